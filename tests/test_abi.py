"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/ecfft_b200.h declares, fails loudly without a GPU, and never routes through the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def lib():
    from ecfft_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ecfft_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ecfft_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from ecfft_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/ecfft_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names


def test_header_cites_the_reference_interface():
    text = open(os.path.join(ROOT, "include", "ecfft_b200.h")).read()
    for cite in ("src/fftree.rs", "src/lib.rs:39-85", "src/fftree.rs:510-660"):
        assert cite in text


def test_status_codes_match_header():
    from ecfft_b200 import _lib
    text = open(os.path.join(ROOT, "include", "ecfft_b200.h")).read()
    codes = dict(re.findall(r"#define (ECFFT_ERR_[A-Z0-9_]+) (\d+)", text))
    assert int(codes["ECFFT_ERR_NOT_POW2"]) == _lib.ERR_NOT_POW2
    assert int(codes["ECFFT_ERR_TREE_TOO_SMALL"]) == _lib.ERR_TREE_TOO_SMALL
    assert int(codes["ECFFT_ERR_BAD_BYTES"]) == _lib.ERR_BAD_BYTES
    assert int(codes["ECFFT_ERR_CUDA"]) == _lib.ERR_CUDA
    assert int(codes["ECFFT_ERR_TOO_LARGE"]) == _lib.ERR_TOO_LARGE
    assert int(codes["ECFFT_ERR_MISSING_TABLES"]) == _lib.ERR_MISSING_TABLES


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(lib):
    import ecfft_b200
    from ecfft_b200 import _lib
    with pytest.raises(ecfft_b200.EcfftError) as e:
        ecfft_b200.build_fftree(64)
    assert e.value.code == _lib.ERR_CUDA
    assert lib.ecfft_last_error()


def test_argument_errors_do_not_need_a_gpu(lib):
    import ecfft_b200
    from ecfft_b200 import _lib
    with pytest.raises(ecfft_b200.EcfftError) as e:
        ecfft_b200.build_fftree(48)
    assert e.value.code == _lib.ERR_NOT_POW2
    h = ctypes.c_void_p()
    assert lib.ecfft_tree_build_secp256k1(0, 0, 0, ctypes.byref(h)) == _lib.ERR_NOT_POW2
    assert lib.ecfft_tree_build_secp256k1(1 << 36, 0, 0, ctypes.byref(h)) == _lib.ERR_TOO_LARGE
    assert lib.ecfft_tree_build_secp256k1(64, 9, 0, ctypes.byref(h)) == _lib.ERR_INVALID_ARG
    assert lib.ecfft_tree_leaves(None) == 0
    lib.ecfft_tree_free(None)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "ecfft_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "oracle" not in text.lower(), f"{fn} mentions the oracle"
    import subprocess
    from ecfft_b200 import _lib
    deps = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in deps


def test_moiety_values_follow_declaration_order():
    from ecfft_b200 import Moiety
    assert int(Moiety.S0) == 0 and int(Moiety.S1) == 1   # src/fftree.rs:17-21
