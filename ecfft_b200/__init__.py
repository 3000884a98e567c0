"""ecfft_b200 — B200-native ECFFT engine behind the reference crate's FFTree<secp256k1::Fp> surface."""
from ._lib import EcfftError, LIB_PATH  # noqa: F401
from .fftree import FFTree, Moiety, build_fftree, PARTS_FULL, PARTS_ENTER_ONLY  # noqa: F401
from .poly import poly_mul  # noqa: F401
from . import m31  # noqa: F401  (FFTree<m31::Fp>, the reference's second field)
