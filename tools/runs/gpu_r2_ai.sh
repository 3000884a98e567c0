#!/bin/bash
# round 2: compute-sanitizer on the final code (TMA tile load + mbarrier, folded combines, conditional launches, m31)
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
python tools/sanitize_small.py 2>&1 | tail -1
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_ai_sanitize_memcheck.log python tools/sanitize_small.py 2>&1 | tail -1
tail -3 gpurun_out/r02_ai_sanitize_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_ai_sanitize_racecheck.log python tools/sanitize_small.py 2>&1 | tail -1
tail -3 gpurun_out/r02_ai_sanitize_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --log-file gpurun_out/r02_ai_sanitize_synccheck.log python tools/sanitize_small.py 2>&1 | tail -1
tail -3 gpurun_out/r02_ai_sanitize_synccheck.log
