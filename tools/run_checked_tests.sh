#!/bin/bash
# GPU tests against the bounds-checked library (make -C yagre_mcmc_b200/csrc checked), then the proof that its device
# asserts are live.  One gpurun call:  bash tools/run_checked_tests.sh > gpurun_out/checked_tests.log 2>&1
set -u
export YAGRE_B200_LIB=$PWD/tools/_build/libyagre_b200_checked.so
python - <<'PY'
from yagre_mcmc_b200 import _lib
print("library under test:", _lib.LIB_PATH)
PY
python -m pytest tests -m gpu -q 2>&1 | tail -4
python - <<'PY'
import ctypes, os
lib = ctypes.CDLL(os.environ["YAGRE_B200_LIB"])
print("deliberate violation of YG_CHK: device assert fired =", bool(lib.yg_bounds_check_selftest()))
PY
