# compute-sanitizer racecheck (shared-memory hazards) over the warp-level kernels: Raman elastic / unit kernels, small-matrix kernel
set -x
compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 python -m pytest tests/test_gpu_rrs.py -q -m gpu -x \
  -k "small_N and IQU-False or brdf" 2>&1 | tail -25
compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
  -k "small_matrix" 2>&1 | tail -25
