# compute-sanitizer memcheck over the round-2 late additions: Raman kernels, small-matrix column kernel, single-scatter
# kernel, L1 on the tensor-core kernels (small scenes: memcheck slows kernels ~50x)
set -x
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_rrs.py -q -m gpu -x \
  -k "small_N and IQU-False or owner_range or brdf or wide_kernels" 2>&1 | tail -15
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_sources.py -q -m gpu -x \
  -k "small_matrix or rt_run_ss or (rt_layer and IQUV)" 2>&1 | tail -15
