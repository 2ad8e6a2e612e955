mkdir -p gpurun_out
AD4CL_CUDA_LIB=$PWD/variants/threephase.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_divergence.py tests/test_gpu_fullsize.py tests/test_gpu_general_epilogue.py -m gpu -q --tb=short -x 2>&1 | tail -5
bash profiles/scripts/r02_ab2.sh
