timeout 900 python -m pytest tests/test_gpu_irregular.py tests/test_gpu_setup_kernels.py -x -q > gpurun_out/r02l_pytest_irr.log 2>&1; tail -12 gpurun_out/r02l_pytest_irr.log
