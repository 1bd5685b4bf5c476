timeout 900 python -m pytest tests/test_gpu_irregular.py tests/test_gpu_parity.py -x -q > gpurun_out/r02k_pytest_irr.log 2>&1; tail -12 gpurun_out/r02k_pytest_irr.log
