./tools/microbench/graph_while > gpurun_out/r02g_graph_while.log 2>&1; cat gpurun_out/r02g_graph_while.log
python tools/grid_gpu.py --out gpurun_out/grid_gpu_r02g.npz > gpurun_out/r02g_grid_gpu.log 2>&1; cat gpurun_out/r02g_grid_gpu.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1; tail -5 gpurun_out/r02g_pytest.log
