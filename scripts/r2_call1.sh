mkdir -p gpurun_out/r2
(timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5) > gpurun_out/r2/c1_tests.txt
(timeout 300 python -m pytest tests/pending_gpu_forward_golden.py -q -m gpu 2>&1 | tail -40) > gpurun_out/r2/c1_pending.txt
(SGB_BFS_MODE=1 timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -m gpu 2>&1 | tail -15) > gpurun_out/r2/c1_bfs1.txt
(SGB_TC_SPLIT_POLICY=1 timeout 300 python -m pytest tests/test_gpu_spconv.py tests/test_gpu_model.py -q -m gpu 2>&1 | tail -15) > gpurun_out/r2/c1_split1.txt
(SGB_TC_GATHER=1 timeout 300 python -m pytest tests/test_gpu_spconv.py tests/test_gpu_model.py -q -m gpu 2>&1 | tail -15) > gpurun_out/r2/c1_gather1.txt
(timeout 400 python scripts/tc_tuning_ab.py 2>&1 | tail -40) > gpurun_out/r2/c1_ab.txt
(timeout 300 python bench.py 2>gpurun_out/r2/c1_bench.err) > gpurun_out/r2/c1_bench.json
cat gpurun_out/r2/c1_*.txt
