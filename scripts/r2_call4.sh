# round-2 GPU call 4: cp.async-ring conv kernel (parity + per-level A/B), reference-model test details, BFS with L1-cached reads
mkdir -p gpurun_out/r2
(timeout 300 python -m pytest tests/test_gpu_reference_model.py -q -m gpu --timeout 200 --tb=short 2>&1 | tail -90) > gpurun_out/r2/c4_refmodel.txt
(SGB_CONV_IMPL=ss timeout 300 python -m pytest tests/test_gpu_spconv.py -q -m gpu --timeout 120 2>&1 | tail -25) > gpurun_out/r2/c4_ss_tests.txt
(timeout 200 python scripts/conv_levels_ab.py tc ss 2>&1 | tail -12) > gpurun_out/r2/c4_levels.txt
(timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_vs_reference.py -q -m gpu --timeout 200 2>&1 | tail -15) > gpurun_out/r2/c4_ops_tests.txt
(SGB_CONV_IMPL=ss timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_forward_golden.py -q -m gpu --timeout 120 2>&1 | tail -8) > gpurun_out/r2/c4_ss_model.txt
(timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/r2/c4_bench.err) > gpurun_out/r2/c4_bench.json
(SGB_CONV_IMPL=ss timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/r2/c4_bench_ss.err) > gpurun_out/r2/c4_bench_ss.json
cat gpurun_out/r2/c4_ss_tests.txt gpurun_out/r2/c4_levels.txt gpurun_out/r2/c4_ops_tests.txt gpurun_out/r2/c4_ss_model.txt
