# round-2 GPU call 2: TMA probe v2, the new GPU tests, first run of the TMA-gather conv kernel (under timeouts), A/B, bench
mkdir -p gpurun_out/r2
timeout 150 scripts/bin/tma_probe > gpurun_out/r2/tma_probe2.txt 2>&1
(timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -25) > gpurun_out/r2/c2_tests.txt
(SGB_CONV_IMPL=tma timeout 300 python -m pytest tests/test_gpu_spconv.py -q -m gpu --timeout 120 2>&1 | tail -25) > gpurun_out/r2/c2_tma_tests.txt
(timeout 200 python scripts/conv_levels_ab.py tc tma 2>&1 | tail -12) > gpurun_out/r2/c2_levels.txt
(SGB_CONV_IMPL=tma timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_forward_golden.py -q -m gpu --timeout 120 2>&1 | tail -8) > gpurun_out/r2/c2_tma_model.txt
(timeout 300 python bench.py 2>gpurun_out/r2/c2_bench.err) > gpurun_out/r2/c2_bench.json
(SGB_CONV_IMPL=tma timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/r2/c2_bench_tma.err) > gpurun_out/r2/c2_bench_tma.json
grep "^D:" gpurun_out/r2/tma_probe2.txt | head -70
cat gpurun_out/r2/c2_tests.txt gpurun_out/r2/c2_tma_tests.txt gpurun_out/r2/c2_levels.txt gpurun_out/r2/c2_tma_model.txt
