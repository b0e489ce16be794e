# round-2 GPU call 3: all GPU tests, TMA-gather kernel parity, per-level A/B in the original and in Z-order, bench
mkdir -p gpurun_out/r2
(timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -30) > gpurun_out/r2/c3_tests.txt
(SGB_CONV_IMPL=tma timeout 300 python -m pytest tests/test_gpu_spconv.py -q -m gpu --timeout 120 2>&1 | tail -25) > gpurun_out/r2/c3_tma_tests.txt
(timeout 200 python scripts/conv_levels_ab.py tc tma 2>&1 | tail -12) > gpurun_out/r2/c3_levels.txt
(timeout 200 python scripts/conv_levels_ab.py tc tma --morton 2>&1 | tail -12) > gpurun_out/r2/c3_levels_morton.txt
(timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/r2/c3_bench.err) > gpurun_out/r2/c3_bench.json
(timeout 300 python bench.py --no-cpu-baseline --workload c2frag 2>gpurun_out/r2/c3_bench_frag.err) > gpurun_out/r2/c3_bench_frag.json
cat gpurun_out/r2/c3_tests.txt gpurun_out/r2/c3_tma_tests.txt gpurun_out/r2/c3_levels.txt gpurun_out/r2/c3_levels_morton.txt
