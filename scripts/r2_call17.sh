# round-2 GPU call 17: spconv_ss_kernel with pair stages: parity (forced on), per-level A/B with counters, product-library A/B, bench
mkdir -p gpurun_out/r2
(SGB_CONV_SS=1 timeout 400 python -m pytest tests/test_gpu_spconv.py -q -m gpu --timeout 120 --tb=line 2>&1 | tail -8) > gpurun_out/r2/c17_tests_ss.txt
(SS_LEVELS=0,1,2,3,4 timeout 200 python scripts/ss_timeline.py 2>&1 | tail -20) > gpurun_out/r2/c17_ss_ab.txt
(SS_LEVELS=0,1,2,3,4 timeout 200 python scripts/ss_timeline.py --no-tl 2>&1 | tail -20) > gpurun_out/r2/c17_ss_ab_product.txt
(timeout 300 python bench.py --no-cpu-baseline 2>/dev/null) > gpurun_out/r2/c17_bench_default.json
cat gpurun_out/r2/c17_tests_ss.txt gpurun_out/r2/c17_ss_ab.txt gpurun_out/r2/c17_ss_ab_product.txt; cut -c1-400 gpurun_out/r2/c17_bench_default.json
