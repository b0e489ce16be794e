# round-2 GPU call 16: tcgen05.mma minimum per k-step across CTAs / issuing warps / fixed B (umma_rate3), ss kernel with the rulebook entries preloaded, bench with the cached plan version
mkdir -p gpurun_out/r2
timeout 120 scripts/experiments/build/umma_rate3 > gpurun_out/r2/c16_umma_rate3.txt 2>&1
(SS_LEVELS=0,1,2,3 timeout 200 python scripts/ss_timeline.py 2>&1 | tail -20) > gpurun_out/r2/c16_ss_ab.txt
(SGB_SS_FLAGS=1 SS_LEVELS=0,1,2,3 timeout 200 python scripts/ss_timeline.py 2>&1 | tail -20) > gpurun_out/r2/c16_ss_ab_nob.txt
(SS_LEVELS=0 timeout 200 python scripts/ss_timeline.py --trace 2>&1 | tail -64 | head -45) > gpurun_out/r2/c16_trace0.txt
(SGB_CONV_SS=0 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null) > gpurun_out/r2/c16_bench_tc.json
(timeout 300 python bench.py --no-cpu-baseline 2>/dev/null) > gpurun_out/r2/c16_bench_default.json
cat gpurun_out/r2/c16_umma_rate3.txt gpurun_out/r2/c16_ss_ab.txt gpurun_out/r2/c16_ss_ab_nob.txt gpurun_out/r2/c16_trace0.txt; cut -c1-400 gpurun_out/r2/c16_bench_tc.json; echo; cut -c1-400 gpurun_out/r2/c16_bench_default.json
