# round-2 GPU call 6: the unified conv path (packed activations between convs, shift 11, fused epilogue) -- all GPU tests + bench
mkdir -p gpurun_out/r2
(timeout 1200 python -m pytest tests -q -m gpu --timeout 300 --tb=short 2>&1 | tail -70) > gpurun_out/r2/c6_tests.txt
(timeout 200 python scripts/conv_levels_ab.py 2>&1 | tail -10) > gpurun_out/r2/c6_levels.txt
(timeout 300 python bench.py 2>gpurun_out/r2/c6_bench.err) > gpurun_out/r2/c6_bench.json
(timeout 300 python bench.py --no-cpu-baseline --workload c2frag 2>gpurun_out/r2/c6_bench_frag.err) > gpurun_out/r2/c6_bench_frag.json
cat gpurun_out/r2/c6_tests.txt gpurun_out/r2/c6_levels.txt
