# round-2 GPU call 9: segmented plan executor, reference model with the torch>=2 shim -- all tests + bench
mkdir -p gpurun_out/r2
(timeout 1200 python -m pytest tests -q -m gpu --timeout 300 --tb=short 2>&1 | tail -60) > gpurun_out/r2/c9_tests.txt
(timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/r2/c9_bench.err) > gpurun_out/r2/c9_bench.json
(timeout 300 python bench.py --no-cpu-baseline --workload c2frag 2>gpurun_out/r2/c9_bench_frag.err) > gpurun_out/r2/c9_bench_frag.json
cat gpurun_out/r2/c9_tests.txt
