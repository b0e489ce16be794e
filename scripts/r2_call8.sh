# round-2 GPU call 8: plan executor (one C call per U-Net), two-pass ball query, reference model on torch>=2 -- all tests + bench
mkdir -p gpurun_out/r2
(timeout 1200 python -m pytest tests -q -m gpu --timeout 300 --tb=short 2>&1 | tail -70) > gpurun_out/r2/c8_tests.txt
(timeout 300 python bench.py 2>gpurun_out/r2/c8_bench.err) > gpurun_out/r2/c8_bench.json
(timeout 300 python bench.py --no-cpu-baseline --workload c2frag 2>gpurun_out/r2/c8_bench_frag.err) > gpurun_out/r2/c8_bench_frag.json
cat gpurun_out/r2/c8_tests.txt
