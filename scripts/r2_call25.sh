# round-2 GPU call 25: full GPU suite, smoke(), bench (default line incl. cpu baseline + reference GPU ops), reference arm
mkdir -p gpurun_out/r2
(timeout 1500 python -m pytest tests -q -m gpu --timeout 300 --tb=short 2>&1 | tail -12) > gpurun_out/r2/c25_tests.txt
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/r2/c25_smoke.txt
(timeout 600 python bench.py 2>gpurun_out/r2/c25_bench.err) > gpurun_out/r2/c25_bench.json
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null) > gpurun_out/r2/c25_bench_ref.json
cat gpurun_out/r2/c25_tests.txt gpurun_out/r2/c25_smoke.txt; cut -c1-700 gpurun_out/r2/c25_bench.json; echo; tail -3 gpurun_out/r2/c25_bench.err; cut -c1-500 gpurun_out/r2/c25_bench_ref.json
