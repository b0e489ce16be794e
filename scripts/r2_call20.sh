# round-2 GPU call 20: all GPU tests (both conv kernels parametrised, fused grouping selection), bench, step timeline
mkdir -p gpurun_out/r2
(timeout 1500 python -m pytest tests -q -m gpu --timeout 300 --tb=short 2>&1 | tail -25) > gpurun_out/r2/c20_tests.txt
(timeout 300 python bench.py --no-cpu-baseline 2>/dev/null) > gpurun_out/r2/c20_bench.json
(timeout 300 python bench.py --no-cpu-baseline --workload c2frag 2>/dev/null) > gpurun_out/r2/c20_bench_frag.json
(timeout 200 python scripts/step_trace.py 2>&1) > gpurun_out/r2/c20_step_trace.txt
cat gpurun_out/r2/c20_tests.txt; cut -c1-330 gpurun_out/r2/c20_bench.json; echo; cut -c1-330 gpurun_out/r2/c20_bench_frag.json; echo; grep "^# total" gpurun_out/r2/c20_step_trace.txt
