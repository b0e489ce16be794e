# round-2 GPU call 10 (session 2): recover the measured state after the container was re-created -- all tests, bench (c2 + c2frag), ncu launch list
mkdir -p gpurun_out/r2
(timeout 1200 python -m pytest tests -q -m gpu --timeout 300 --tb=short 2>&1 | tail -40) > gpurun_out/r2/c10_tests.txt
(timeout 400 python bench.py 2>gpurun_out/r2/c10_bench.err) > gpurun_out/r2/c10_bench.json
(timeout 300 python bench.py --no-cpu-baseline --workload c2frag 2>gpurun_out/r2/c10_bench_frag.err) > gpurun_out/r2/c10_bench_frag.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2/c10_launches.csv python scripts/one_step.py 1 > /dev/null 2>&1
cat gpurun_out/r2/c10_tests.txt; head -c 1500 gpurun_out/r2/c10_bench.json; tail -3 gpurun_out/r2/c10_bench.err
