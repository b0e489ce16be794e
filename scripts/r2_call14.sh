# round-2 GPU call 14: spconv_ss_kernel after the ring-phase fix: all conv parity tests with it forced on, per-level A/B + role counters, step timeline
mkdir -p gpurun_out/r2
(SGB_CONV_SS=1 timeout 400 python -m pytest tests/test_gpu_spconv.py -q -m gpu --timeout 120 --tb=line 2>&1 | tail -15) > gpurun_out/r2/c14_tests_ss.txt
(timeout 300 python scripts/ss_timeline.py 2>&1 | tail -30) > gpurun_out/r2/c14_ss_ab.txt
(timeout 200 python scripts/step_trace.py 2>&1) > gpurun_out/r2/c14_step_trace.txt
cat gpurun_out/r2/c14_tests_ss.txt gpurun_out/r2/c14_ss_ab.txt; tail -45 gpurun_out/r2/c14_step_trace.txt
