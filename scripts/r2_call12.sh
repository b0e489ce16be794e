# round-2 GPU call 12: first run of the persistent shared-memory-ring conv kernel (spconv_ss.cu): parity tests with it forced on,
# per-level A/B against the register-gather kernel with role counters, CUPTI timeline of one step (idle gaps)
mkdir -p gpurun_out/r2
(SGB_CONV_SS=1 timeout 300 python -m pytest tests/test_gpu_spconv.py -q -m gpu --timeout 120 --tb=short -x 2>&1 | tail -30) > gpurun_out/r2/c12_tests_ss.txt
(timeout 300 python scripts/ss_timeline.py 2>&1 | tail -30) > gpurun_out/r2/c12_ss_ab.txt
(timeout 200 python scripts/step_trace.py 2>&1) > gpurun_out/r2/c12_step_trace.txt
cat gpurun_out/r2/c12_tests_ss.txt gpurun_out/r2/c12_ss_ab.txt; tail -45 gpurun_out/r2/c12_step_trace.txt
