# round-2 GPU call 31: host CPU time per end-to-end scan by function; new full-size tests
mkdir -p gpurun_out/r2
(timeout 300 python scripts/host_cpu_profile.py 2>&1 | head -130) > gpurun_out/r2/c31_cpu.txt
(timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 --tb=short -k full_size 2>&1 | tail -8) > gpurun_out/r2/c31_tests.txt
cat gpurun_out/r2/c31_tests.txt; head -110 gpurun_out/r2/c31_cpu.txt
