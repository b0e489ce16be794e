# round-2 GPU call 5: UMMA issue/latency micro-benchmark, reference-model drop-in tests, fused (packed) graph on the ss kernel
mkdir -p gpurun_out/r2
(timeout 120 python scripts/umma_rate.py 2>&1 | tail -40) > gpurun_out/r2/c5_umma.txt
(timeout 300 python -m pytest tests/test_gpu_reference_model.py -q -m gpu --timeout 200 --tb=short 2>&1 | tail -60) > gpurun_out/r2/c5_refmodel.txt
(SGB_CONV_IMPL=ss timeout 400 python -m pytest tests/test_gpu_spconv.py tests/test_gpu_model.py tests/test_gpu_forward_golden.py tests/test_gpu_reference_model.py -q -m gpu --timeout 150 --tb=short 2>&1 | tail -40) > gpurun_out/r2/c5_ss_fused.txt
(SGB_CONV_IMPL=ss timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/r2/c5_bench_ss.err) > gpurun_out/r2/c5_bench_ss.json
(timeout 300 python -m pytest tests/test_gpu_dataprep.py tests/test_gpu_evaluation.py -q -m gpu --timeout 150 --tb=short 2>&1 | tail -30) > gpurun_out/r2/c5_prep_eval.txt
cat gpurun_out/r2/c5_umma.txt gpurun_out/r2/c5_refmodel.txt gpurun_out/r2/c5_ss_fused.txt gpurun_out/r2/c5_prep_eval.txt
