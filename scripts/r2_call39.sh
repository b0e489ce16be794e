# round-2 GPU call 39: per-slot release of the ring (commit after each slot's products): conv parity + per-level A/B (product build)
mkdir -p gpurun_out/r2
(timeout 600 python -m pytest tests/test_gpu_spconv.py -q -m gpu --timeout 120 --tb=line 2>&1 | tail -3) > gpurun_out/r2/c39_tests.txt
(SS_LEVELS=0,1,2,3,4 timeout 200 python scripts/ss_timeline.py --no-tl 2>&1 | tail -6) > gpurun_out/r2/c39_ss_ab.txt
(SS_LEVELS=0,1,2,3,4 timeout 200 python scripts/ss_timeline.py --no-tl 2>&1 | tail -6) >> gpurun_out/r2/c39_ss_ab.txt
cat gpurun_out/r2/c39_tests.txt gpurun_out/r2/c39_ss_ab.txt
