# round-2 GPU call 19: ss kernel with the next pair's barrier wait between the slots; per-level A/B (product build); launch-by-launch A/B inside the real step
mkdir -p gpurun_out/r2
(SGB_CONV_SS=1 timeout 400 python -m pytest tests/test_gpu_spconv.py -q -m gpu --timeout 120 --tb=line 2>&1 | tail -4) > gpurun_out/r2/c19_tests_ss.txt
(SS_LEVELS=0,1,2,3,4 timeout 200 python scripts/ss_timeline.py --no-tl 2>&1 | tail -8) > gpurun_out/r2/c19_ss_ab_product.txt
(SS_LEVELS=0,2 timeout 200 python scripts/ss_timeline.py 2>&1 | tail -8) > gpurun_out/r2/c19_ss_ab_tl.txt
for m in 0 1; do SGB_CONV_SS=$m timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2/c19_launches_ss$m.csv python scripts/one_step.py 1 > /dev/null 2>&1; done
python scripts/conv_launch_ab.py gpurun_out/r2/c19_launches_ss0.csv gpurun_out/r2/c19_launches_ss1.csv > gpurun_out/r2/c19_launch_ab.txt 2>&1
cat gpurun_out/r2/c19_tests_ss.txt gpurun_out/r2/c19_ss_ab_product.txt gpurun_out/r2/c19_ss_ab_tl.txt gpurun_out/r2/c19_launch_ab.txt
