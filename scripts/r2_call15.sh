# round-2 GPU call 15: per-iteration timeline of spconv_ss_kernel (levels 0 and 2), what-if runs (no weight copies / no row copies / ring depth), host profile of a step
mkdir -p gpurun_out/r2
(SS_LEVELS=0,2 timeout 200 python scripts/ss_timeline.py --trace 2>&1 | tail -150) > gpurun_out/r2/c15_trace.txt
for f in 1 2 3; do echo "== SGB_SS_FLAGS=$f (1: no weight copies, 2: no row copies, 3: neither)"; SGB_SS_FLAGS=$f SS_LEVELS=0,1,2,3 timeout 200 python scripts/ss_timeline.py 2>&1 | tail -12; done > gpurun_out/r2/c15_whatif.txt 2>&1
for s in 3 5; do echo "== SGB_SS_S=$s"; SGB_SS_S=$s SS_LEVELS=0,1,2,3 timeout 200 python scripts/ss_timeline.py 2>&1 | tail -12; done >> gpurun_out/r2/c15_whatif.txt 2>&1
(timeout 200 python scripts/host_profile.py 2>&1 | head -150) > gpurun_out/r2/c15_host_profile.txt
cat gpurun_out/r2/c15_trace.txt gpurun_out/r2/c15_whatif.txt; head -70 gpurun_out/r2/c15_host_profile.txt
