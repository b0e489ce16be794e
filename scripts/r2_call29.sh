# round-2 GPU call 29: why c3 finds no proposals and c5 fails in pyramid_inverse_map at full size
mkdir -p gpurun_out/r2
(timeout 300 python scripts/debug_workload.py c3 2>&1 | tail -40) > gpurun_out/r2/c29_c3.txt
(timeout 300 python scripts/debug_workload.py c5 2>&1 | tail -60) > gpurun_out/r2/c29_c5.txt
cat gpurun_out/r2/c29_c3.txt gpurun_out/r2/c29_c5.txt
