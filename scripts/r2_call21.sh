# round-2 GPU call 21: host profile of the end-to-end call
mkdir -p gpurun_out/r2
(timeout 300 python scripts/e2e_breakdown.py 2>&1 | head -70) > gpurun_out/r2/c21_e2e.txt
cat gpurun_out/r2/c21_e2e.txt
