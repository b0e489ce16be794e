# round-2 GPU call 37: GPU occupancy with scans in flight (CUPTI)
mkdir -p gpurun_out/r2
(timeout 300 python scripts/inflight_trace.py 3 2>&1 | tail -16) > gpurun_out/r2/c37_inflight3.txt
(timeout 300 python scripts/inflight_trace.py 1 2>&1 | tail -16) > gpurun_out/r2/c37_inflight1.txt
cat gpurun_out/r2/c37_inflight3.txt gpurun_out/r2/c37_inflight1.txt
