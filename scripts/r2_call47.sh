# round-2 GPU call 47: GPU occupancy of the END-TO-END path with 3 scans in flight
mkdir -p gpurun_out/r2
(timeout 300 python scripts/inflight_trace.py 3 e2e 2>&1 | tail -15) > gpurun_out/r2/c47_inflight3_e2e.txt
(timeout 300 python scripts/inflight_trace.py 3 2>&1 | tail -15) > gpurun_out/r2/c47_inflight3_dev.txt
cat gpurun_out/r2/c47_inflight3_e2e.txt gpurun_out/r2/c47_inflight3_dev.txt
