# round-2 GPU call 42: compute-sanitizer memcheck over the kernels touched this round (BFS select/push/emit, grouping selection, ball query loops, both conv kernels at small size)
mkdir -p gpurun_out/r2
(timeout 900 compute-sanitizer --tool memcheck --print-limit 3 --error-exitcode 7 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 800 --tb=line -k "bfs_cluster_random or bfs_cluster_edges or bfs_cluster_capped or group_entries or ballquery" 2>&1 | grep -v "^$" | tail -12) > gpurun_out/r2/c42_memcheck_ops.txt
(timeout 600 compute-sanitizer --tool memcheck --print-limit 3 --error-exitcode 7 python -m pytest tests/test_gpu_spconv.py -q -m gpu --timeout 500 --tb=line -k "test_subm_conv_vs_oracle or fused_act or down_and_inverse or beyond_fp16" 2>&1 | grep -v "^$" | tail -12) > gpurun_out/r2/c42_memcheck_conv.txt
cat gpurun_out/r2/c42_memcheck_ops.txt gpurun_out/r2/c42_memcheck_conv.txt
