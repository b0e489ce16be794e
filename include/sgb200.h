/*
 * sgb200.h -- C ABI of libsgb200.so, the B200 (sm_100a) implementation of SoftGroup's per-scan
 * inference hot path (reference: thangvubk/SoftGroup, paths below are relative to its root).
 *
 * Boundary rules
 *   - extern "C", plain pointers and sizes, no torch / ATen types.
 *   - Every pointer named d_* is DEVICE memory; h_* is HOST memory. `stream` is a cudaStream_t
 *     passed as void*. All work is enqueued on `stream`; functions documented as "blocking"
 *     synchronise that stream before returning (they return a count the caller needs to size
 *     the next allocation, exactly where the reference blocks on cudaMemcpy D2H).
 *   - The library owns no memory between calls. Scratch is caller-provided: ask
 *     sgb_*_workspace_bytes() and pass a device buffer of at least that size.
 *   - Return value: >= 0 on success (a count where documented), < 0 = SGB_ERR_*; the message is
 *     available from sgb_last_error(). Nothing falls back to the CPU.
 *   - Out-parameter convention follows the reference extension (softgroup/ops/src/softgroup_api.cpp:6-29):
 *     the caller pre-allocates outputs. Where the reference resize_()s an empty tensor inside the
 *     callee (voxelize.cpp:29-33, bfs_cluster.cpp:119-122) the C ABI is two-phase:
 *     *_count (returns sizes) then *_fill.
 */
#ifndef SGB200_H
#define SGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGB_OK 0
#define SGB_ERR_CUDA (-1)      /* a CUDA call failed; see sgb_last_error() */
#define SGB_ERR_ARG (-2)       /* bad argument (null pointer, negative size, unsupported mode) */
#define SGB_ERR_WORKSPACE (-3) /* workspace too small */
#define SGB_ERR_RANGE (-4)     /* input outside the packed-key range (see each function) */
#define SGB_ERR_OVERFLOW (-5)  /* fixed-capacity structure overflowed */

#define SGB_MAX_NEIGHBORS 1000 /* softgroup/ops/src/bfs_cluster/bfs_cluster.cu:24, octree_ball_query.cu:11 */

const char *sgb_last_error(void);
/* ABI version of this header (bumped on any signature change). */
int sgb_abi_version(void);
/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches evidence). */
long long sgb_launch_count(void);
/* 1 if a CUDA device is usable, 0 otherwise (never initialises a context when none exists). */
int sgb_device_available(void);

/* ---------------------------------------------------------------------------------------------
 * voxelize_idx -- replaces voxelize_idx_3d (softgroup/ops/src/softgroup_ops.cpp:13-19) =
 * voxelize_idx<3> / voxelize_inputmap / voxelize_outputmap (voxelize/voxelize.cpp:11-165).
 * coords: int64 [N, ncol], ncol = 4 (batch,x,y,z) or 3 (x,y,z). Voxel id = first-occurrence rank
 * in point order; output_map rows are [count, p0, p1, ..., 0-pad] with ascending point index.
 * mode: 0 unique, 1 first, 2 last, 3 sum, 4 mean (only the map differs; 4 is what the model uses).
 *
 * GPU path: keys are packed into 64 bits: batch in [0, 65535], x,y,z in [-32768, 32767]
 * (all reference configs fit); anything outside returns SGB_ERR_RANGE.
 *   count (blocking): fills d_input_map [N]; returns M and maxActive through h_M / h_maxActive.
 *   fill: d_output_coords int64 [M, ncol], d_output_map int32 [M, maxActive+1]. Must be called with
 *         the SAME workspace contents left by count.
 * CPU path (host pointers, for DataLoader workers -- softgroup/data/custom.py:239; never touches
 * CUDA): sgb_voxelize_idx_cpu_begin / _finish with the same semantics and no key-range limit.
 * ------------------------------------------------------------------------------------------- */
size_t sgb_voxelize_idx_workspace_bytes(int N);
int sgb_voxelize_idx_count(const int64_t *d_coords, int N, int ncol, int mode, int32_t *d_input_map, void *d_ws,
                           size_t ws_bytes, int *h_M, int *h_maxActive, void *stream);
int sgb_voxelize_idx_fill(const int64_t *d_coords, int N, int ncol, int mode, int M, int maxActive,
                          int64_t *d_output_coords, int32_t *d_output_map, void *d_ws, size_t ws_bytes, void *stream);
void *sgb_voxelize_idx_cpu_begin(const int64_t *h_coords, int N, int ncol, int mode, int32_t *h_input_map, int *h_M,
                                 int *h_maxActive);
int sgb_voxelize_idx_cpu_finish(void *handle, const int64_t *h_coords, int64_t *h_output_coords,
                                int32_t *h_output_map);

/* voxelize_fp / voxelize_bp -- replace voxelize_fp_feat / voxelize_bp_feat (softgroup_ops.cpp:21-38,
 * voxelize/voxelize.cu:9-62). feats f32 [N,C], rules int32 [M, maxActive+1], out f32 [M,C].
 * fp overwrites d_out (no pre-zeroing needed); bp ACCUMULATES into d_d_feats like the reference. */
int sgb_voxelize_fp(const float *d_feats, float *d_out, const int32_t *d_rules, int mode, int M, int maxActive, int C,
                    void *stream);
int sgb_voxelize_bp(const float *d_d_out, float *d_d_feats, const int32_t *d_rules, int mode, int M, int maxActive,
                    int C, void *stream);

/* ---------------------------------------------------------------------------------------------
 * ballquery_batch_p -- replaces ballquery_batch_p_cuda (bfs_cluster/bfs_cluster.cu:68-101, kernel :15-66).
 * xyz f32 [n,3]; batch_idxs int32 [n]; batch_offsets int32 [B+1] (non-decreasing; query i scans the index
 * range of its batch). Per point: neighbours with d2 < radius^2 (strict, fp32, d2 evaluated as
 * fma(dz,dz,fma(dx,dx,dy*dy)) like the compiled reference), ascending point index, self included, first
 * 1000 kept. d_idx has capacity n*meanActive entries; lists that would cross the capacity are truncated
 * exactly like the reference (:55-61) and the caller relaunches with a larger meanActive
 * (softgroup/ops/functions.py:258-266). d_start_len int32 [n,2] = (start, count).
 * Returns (blocking) the total neighbour count, like the reference's `return cumsum`.
 * The *_async form never synchronises: it leaves d_total[0] = total, d_total[1] = range-error flag (int32 [2], device).
 * Points with a NaN / Inf coordinate get the empty list and are nobody's neighbour -- the reference's result, every
 * comparison with them being false (:38-44). B must be <= 1023 and finite |xyz/radius| < 131070: SGB_ERR_RANGE
 * otherwise (blocking form), or the flag in d_total[1] (async form; sgb_bfs_cluster_count reports it).
 * capacity must stay below 2^31 entries (int32 cursor).
 * ------------------------------------------------------------------------------------------- */
size_t sgb_ballquery_workspace_bytes(int n);
long long sgb_ballquery_batch_p(int n, int meanActive, float radius, const float *d_xyz, const int32_t *d_batch_idxs,
                                const int32_t *d_batch_offsets, int B, int32_t *d_idx, int32_t *d_start_len,
                                void *d_ws, size_t ws_bytes, void *stream);
int sgb_ballquery_batch_p_async(int n, long long capacity, float radius, const float *d_xyz,
                                const int32_t *d_batch_idxs, const int32_t *d_batch_offsets, int B, int32_t *d_idx,
                                int32_t *d_start_len, int32_t *d_total, void *d_ws, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Octree ball query (SoftGroup++) -- replaces build_and_export_octree (octree_ball_query/octree_ball_query.cpp:150-165,
 * CPU pointer tree) and octree_ball_query (octree_ball_query.cu:128-147). Fixed 3 levels: boxes f32 [585,6] in BFS
 * order from xyzwhl = (centre, extent) f32 [6]; pt_inds int32 [n] (leaf-major, ascending inside a leaf);
 * pt_start_len int32 [512,2]. Query: neighbours in LEAF-MAJOR order, first 1000 kept; d_out_inds has capacity
 * n*mean_active with the reference's truncation; returns (blocking) the total count, like the reference.
 * The tree ignores batch indices (batch size 1 only), like the reference.
 * ------------------------------------------------------------------------------------------- */
size_t sgb_octree_workspace_bytes(int n);
int sgb_octree_build(const float *d_points, int n, const float *d_xyzwhl, float *d_boxes, int32_t *d_pt_inds,
                     int32_t *d_pt_start_len, void *d_ws, size_t ws_bytes, void *stream);
long long sgb_octree_ball_query(const float *d_points, const float *d_boxes, const int32_t *d_pt_inds,
                                const int32_t *d_pt_start_len, int n, int mean_active, float radius, int32_t *d_out_inds,
                                int32_t *d_out_start_len, int32_t *d_total, void *stream);

/* ---------------------------------------------------------------------------------------------
 * bfs_cluster -- replaces bfs_cluster / get_clusters / find_cc / fill_cluster_idxs_
 * (bfs_cluster/bfs_cluster.cpp:33-126), on the GPU, bit-exact including BFS visitation order.
 * d_ball_query_idxs int32 [nActive], d_start_len int32 [N,2] (DEVICE memory; the reference takes CPU
 * tensors). threshold semantics (:70-82): a component is kept when (float)size >= thr, where the caller
 * passes thr = threshold if class_numpoint_mean[class_id] == -1 else threshold * mean.
 *   count (blocking): returns nCluster, writes sumNPoint to *h_sumNPoint and the longest list length to *h_maxLen.
 *   fill: d_cluster_idxs int32 [sumNPoint,2] (cluster id, point idx), d_cluster_offsets int32 [nCluster+1].
 * Optional per-node segments (d_node_seg int32 [N], d_seg_thr f32 [nSeg]) give every node the threshold of
 * its segment (used to cluster all classes of a scan in one call); pass NULL for a single threshold.
 * The labelling is the exact directed one (lists cut by the 1000 cap make the graph asymmetric); label passes are
 * enqueued in batches and the host waits ONCE, for the totals. d_upstream_err (nullable): device int32 error flag of
 * the asynchronous ball query that produced the lists (d_total + 1 of sgb_ballquery_batch_p_async); it is read at that
 * same wait and reported as SGB_ERR_RANGE.
 * ------------------------------------------------------------------------------------------- */
size_t sgb_bfs_cluster_workspace_bytes(int N);
int sgb_bfs_cluster_count(const int32_t *d_ball_query_idxs, const int32_t *d_start_len, int N, float thr,
                          const int32_t *d_node_seg, const float *d_seg_thr, const int32_t *d_upstream_err, void *d_ws,
                          size_t ws_bytes, int *h_sumNPoint, int *h_maxLen, void *stream);
/* scratch for fill: one bitmap row of ceil(maxLen/32) words per emitted point (0 when maxLen > 2048: fallback path) */
size_t sgb_bfs_cluster_scratch_bytes(int sumNPoint, int maxLen);
int sgb_bfs_cluster_fill(const int32_t *d_ball_query_idxs, const int32_t *d_start_len, int N, int nCluster,
                         int sumNPoint, int maxLen, int32_t *d_cluster_idxs, int32_t *d_cluster_offsets, void *d_ws,
                         size_t ws_bytes, void *d_scratch, size_t scratch_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Segment reductions -- replace sec_mean / sec_min / sec_max (sec_mean/sec_mean.cu:13-93) and
 * global_avg_pool_fp / _bp (roipool/roipool.cu:12-72). inp f32 [S,C]; offsets int32 [nP+1]; out f32 [nP,C].
 * min/max are exact; mean / avg-pool use a tree reduction (within 1e-5 rel of the sequential sums).
 * ------------------------------------------------------------------------------------------- */
int sgb_sec_mean(const float *d_inp, const int32_t *d_offsets, float *d_out, int nProposal, int C, void *stream);
int sgb_sec_min(const float *d_inp, const int32_t *d_offsets, float *d_out, int nProposal, int C, void *stream);
int sgb_sec_max(const float *d_inp, const int32_t *d_offsets, float *d_out, int nProposal, int C, void *stream);
int sgb_global_avg_pool_fp(const float *d_feats, const int32_t *d_offsets, float *d_out, int nProposal, int C,
                           void *stream);
int sgb_global_avg_pool_bp(float *d_d_feats, const int32_t *d_offsets, const float *d_d_out, int nProposal, int C,
                           void *stream);

/* ---------------------------------------------------------------------------------------------
 * Mask IoU / labels -- replace get_mask_iou_on_cluster / get_mask_iou_on_pred / get_mask_label
 * (cal_iou_and_masklabel/cal_iou_and_masklabel.cu:9-164). Training-side ops of the binding surface.
 * proposals_idx here is the POINT-INDEX column (int32 [sumNPoint]); instance_labels int64 [N];
 * d_mask_scores_sigmoid may be NULL (on_cluster). iou f32 [nProposal, nInstance].
 * ------------------------------------------------------------------------------------------- */
int sgb_get_mask_iou(const int32_t *d_proposals_idx, const int32_t *d_proposals_offset,
                     const int64_t *d_instance_labels, const int32_t *d_instance_pointnum,
                     const float *d_mask_scores_sigmoid, float *d_proposals_iou, int nInstance, int nProposal,
                     void *stream);
int sgb_get_mask_label(const int32_t *d_proposals_idx, const int32_t *d_proposals_offset,
                       const int64_t *d_instance_labels, const int64_t *d_instance_cls,
                       const float *d_proposals_iou, int nInstance, int nProposal, float iou_thr,
                       float *d_mask_label, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Sparse convolution -- replaces what the reference delegates to spconv 2.x (call sites
 * softgroup/model/blocks.py:31-41,50-70,96-129; softgroup/model/softgroup.py:60-65,73-74).
 * indices int32 [M,4] (b,x,y,z), b in [0,65535], x,y,z in [0,32766].
 *
 * Rulebooks ("maps"): int32 [K, Mout]; map[k][j] = input row feeding output row j through kernel offset k,
 * or -1. Kernel offset order = row-major (k0,k1,k2) like the weight layout [out,k0,k1,k2,in]
 * (tools/convert_checkpoint.py:17-19).
 *   subm3:   K=27, Mout=M, map[k][j] = row of indices[j] + (k0-1,k1-1,k2-1)           (SubMConv3d k3 p1)
 *   down2:   K=8; output voxels = distinct (b, x>>1, y>>1, z>>1) with every coordinate < out_shape
 *            = floor(shape/2) (inputs on the max plane of an odd dim are dropped), numbered by first
 *            occurrence in input order. count (blocking) returns Mout; fill writes d_out_indices [Mout,4],
 *            d_map [8,Mout] (children) and d_inv_map [8,M] (inv_map[k][i] = parent row if input i is child k
 *            of it, else -1) -- the inverse conv is the same kernel run with inv_map.   (SparseConv3d k2 s2,
 *            SparseInverseConv3d k2)
 * ------------------------------------------------------------------------------------------- */
size_t sgb_rulebook_workspace_bytes(int M);
int sgb_rulebook_subm3(const int32_t *d_indices, int M, int32_t *d_map, void *d_ws, size_t ws_bytes, void *stream);
int sgb_rulebook_down2_count(const int32_t *d_indices, int M, const int32_t *h_spatial_shape /*[3]*/, void *d_ws,
                             size_t ws_bytes, void *stream);
int sgb_rulebook_down2_fill(const int32_t *d_indices, int M, int Mout, int32_t *d_out_indices, int32_t *d_map,
                            int32_t *d_inv_map, void *d_ws, size_t ws_bytes, void *stream);

/* out[j, out_off + n] = sum_k sum_c act(in[map[k][j], in_off + c]) * W[k][c][n]  (+ residual[j, res_off + n]) (+ bias[n])
 * act(x) = x                       if d_in_scale == NULL
 *        = max(x*scale[c]+shift[c], 0)   otherwise (eval-mode BatchNorm1d folded to scale/shift, then ReLU --
 *          the pre-activation of blocks.py:55-70); absent neighbours contribute exactly 0.
 * W is f32 [K, Cin, Cout] (the wrapper permutes the reference layout once). d_map == NULL means K == 1 and the
 * identity map (Custom1x1Subm3d, blocks.py:31-41, and nn.Linear). Row strides are in floats.
 * fp32 accumulate on CUDA cores (fp32-exact products; see DESIGN.md on why not plain TF32). */
int sgb_spconv_forward(const float *d_in, int in_stride, int in_off, const int32_t *d_map, int K, int Mout,
                       const float *d_W, int Cin, int Cout, const float *d_in_scale, const float *d_in_shift,
                       const float *d_residual, int res_stride, int res_off, const float *d_bias, float *d_out,
                       int out_stride, int out_off, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core sparse convolution (spconv_tc.cu; tcgen05.mma kind::f16, fp32 accumulators in TMEM) -- the production path
 * for Cin <= 512, Cout <= 256. Every operand value x is carried as two fp16 numbers hi = fp16(x),
 * lo = fp16((x - hi) * 2^shift), shift = sgb_spconv_lo_shift(), and the product is hi*hi + hi*lo + lo*hi: fp32-grade
 * (DESIGN.md 3.2). Activations travel PACKED between convolutions: per row and 32-channel chunk 16 words of fp16 hi pairs
 * then 16 words of fp16 lo pairs (one 128-byte line), float32-typed buffers [rows][ceil(C/32)*32 words].
 *   Weights, pre-split and pre-packed by the host (softgroup_b200/spconv/core.py:pack_weight_tc) with the same shift:
 *     N = Cout rounded up to 16, nkc = ceil(Cin/32);
 *     Wp fp16 [K][nkc][4][2][N][8]: element (k, kc, q, part, n, e) = part(W[k][32*kc + 8*q + e][n]) (0 outside Cin/Cout),
 *     part 0 = hi, part 1 = lo -- per (k, kc) a UMMA K-major / no-swizzle operand [B_hi | B_lo] of 16-byte chunks;
 *     passed as float* (two halves per word); sgb_spconv_tc_packed_floats = its length in words.
 *   sgb_spconv_forward_tc: d_in_pk packed input [Min][in_stride words] ALREADY activated (the consumer's
 *     BatchNorm(eval)+ReLU was applied by whoever packed it); d_map int32 [K][Mout] (-1 = no neighbour) or NULL for the
 *     identity (K == 1: 1x1 conv / nn.Linear). Outputs, either or both: d_out fp32 rows (strided: concat buffers);
 *     d_pk_out packed rows at channel offset pk_coff (multiple of 8) after y = relu?(x * pk_scale[c] + pk_shift[c]) per
 *     OUTPUT channel c -- the NEXT layer's BatchNorm(eval)+ReLU folded into this producer (an intermediate with a single
 *     consumer never exists in fp32); pk_fill != 0 also zeroes the upper half of a last chunk that N leaves half written.
 *   sgb_act_pack: the same packing for tensors without a producing convolution (C real channels, zeros up to Cfill).
 *   sgb_spconv_overflow: a packed value beyond the fp16 range (|y| > 65504, or NaN) raises a per-device flag instead of
 *     saturating silently; this call reads and clears it (blocking 4-byte read on `stream`).
 *   sgb_spconv_tc_plan: the tile configuration the launch uses for a problem size (pure function; sms <= 0: 148):
 *     out[0] = NT (columns per CTA), [1] = column parts, [2] = split-K cluster size, [3] = TMEM columns per CTA,
 *     [4] = pipeline stages, [5] = row tiles.
 * ------------------------------------------------------------------------------------------- */
long long sgb_spconv_tc_packed_floats(int K, int Cin, int Cout);
int sgb_spconv_lo_shift(void);
int sgb_spconv_overflow(int *h_flag, void *stream);
int sgb_spconv_tc_plan(int K, int Mout, int Cin, int Cout, int has_map, int sms, int *out);
int sgb_act_pack(const float *d_x, int x_stride, int x_off, const float *d_scale, const float *d_shift, int relu,
                 float *d_pk, int pk_stride, int pk_coff, int M, int C, int Cfill, void *stream);
int sgb_spconv_forward_tc(const float *d_in_pk, int in_stride, int Min, const int32_t *d_map, int K, int Mout,
                          const float *d_Wp, int Cin, int Cout, const float *d_residual, int res_stride, int res_off,
                          const float *d_bias, float *d_out, int out_stride, int out_off, float *d_pk_out, int pk_stride,
                          int pk_coff, const float *d_pk_scale, const float *d_pk_shift, int pk_relu, int pk_fill,
                          void *stream);
/* Two kernels implement sgb_spconv_forward_tc with bit-identical results (same products, same accumulation order):
 *   0 = register gather (spconv_tc.cu: rows -> registers -> tensor memory, column split and split-K clusters for the deep
 *       U-Net levels), 1 = persistent shared-memory ring (spconv_ss.cu: cp.async row gather into SWIZZLE_128B tiles, one
 *       CTA per SM walking the row tiles, accumulators double-buffered in tensor memory).
 * sgb_spconv_kernel_choice: which one sgb_spconv_forward_tc runs for a problem size (sms <= 0: 148).
 * sgb_spconv_forward_tc_ex: the same call with the kernel named explicitly (-1 = choose) -- parity tests and A/B tools. */
int sgb_spconv_kernel_choice(int K, int Mout, int Cin, int Cout, int sms);
int sgb_spconv_forward_tc_ex(const float *d_in_pk, int in_stride, int Min, const int32_t *d_map, int K, int Mout,
                             const float *d_Wp, int Cin, int Cout, const float *d_residual, int res_stride, int res_off,
                             const float *d_bias, float *d_out, int out_stride, int out_off, float *d_pk_out, int pk_stride,
                             int pk_coff, const float *d_pk_scale, const float *d_pk_shift, int pk_relu, int pk_fill, int kernel,
                             void *stream);

/* ---------------------------------------------------------------------------------------------
 * Per-class point selection in front of the ball query, all classes in one pass (grouping.cu) -- replaces the loop body
 * softgroup/model/softgroup.py:430-446 (`(scores[:, class_id] > score_thr).nonzero()`, the `min_npoint` skip, the gathers
 * of batch_idxs / coords_float / pt_offsets) for the classes h_classes[0..nc) (host array, nc <= 32):
 *   entries in class-major, ascending point order: d_pts[e] = point, d_seg[e] = rank(class) * B + batch_idxs[point],
 *   d_shifted[e] = coords[point] + offsets[point] (fp32), all with capacity N * nc; d_seg_offsets int32 [nc*B + 1] =
 *   first entry of every (class, batch item) segment; d_total int32 [1 + nc] = number of entries, then entries per class
 *   (0 for a class below min_npoint). d_scores = softmax scores [N][C]. No host synchronisation.
 * ------------------------------------------------------------------------------------------- */
size_t sgb_group_entries_workspace_bytes(int N, int nc, int B);
int sgb_group_entries(const float *d_scores, int N, int C, const int *h_classes, int nc, float score_thr, int min_npoint,
                      const int32_t *d_batch_idxs, int B, const float *d_coords, const float *d_offsets, int32_t *d_pts,
                      int32_t *d_seg, float *d_shifted, int32_t *d_seg_offsets, int32_t *d_total, void *d_ws, size_t ws_bytes,
                      void *stream);

/* y[i, c] = max(x[i, c]*scale[c] + shift[c], 0) (relu != 0) -- BatchNorm1d(eval)+ReLU over rows. */
int sgb_bn_relu(const float *d_x, int x_stride, const float *d_scale, const float *d_shift, int relu, float *d_y,
                int y_stride, int M, int C, void *stream);

/* out[i, :] = in[index[i], :]  (the "devoxelize" gather of softgroup.py:374); index int32 [N]. */
int sgb_gather_rows(const float *d_in, const int32_t *d_index, float *d_out, int N, int C, void *stream);

/* Host-side serialisation of instance masks into the reference's RLE wire format (softgroup/util/rle.py:5-19:
 * dict(length, counts='start len start len ...'), 1-based starts). h_ids: ascending point ids of all masks back to
 * back (int32), h_offs int64 [n_masks+1]. Writes the `counts` strings back to back into h_out and their byte ranges
 * into h_out_offs [n_masks+1]. Returns bytes written, or SGB_ERR_OVERFLOW if out_cap is too small
 * (12 bytes per run + 2 always suffices). No CUDA involved. */
long long sgb_rle_format_ids(const int32_t *h_ids, const long long *h_offs, int n_masks, char *h_out,
                             long long out_cap, long long *h_out_offs);

/* ---------------------------------------------------------------------------------------------
 * Instance masks, RLE, panoptic paste and evaluation intersections on the GPU (instances.cu) -- replace the dense
 * [nProposal, N] masks + numpy RLE of get_instances (softgroup/model/softgroup.py:537-604, softgroup/util/rle.py:5-19),
 * the numpy paste loop of panoptic_fusion (softgroup.py:606-639) and the per-pair np.count_nonzero of
 * ScanNetEval.assign_instances_for_scan (softgroup/evaluation/instance_eval.py:262-293).
 * Masks are bitmaps [rows][W], W = sgb_bitmap_words(N) = ceil((N+1)/32) 32-bit words (bit N always 0).
 *   sgb_inst_count:   npoint[p*nI+i] = #entries of proposal p with mask_scores[e,i] > thr (zeroed here).
 *   sgb_inst_scatter: bitmaps (zeroed here) of the kept instances; slot int32 [nI*nP], class-major (i*nP+p): bitmap row or -1.
 *   sgb_bitmap_set:   generic (row, point) -> bit; clear != 0 zeroes the bitmaps first.
 *   sgb_rle_count (blocking) / sgb_rle_fill: 1-based transition positions of every mask, ascending = `runs` of rle.py:15
 *     before `runs[1::2] -= runs[::2]`; inst_off int32 [n_inst+1]. sgb_rle_format_runs (host): the `counts` strings.
 *   sgb_bitmap_intersections: inter[r*nG+g], vert[r], void[r] from gslot int32 [N] (>= 0: gt column, -2: void, -1: other).
 *   sgb_panoptic_paste: instances visited in d_order; skipped when intersect/(size+1e-5) > skip_iou (float64), else
 *     pasted where nothing was pasted before: pan_cls[pt] = d_cls[row], pan_ids[pt] = 1, 2, ... in paste order.
 * ------------------------------------------------------------------------------------------- */
int sgb_inst_count(const int32_t *d_proposals_idx, const float *d_mask_scores, int ms_stride, int S, int nI, float thr,
                   int32_t *d_npoint, int nP, void *stream);
size_t sgb_bitmap_words(int N);
int sgb_inst_scatter(const int32_t *d_proposals_idx, const float *d_mask_scores, int ms_stride, int S, int nI, int nP, float thr,
                     const int32_t *d_slot, uint32_t *d_bitmaps, int n_inst, int N, void *stream);
int sgb_bitmap_set(const int32_t *d_row, const int32_t *d_pt, long long n, uint32_t *d_bitmaps, int n_rows, int N, int clear,
                   void *stream);
size_t sgb_rle_workspace_bytes(int n_inst, int N);
long long sgb_rle_count(const uint32_t *d_bitmaps, int n_inst, int N, void *d_ws, size_t ws_bytes, void *stream);
int sgb_rle_fill(const uint32_t *d_bitmaps, int n_inst, int N, int total_trans, int32_t *d_trans, int32_t *d_inst_off, void *d_ws,
                 size_t ws_bytes, void *stream);
long long sgb_rle_format_runs(const int32_t *h_trans, const int32_t *h_offs, int n_masks, char *h_out, long long out_cap,
                              long long *h_out_offs);
int sgb_bitmap_intersections(const uint32_t *d_bitmaps, int n_rows, int N, const int32_t *d_gslot, int nG, int32_t *d_inter,
                             int32_t *d_vert, int32_t *d_void, void *stream);
int sgb_panoptic_paste(const uint32_t *d_bitmaps, int N, const int32_t *d_order, const int32_t *d_cls, int n_inst, double skip_iou,
                       uint32_t *d_prev, uint32_t *d_pan_cls, uint32_t *d_pan_ids, void *stream);

/* Test-time transform of the reference dataloader, first step (softgroup/data/custom.py:87-107,162-164):
 * out[i, :] = xyz[i, :] @ m in float64 (xyz float32 [N,3] device, m float64 [3,3] row-major HOST, out float64 [N,3] device),
 * accumulated like the BLAS kernels numpy calls: fma(z, m2j, fma(y, m1j, x * m0j)). */
int sgb_affine3_f64(const float *d_xyz, const double *h_m9, double *d_out, int N, void *stream);


/* ---------------------------------------------------------------------------------------------
 * Sparse U-Net executor (unet.cu): one backbone pass = ONE call. The plan (array of sgb_unet_op) is compiled once per
 * model by softgroup_b200/model/unet_plan.py from the module tree of softgroup/model/blocks.py:44-143 with the fusion the
 * module path applies (consumer BatchNorm+ReLU in the producer's epilogue, packed activations, residual / concat writes in
 * epilogues); per scan the caller supplies row counts M[level], rulebooks per level (subm [27][M_l], down [8][M_{l+1}],
 * inverse [8][M_l]) and one pointer per plan buffer. Results are those of the same launches issued one by one.
 * ------------------------------------------------------------------------------------------- */
enum { SGB_UNET_CONV = 1, SGB_UNET_ACT_PACK = 2, SGB_UNET_COPY_COLS = 3, SGB_UNET_BN_RELU = 4 };
typedef struct sgb_unet_op {
  int32_t kind;                     /* SGB_UNET_* */
  int32_t level_in, level_out;      /* rows of the input / output: M[level] */
  int32_t map_kind;                 /* CONV: 0 identity (K = 1), 1 subm of level_out, 2 down level_in -> level_in + 1, 3 inverse -> level_out */
  int32_t K, Cin, Cout;             /* ACT_PACK: Cin = real channels, Cout = fill width; COPY / BN_RELU: Cin = channels */
  int32_t in_buf, in_stride, in_off;          /* fp32 source (ACT_PACK, COPY_COLS, BN_RELU) */
  int32_t pk_in_buf, pk_in_stride;            /* CONV: packed input, row stride in words */
  int32_t out_buf, out_stride, out_off;       /* fp32 destination (-1: none) */
  int32_t pk_out_buf, pk_out_stride, pk_out_coff, pk_fill, relu;  /* packed destination (-1: none) */
  int32_t res_buf, res_stride, res_off;       /* CONV: residual rows (-1: none) */
  const float *Wp, *bias, *scale, *shift;     /* device pointers: packed weights, bias, BatchNorm scale / shift of the op */
} sgb_unet_op;
int sgb_unet_run(const sgb_unet_op *ops, int n_ops, float *const *bufs, const int32_t *const *subm_maps,
                 const int32_t *const *down_maps, const int32_t *const *inv_maps, const int *M, int n_levels, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SGB200_H */
