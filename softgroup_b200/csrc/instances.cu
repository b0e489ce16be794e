// instances.cu -- the tail of the forward on the GPU: instance mask selection, run-length encoding, panoptic paste, and
// the prediction/ground-truth intersection matrix of the evaluation.
//
// Replaces (reference file:line):
//   get_instances           softgroup/model/softgroup.py:537-604  18 dense int32 [nProposal, N] masks -> D2H -> numpy RLE
//   rle_encode              softgroup/util/rle.py:5-19            np.where(mask[1:] != mask[:-1]) per instance on the host
//   panoptic_fusion         softgroup/model/softgroup.py:606-639  rle_decode + numpy paste loop per instance
//   assign_instances_for_scan  softgroup/evaluation/instance_eval.py:228-309  np.count_nonzero per (pred, gt) pair
// Masks are bitmaps [instance][W words], W = ceil((N + 1) / 32): bit N is always 0, so the closing transition of a mask
// that ends at point N-1 falls inside the bitmap like the reference's trailing pad (rle.py:14). A transition at point k
// (mask[k] != mask[k-1], mask[-1] = 0) is bit k of  t = x ^ ((x << 1) | carry)  -- one word-parallel pass finds every
// run boundary; the host only formats "start len" pairs from the (already sorted) transition positions.
#include <algorithm>

#include "common.cuh"

namespace sgb {

// npoint[p * nI + i] = #entries e of proposal p with mask_scores[e, i] > thr  (softgroup.py:553, 563-565 `mask_pred.sum(1)`).
// proposals_idx [S, 2] (proposal id, point); entries of one proposal are contiguous (cluster order), so a warp mostly
// holds one proposal: match_any groups + ballot popcounts keep the atomics at one per (warp, proposal, class).
__global__ void inst_count_kernel(const int32_t *__restrict__ pidx, const float *__restrict__ mask_scores, int ms_stride, int S,
                                  int nI, float thr, int32_t *__restrict__ npoint) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool ok = e < S;
  const int p = ok ? pidx[2 * (size_t)e] : -1;
  const unsigned peers = __match_any_sync(0xffffffffu, p);
  const bool leader = (__ffs(peers) - 1) == lane;
  for (int i = 0; i < nI; i++) {
    const bool on = ok && mask_scores[(size_t)e * ms_stride + i] > thr;
    const unsigned b = __ballot_sync(0xffffffffu, on) & peers;
    if (leader && p >= 0 && b) atomicAdd(&npoint[(size_t)p * nI + i], __popc(b));
  }
}

// slot[i * nP + p] (class-major, the order in which the reference appends instances) -> bitmap rows
__global__ void inst_scatter_kernel(const int32_t *__restrict__ pidx, const float *__restrict__ mask_scores, int ms_stride, int S,
                                    int nI, int nP, float thr, const int32_t *__restrict__ slot, uint32_t *__restrict__ bm, int W) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= S) return;
  const int p = pidx[2 * (size_t)e], pt = pidx[2 * (size_t)e + 1];
  for (int i = 0; i < nI; i++) {
    const int s = __ldg(&slot[(size_t)i * nP + p]);
    if (s >= 0 && mask_scores[(size_t)e * ms_stride + i] > thr) atomicOr(&bm[(size_t)s * W + (pt >> 5)], 1u << (pt & 31));
  }
}

// generic: set bit `pt` of row `row[e]` for every entry with row[e] >= 0
__global__ void bitmap_set_kernel(const int32_t *__restrict__ row, const int32_t *__restrict__ pt, long long n, uint32_t *__restrict__ bm,
                                  int W) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int r = row[e];
  if (r >= 0) atomicOr(&bm[(size_t)r * W + (pt[e] >> 5)], 1u << (pt[e] & 31));
}

__device__ __forceinline__ uint32_t transitions_of(const uint32_t *__restrict__ bm, int W, long long g) {
  const int w = (int)(g % W);
  const uint32_t x = bm[g];
  const uint32_t carry = (w > 0) ? (bm[g - 1] >> 31) : 0u;
  return x ^ ((x << 1) | carry);
}

__global__ void rle_count_kernel(const uint32_t *__restrict__ bm, int W, long long total, int32_t *__restrict__ cnt) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  cnt[g] = __popc(transitions_of(bm, W, g));
}

// trans[off[g] ...] = 1-based positions of the transitions of word g (ascending): exactly `runs` of rle.py:15 before the
// `runs[1::2] -= runs[::2]` step. inst_off[r] = off[r * W] (first transition of instance r).
__global__ void rle_fill_kernel(const uint32_t *__restrict__ bm, int W, long long total, const int32_t *__restrict__ off,
                                int32_t *__restrict__ trans, int32_t *__restrict__ inst_off, int n_inst, int grand_total) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g == 0) inst_off[n_inst] = grand_total;
  if (g >= total) return;
  const int w = (int)(g % W);
  int o = off[g];
  if (w == 0) inst_off[g / W] = o;
  uint32_t t = transitions_of(bm, W, g);
  while (t) {
    const int b = __ffs(t) - 1;
    t &= t - 1;
    trans[o++] = w * 32 + b + 1;
  }
}

// inter[r * nG + gslot[pt]] += 1 for every set bit (r, pt); vert[r] = popcount; void_inter[r] = set bits with gslot < 0
// marked void (gslot == -2). One warp per (row, 32-word strip).
__global__ void bitmap_hist_kernel(const uint32_t *__restrict__ bm, int W, int n_rows, int N, const int32_t *__restrict__ gslot,
                                   int nG, int32_t *__restrict__ inter, int32_t *__restrict__ vert, int32_t *__restrict__ void_inter) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)n_rows * W) return;
  const int r = (int)(g / W), w = (int)(g % W);
  uint32_t x = bm[g];
  if (!x) return;
  atomicAdd(&vert[r], __popc(x));
  int nvoid = 0;
  while (x) {
    const int b = __ffs(x) - 1;
    x &= x - 1;
    const int pt = w * 32 + b;
    if (pt >= N) continue;
    const int s = __ldg(&gslot[pt]);
    if (s >= 0) atomicAdd(&inter[(size_t)r * nG + s], 1);
    else if (s == -2) nvoid++;
  }
  if (nvoid) atomicAdd(&void_inter[r], nvoid);
}

// panoptic_fusion (softgroup.py:606-639): instances in descending confidence (order[] given by the caller), each is
// skipped when intersect / (size + 1e-5) > skip_iou (float64 like numpy), else pasted where nothing was pasted before.
// One CTA walks the instances in order (the loop is sequential by definition); words are spread over the threads.
constexpr int kPanThreads = 1024;
__global__ void __launch_bounds__(kPanThreads) panoptic_paste_kernel(const uint32_t *__restrict__ bm, int W, int N,
                                                                     const int32_t *__restrict__ order, const int32_t *__restrict__ cls,
                                                                     int n_inst, double skip_iou, uint32_t *__restrict__ prev /*[W] zeroed*/,
                                                                     uint32_t *__restrict__ pan_cls, uint32_t *__restrict__ pan_ids) {
  __shared__ int s_red[2][kPanThreads / 32];
  __shared__ int s_tot[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t next_id = 1;
  for (int k = 0; k < n_inst; k++) {
    const int r = order[k];
    const uint32_t *row = bm + (size_t)r * W;
    int inter = 0, size = 0;
    for (int w = tid; w < W; w += kPanThreads) {
      const uint32_t x = row[w];
      inter += __popc(x & prev[w]);
      size += __popc(x);
    }
    inter = warp_sum(inter);
    size = warp_sum(size);
    if (lane == 0) { s_red[0][warp] = inter; s_red[1][warp] = size; }
    __syncthreads();
    if (warp == 0) {
      int a = (lane < kPanThreads / 32) ? s_red[0][lane] : 0, b = (lane < kPanThreads / 32) ? s_red[1][lane] : 0;
      a = warp_sum(a);
      b = warp_sum(b);
      if (lane == 0) { s_tot[0] = a; s_tot[1] = b; }
    }
    __syncthreads();
    const bool skip = (double)s_tot[0] / ((double)s_tot[1] + 1e-5) > skip_iou;
    if (!skip) {
      const uint32_t c = (uint32_t)cls[r];
      for (int w = tid; w < W; w += kPanThreads) {
        uint32_t paste = row[w] & ~prev[w];
        if (paste) {
          prev[w] |= paste;
          while (paste) {
            const int b = __ffs(paste) - 1;
            paste &= paste - 1;
            const int pt = w * 32 + b;
            if (pt < N) { pan_cls[pt] = c; pan_ids[pt] = next_id; }
          }
        }
      }
      next_id++;
    }
    __syncthreads();
  }
}

}  // namespace sgb

using namespace sgb;

extern "C" {

int sgb_inst_count(const int32_t *d_proposals_idx, const float *d_mask_scores, int ms_stride, int S, int nI, float thr,
                   int32_t *d_npoint, int nP, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SGB_REQUIRE(S >= 0 && nI >= 0 && nP >= 0 && d_npoint, SGB_ERR_ARG, "inst_count arguments");
  SGB_CUDA_CHECK(cudaMemsetAsync(d_npoint, 0, (size_t)nP * nI * 4, st));
  if (S == 0 || nI == 0) return SGB_OK;
  SGB_REQUIRE(d_proposals_idx && d_mask_scores && ms_stride >= nI, SGB_ERR_ARG, "inst_count pointers");
  inst_count_kernel<<<div_up(S, 256), 256, 0, st>>>(d_proposals_idx, d_mask_scores, ms_stride, S, nI, thr, d_npoint);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}

size_t sgb_bitmap_words(int N) { return (size_t)(N + 1 + 31) / 32; }

int sgb_inst_scatter(const int32_t *d_proposals_idx, const float *d_mask_scores, int ms_stride, int S, int nI, int nP, float thr,
                     const int32_t *d_slot, uint32_t *d_bitmaps, int n_inst, int N, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int W = (int)sgb_bitmap_words(N);
  SGB_REQUIRE(S >= 0 && n_inst >= 0 && N >= 0, SGB_ERR_ARG, "inst_scatter arguments");
  if (n_inst == 0) return SGB_OK;
  SGB_REQUIRE(d_bitmaps, SGB_ERR_ARG, "null bitmaps");
  SGB_CUDA_CHECK(cudaMemsetAsync(d_bitmaps, 0, (size_t)n_inst * W * 4, st));
  if (S == 0 || nI == 0) return SGB_OK;
  SGB_REQUIRE(d_proposals_idx && d_mask_scores && d_slot, SGB_ERR_ARG, "inst_scatter pointers");
  inst_scatter_kernel<<<div_up(S, 256), 256, 0, st>>>(d_proposals_idx, d_mask_scores, ms_stride, S, nI, nP, thr, d_slot, d_bitmaps, W);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}

int sgb_bitmap_set(const int32_t *d_row, const int32_t *d_pt, long long n, uint32_t *d_bitmaps, int n_rows, int N, int clear,
                   void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int W = (int)sgb_bitmap_words(N);
  SGB_REQUIRE(n >= 0 && n_rows >= 0 && (n_rows == 0 || d_bitmaps), SGB_ERR_ARG, "bitmap_set arguments");
  if (clear && n_rows) SGB_CUDA_CHECK(cudaMemsetAsync(d_bitmaps, 0, (size_t)n_rows * W * 4, st));
  if (n == 0) return SGB_OK;
  SGB_REQUIRE(d_row && d_pt, SGB_ERR_ARG, "bitmap_set pointers");
  bitmap_set_kernel<<<div_up(n, 256), 256, 0, st>>>(d_row, d_pt, n, d_bitmaps, W);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}

size_t sgb_rle_workspace_bytes(int n_inst, int N) {
  const size_t tot = (size_t)std::max(n_inst, 0) * sgb_bitmap_words(N) + 1;
  return align_up(tot * 4) + align_up(scan_temp_elems(tot) * 4) + align_up(64) + 1024;
}

// count (blocking): returns the number of transitions of all masks (2 per run), or < 0.
long long sgb_rle_count(const uint32_t *d_bitmaps, int n_inst, int N, void *d_ws, size_t ws_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_inst == 0) return 0;
  const int W = (int)sgb_bitmap_words(N);
  const long long total = (long long)n_inst * W;
  SGB_REQUIRE(total < (1ll << 31), SGB_ERR_RANGE, "rle: n_inst * words must stay below 2^31");
  Arena a(d_ws, ws_bytes);
  int32_t *cnt = a.take<int32_t>((size_t)total + 1);
  int32_t *tmp = a.take<int32_t>(scan_temp_elems((size_t)total + 1));
  int32_t *tot_d = a.take<int32_t>(16);
  SGB_REQUIRE(d_bitmaps && cnt && tmp && tot_d, SGB_ERR_WORKSPACE, "rle workspace too small");
  rle_count_kernel<<<div_up(total, 256), 256, 0, st>>>(d_bitmaps, W, total, cnt);
  SGB_LAUNCH_CHECK();
  int rc = exclusive_scan_i32(cnt, cnt, (size_t)total, tot_d, tmp, st);
  if (rc) return rc;
  int h = 0;
  SGB_CUDA_CHECK(cudaMemcpyAsync(&h, tot_d, 4, cudaMemcpyDeviceToHost, st));
  SGB_CUDA_CHECK(cudaStreamSynchronize(st));
  return h;
}

// fill: after sgb_rle_count on the same workspace. d_trans int32 [total], d_inst_off int32 [n_inst + 1].
int sgb_rle_fill(const uint32_t *d_bitmaps, int n_inst, int N, int total_trans, int32_t *d_trans, int32_t *d_inst_off, void *d_ws,
                 size_t ws_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SGB_REQUIRE(d_inst_off, SGB_ERR_ARG, "null inst_off");
  if (n_inst == 0) { SGB_CUDA_CHECK(cudaMemsetAsync(d_inst_off, 0, 4, st)); return SGB_OK; }
  const int W = (int)sgb_bitmap_words(N);
  const long long total = (long long)n_inst * W;
  Arena a(d_ws, ws_bytes);
  int32_t *off = a.take<int32_t>((size_t)total + 1);
  SGB_REQUIRE(d_bitmaps && off && (total_trans == 0 || d_trans), SGB_ERR_WORKSPACE, "rle workspace too small");
  rle_fill_kernel<<<div_up(total, 256), 256, 0, st>>>(d_bitmaps, W, total, off, d_trans, d_inst_off, n_inst, total_trans);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}

// Host-side formatter: transitions (1-based positions, ascending per mask; offsets int32 [n+1]) -> "start len ..." strings.
long long sgb_rle_format_runs(const int32_t *h_trans, const int32_t *h_offs, int n_masks, char *h_out, long long out_cap,
                              long long *h_out_offs) {
  auto put = [](char *p, long long v) -> char * {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
  };
  char *p = h_out;
  char *end = h_out + out_cap;
  for (int m = 0; m < n_masks; m++) {
    h_out_offs[m] = p - h_out;
    const int a = h_offs[m], b = h_offs[m + 1];
    if ((b - a) & 1) { set_error("sgb_rle_format_runs: odd number of transitions in mask %d", m); return SGB_ERR_ARG; }
    for (int i = a; i < b; i += 2) {
      if (end - p < 48) { set_error("sgb_rle_format_runs: output buffer too small"); return SGB_ERR_OVERFLOW; }
      if (i > a) *p++ = ' ';
      p = put(p, (long long)h_trans[i]);
      *p++ = ' ';
      p = put(p, (long long)h_trans[i + 1] - h_trans[i]);
    }
  }
  h_out_offs[n_masks] = p - h_out;
  return p - h_out;
}

// Prediction x ground-truth intersections (instance_eval.py:262-293). d_gslot int32 [N]: column of the point's gt
// instance (>= 0), -2 = void label (counted in d_void), -1 = neither. Outputs are zeroed here.
int sgb_bitmap_intersections(const uint32_t *d_bitmaps, int n_rows, int N, const int32_t *d_gslot, int nG, int32_t *d_inter,
                             int32_t *d_vert, int32_t *d_void, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SGB_REQUIRE(n_rows >= 0 && N >= 0 && nG >= 0, SGB_ERR_ARG, "bitmap_intersections arguments");
  if (n_rows == 0) return SGB_OK;
  SGB_REQUIRE(d_bitmaps && d_gslot && d_vert && d_void && (nG == 0 || d_inter), SGB_ERR_ARG, "bitmap_intersections pointers");
  if (nG) SGB_CUDA_CHECK(cudaMemsetAsync(d_inter, 0, (size_t)n_rows * nG * 4, st));
  SGB_CUDA_CHECK(cudaMemsetAsync(d_vert, 0, (size_t)n_rows * 4, st));
  SGB_CUDA_CHECK(cudaMemsetAsync(d_void, 0, (size_t)n_rows * 4, st));
  const int W = (int)sgb_bitmap_words(N);
  bitmap_hist_kernel<<<div_up((long long)n_rows * W, 256), 256, 0, st>>>(d_bitmaps, W, n_rows, N, d_gslot, nG, d_inter, d_vert, d_void);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}

// panoptic_fusion on bitmaps. d_order int32 [n_inst]: instance rows in descending confidence (the caller sorts: numpy
// argsort()[::-1] order must be reproduced by the caller); d_cls int32 [n_rows]: class value to paste (label_id + offset).
// d_pan_cls must hold the semantic predictions on entry (uint32 [N]), d_pan_ids zeros; d_prev uint32 [W] scratch.
int sgb_panoptic_paste(const uint32_t *d_bitmaps, int N, const int32_t *d_order, const int32_t *d_cls, int n_inst, double skip_iou,
                       uint32_t *d_prev, uint32_t *d_pan_cls, uint32_t *d_pan_ids, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SGB_REQUIRE(n_inst >= 0 && N >= 0, SGB_ERR_ARG, "panoptic_paste arguments");
  if (n_inst == 0 || N == 0) return SGB_OK;
  SGB_REQUIRE(d_bitmaps && d_order && d_cls && d_prev && d_pan_cls && d_pan_ids, SGB_ERR_ARG, "panoptic_paste pointers");
  const int W = (int)sgb_bitmap_words(N);
  SGB_CUDA_CHECK(cudaMemsetAsync(d_prev, 0, (size_t)W * 4, st));
  panoptic_paste_kernel<<<1, kPanThreads, 0, st>>>(d_bitmaps, W, N, d_order, d_cls, n_inst, skip_iou, d_prev, d_pan_cls, d_pan_ids);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}
}
