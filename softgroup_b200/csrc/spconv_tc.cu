// spconv_tc.cu -- sparse convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// One CTA owns a tile of 128 output rows x NT <= 128 output columns (and, with split-K, a contiguous share of the
// (kernel offset, 32-channel slice) iterations). For every kernel offset with at least one active pair in the tile
// and every 32-channel slice of Cin:
//   * the producer threads gather one input row slice each. Inputs are either raw fp32 rows (eval-BatchNorm + ReLU
//     folded in here, then split) or rows already activated and split by sgb_act_split. A value x is carried as
//     two fp16 numbers hi = fp16(x), lo = fp16(x - hi); both go straight into TENSOR MEMORY with tcgen05.st
//     (A operand from TMEM: lane = row, 16 columns of hi pairs + 16 columns of lo pairs per stage) -- the
//     shared-memory pipe only carries the weights;
//   * the pre-packed weight slice (same split, core-matrix order, done once on the host side) is streamed by one
//     elected lane with TMA bulk copies into a ring of stages; the weights of the whole convolution are pulled
//     into L2 with cp.async.bulk.prefetch at kernel start because every CTA walks them in the same order (so they
//     would otherwise always be cold);
//   * one thread issues the error-compensated products as tcgen05.mma.cta_group::1.kind::f16 (M=128, K=16)
//         D[:, 0:2nt]  += A_hi * [B_hi | B_lo]      D[:, nt:2nt] += A_lo * B_hi
//     accumulating fp32 in TMEM, then tcgen05.commit's the stages back to their producers through mbarriers.
// The epilogue reads the accumulator with tcgen05.ld, sums the two column blocks, adds bias / residual and writes
// strided rows (the U-Net concat buffer). With split-K (deep U-Net levels: a handful of row tiles, megabytes of
// weights) the CTAs of one thread-block cluster hold partial tiles; ranks > 0 park theirs in their own shared memory
// and rank 0 sums them in rank order over distributed shared memory -- deterministic, no workspace, no atomics.
// Dropping lo*lo leaves a relative error of ~2^-22 per product, i.e. fp32-grade (the north star's 1e-4 over ~40
// sequential convolutions rules out plain TF32/fp16).
#include <algorithm>

#include <cuda_fp16.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace sgb {

constexpr int TC_ROWS = 128;
constexpr int TC_KC = 32;  // channels per stage
constexpr int TC_THREADS = 320;  // 8 producer warps + 1 MMA warp + 1 weight-loader warp
// The remainder lo = x - fp16(x) is carried as fp16(lo * 2^kLoShift); the correction columns are scaled back by
// 2^-kLoShift when they are added to the main product. 0 = validated behaviour (lo is a subnormal fp16 for |x| < ~0.1:
// 2^-25 absolute error); 11 removes the subnormal range and is the round-2 candidate (sgb_spconv_tc_lo_shift() tells
// the host packer which one the library was built with).
constexpr int kLoShift = 0;
constexpr float kLoScale = (float)(1 << kLoShift), kLoInv = 1.0f / kLoScale;
constexpr int BAR_FULL = 0, BAR_FREE = 3, BAR_BFULL = 9;  // per pair stage (<= 3); bars[8] = accumulator done

struct TcArgs {
  const float *in; int in_stride, in_off;
  const int32_t *map; int K, Mout;
  const float *Wp;  // packed fp16 [K][nkc][4 chunks][2 (hi,lo)][N][8 halves]
  int Cin, N, Cout;        // N = Cout rounded up to 16
  int NT;                  // columns per CTA (multiple of 16); gridDim.y = ceil(N / NT)
  const float *in_scale, *in_shift;
  const float *residual; int res_stride, res_off;
  const float *bias;
  float *out; int out_stride, out_off;
  int nstages, nbstages, tmem_cols, tmem_acols;  // A ring depth (TMEM), weight ring depth (smem), TMEM columns, first A column
  int ksplit;      // CTAs per cluster along z sharing one tile's iterations (1 = no split)
  int prefetch;    // pull the packed weights into L2 at kernel start
  int in_packed;   // input rows are already activated + split: per 32-channel chunk [16 words hi pairs | 16 words lo pairs]
  long long *dbg;  // optional timeline buffer (test hook)
  int skip;        // test hook (timing decomposition only, results are garbage): 1 no tcgen05.st, 2 no MMA, 4 no gather, 8 no weight copy
  int gather_off;  // GATHER = 1: byte offset of the per-warp gather tiles (8 warps x 2 x 4 KB) in dynamic shared memory
};

// Warp roles: warps 0-3 = producer group 0, warps 4-7 = producer group 1 (one output row per thread; group g fills slot g
// of every iteration pair), warp 8 = MMA issuer (warp-uniform, one elected lane), warp 9 = weight loader (TMA bulk copies).
// Barriers per pair stage: full (8 producer-warp arrivals), bfull (weights: expect_tx), free (tcgen05.commit);
// bars[8]: accumulator complete.
// GATHER = 0: every producer lane fetches its own row slice (validated path). GATHER = 1 (round-2 candidate, packed
// inputs only): lanes cooperate on rows through shared memory, see the producer branch.
template <int GATHER>
__global__ void __launch_bounds__(TC_THREADS, 2) spconv_tc_kernel(TcArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[12];  // [0..2] A full, [3..5] pair free, [8] done, [9..11] weights full
  __shared__ uint32_t s_tmem;
  __shared__ long long s_dbg[GATHER ? 8 : 64 * 8];  // in-kernel timeline (test hook): shared memory so that the stamps do not add global stores to the fences
  __shared__ unsigned int s_mask;
  __shared__ int s_list[32];
  __shared__ int s_nact;
  __shared__ __align__(16) float s_scale[GATHER ? 4 : 512], s_shift[GATHER ? 4 : 512];  // raw inputs never use GATHER = 1

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);  // provably warp-uniform copy for the role dispatch
  const bool mark_on = p.dbg && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0;
  if (mark_on) p.dbg[64 * 8 + 0] = clock64();
  if constexpr (GATHER == 0) {
    if (p.dbg && (p.skip & 32))
      for (int i = threadIdx.x; i < 64 * 8; i += TC_THREADS) s_dbg[i] = 0;
  }
  const bool producer = warp < 8;
  const int grp = warp >> 2;          // producer group (0/1)
  const int r = tid & (TC_ROWS - 1);  // row of this producer thread
  const int row0 = blockIdx.x * TC_ROWS;
  const int my_row = row0 + r;
  const bool row_ok = producer && my_row < p.Mout;
  const bool has_act = p.in_scale != nullptr;
  const bool vec_ok = ((p.in_stride & 3) == 0) && ((p.in_off & 3) == 0) && ((((uintptr_t)p.in) & 15) == 0);
  const int N = p.N, NT = p.NT;
  const int n0 = blockIdx.y * NT;
  const int nt = min(NT, N - n0);                     // columns of this CTA (multiple of 16)
  const uint32_t b_bytes = (uint32_t)NT * TC_KC * 2;  // NT * 64 B (fp16)
  const uint32_t bstage_bytes = 2 * b_bytes;    // weight ring stage: hi + lo
  const int NS = p.nstages;   // A stages in TMEM (32 columns each) == weight stages in shared memory; even
  unsigned char *bring = smem;
  int32_t *map_s = reinterpret_cast<int32_t *>(bring + (size_t)NS * bstage_bytes);  // [K][128] (only when p.map)

  if (tid == 0) {
    for (int i = 0; i < NS / 2; i++) {
      mbar_init(smem_u32(&bars[BAR_FULL + i]), 8);  // one arrival per producer warp
      mbar_init(smem_u32(&bars[BAR_FREE + i]), 1);
      mbar_init(smem_u32(&bars[BAR_BFULL + i]), 1);
    }
    mbar_init(smem_u32(&bars[8]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_mask = 0u;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (mark_on) p.dbg[64 * 8 + 1] = clock64();

  if (GATHER == 0 && has_act) {
    for (int c = tid; c < 512; c += TC_THREADS) {
      s_scale[c] = (c < p.Cin) ? __ldg(&p.in_scale[c]) : 0.f;
      s_shift[c] = (c < p.Cin) ? __ldg(&p.in_shift[c]) : 0.f;
    }
  }
  // ---- rulebook slice of this tile -> shared memory; which kernel offsets have any active pair ------------
  if (p.map) {
    unsigned int flags = 0u;
    if (producer) {
      // all of this thread's rulebook entries are requested before the first one is used (one memory latency, not K/2)
      for (int ob = grp; ob < p.K; ob += 28) {
        int srcv[14];
#pragma unroll
        for (int j = 0; j < 14; j++) {
          const int o = ob + 2 * j;
          srcv[j] = (o < p.K && row_ok) ? __ldg(&p.map[(size_t)o * p.Mout + my_row]) : -1;
        }
#pragma unroll
        for (int j = 0; j < 14; j++) {
          const int o = ob + 2 * j;
          if (o < p.K) {
            map_s[o * TC_ROWS + r] = srcv[j];
            if (srcv[j] >= 0) flags |= 1u << o;
          }
        }
      }
    }
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0 && flags) atomicOr(&s_mask, flags);
  }
  __syncthreads();
  if (mark_on) p.dbg[64 * 8 + 2] = clock64();
  if (tid < 32) {  // compact the active offsets in ascending order
    if (p.map) {
      const unsigned int m = s_mask;
      if (tid < p.K && (m >> tid & 1u)) s_list[__popc(m & ((1u << tid) - 1u))] = tid;
      if (tid == 0) s_nact = __popc(m);
    } else if (tid == 0) {
      s_list[0] = 0;
      s_nact = 1;
    }
  }
  __syncthreads();
  if (mark_on) p.dbg[64 * 8 + 3] = clock64();
  const int nkc = (p.Cin + TC_KC - 1) / TC_KC;
  // split-K: the cluster's CTAs take contiguous shares of the tile's (offset, slice) iterations
  const int S = p.ksplit, z = blockIdx.z;
  const int total_all = s_nact * nkc;
  const int i_beg = (int)((long long)total_all * z / S), i_end = (int)((long long)total_all * (z + 1) / S);
  const int total = i_end - i_beg;
  const int a0 = i_beg / nkc, k0 = i_beg - a0 * nkc;  // first iteration of this CTA

  // Pipeline unit = a PAIR of consecutive iterations (2P, 2P+1): producer group g fills A stage (ps, g), the loader
  // brings both weight slices under one transaction barrier, and the MMA thread pays its fixed costs (two barrier
  // waits, one commit) once per pair and issues up to 8 back-to-back MMAs, which is what lets them pipeline in the
  // tensor core. Per pair stage ps: full[ps] (8 producer-warp arrivals), bfull[ps] (weights landed),
  // free[ps] (tcgen05.commit: both the A stages and the weight stages of the pair can be overwritten).
  const int npairs = (total + 1) >> 1;
  const int NP = NS >> 1;
  if (producer && p.in_packed) {
   if constexpr (GATHER == 1) {
    // ---- cooperative gather (packed inputs): the 8 lanes of a quarter-warp fetch one whole 128-byte row slice with
    //      cp.async (16 B each), so a warp-wide copy touches 4 lines instead of 32 -- the per-lane gather of GATHER = 0 is
    //      bound by L1 tag look-ups. Rows land in a per-warp shared-memory tile with the 16-byte chunk index XOR-ed
    //      by (row & 7): both the quarter-warp writes and the lane-owns-a-row reads (LDS.128) are bank-conflict free.
    //      Two tiles per warp: the copy of own iteration +1 is in flight while the current one moves to TMEM.
    const uint32_t tile_u = smem_u32(smem + p.gather_off) + (uint32_t)warp * 8192u;
    const int wrow0 = (warp & 3) * 32;  // first tile row of this warp (TMEM lanes wrow0 .. wrow0 + 31)
    int ia = a0, ikc = k0 + grp;
    while (ikc >= nkc) { ikc -= nkc; ia++; }
    int nload = grp;
    auto issue = [&](int b) {  // always commits one group (possibly empty) so that wait_group counts stay aligned
      if (nload < total) {
        const int o = s_list[ia];
        const int i = lane & 7;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int R = 4 * j + (lane >> 3);
          const int grow = row0 + wrow0 + R;
          const int src = p.map ? map_s[o * TC_ROWS + wrow0 + R] : (grow < p.Mout ? grow : -1);
          const uint32_t dst = tile_u + (uint32_t)b * 4096u + (uint32_t)(R * 128 + ((i ^ (R & 7)) << 4));
          const float *g = p.in + (size_t)max(src, 0) * p.in_stride + p.in_off + ikc * TC_KC + i * 4;
          const int nbytes = (src >= 0 && !(p.skip & 4)) ? 16 : 0;  // 0 source bytes = zero fill (absent neighbour)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(g), "r"(nbytes) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      nload += 2;
      ikc += 2;
      while (ikc >= nkc) { ikc -= nkc; ia++; }
    };
    int ps = 0, u = 0, b = 0;
    issue(0);
    issue(1);
    for (int P = 0; P < npairs; P++) {
      const bool work = 2 * P + grp < total;
      if (u >= 1) mbar_wait(smem_u32(&bars[BAR_FREE + ps]), (uint32_t)((u - 1) & 1));
      if (work) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");  // everything but the newest group: tile b has landed
        __syncwarp();
        uint32_t w[32];
        const uint32_t rbase = tile_u + (uint32_t)b * 4096u + (uint32_t)lane * 128u;
#pragma unroll
        for (int c = 0; c < 8; c++)
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(w[4 * c + 0]), "=r"(w[4 * c + 1]), "=r"(w[4 * c + 2]), "=r"(w[4 * c + 3])
                       : "r"(rbase + (uint32_t)((c ^ (lane & 7)) << 4)));
        __syncwarp();  // every lane has read its row before the tile is refilled
        const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(p.tmem_acols + (2 * ps + grp) * 32);
        if (!(p.skip & 1)) tmem_st32(ta, w);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars[BAR_FULL + ps]));
      if (work) issue(b);  // own iteration +2 into the tile just drained (after the fence, like GATHER = 0)
      b ^= 1;
      if (++ps == NP) { ps = 0; u++; }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
   } else {
    // ---- packed input (activated + split once by sgb_act_split): the gather is pure data movement, so the registers
    //      freed by the missing transform hold TWO future iterations of this thread's row (4 iterations ahead of the
    //      MMA warp counting both groups) -- the L2 latency of the gather is covered without shared memory.
    uint32_t va[32], vb[32];
    int ia = a0, ikc = k0 + grp;  // issue cursor: (offset list position, channel slice) of the next own iteration to load
    while (ikc >= nkc) { ikc -= nkc; ia++; }
    int nload = grp;        // CTA-local iteration index of the next load
    // One gathered row slice = one 128-byte line, every lane a different line: the loads are tag-bound in L1, so the
    // slice is fetched with four 256-bit loads (LDG.E.256) instead of eight 128-bit ones.
    auto load = [&](uint32_t (&buf)[32]) {
      if (nload < total) {
        const int o = s_list[ia];
        const int src = p.map ? map_s[o * TC_ROWS + r] : (row_ok ? my_row : -1);
        if (src >= 0 && !(p.skip & 4)) {
          const float *rp = p.in + (size_t)src * p.in_stride + p.in_off + ikc * TC_KC;
#pragma unroll
          for (int q = 0; q < 4; q++)
            asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(buf[8 * q + 0]), "=r"(buf[8 * q + 1]), "=r"(buf[8 * q + 2]), "=r"(buf[8 * q + 3]),
                           "=r"(buf[8 * q + 4]), "=r"(buf[8 * q + 5]), "=r"(buf[8 * q + 6]), "=r"(buf[8 * q + 7])
                         : "l"(rp + 8 * q));
        } else {
#pragma unroll
          for (int q = 0; q < 32; q++) buf[q] = 0u;
        }
      }
      nload += 2;
      ikc += 2;
      while (ikc >= nkc) { ikc -= nkc; ia++; }
    };
    int ps = 0, u = 0, P = 0;
    auto consume = [&](uint32_t (&buf)[32]) {
      const bool work = 2 * P + grp < total;  // an odd tail leaves group 1 without a slice: it only arrives
      const bool dbg_on = p.dbg && (p.skip & 32) && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0 && r == 0 && 2 * P + grp < 64;
      const int di = 2 * P + grp;
      if (dbg_on) s_dbg[di * 8 + 0] = clock64();
      if (u >= 1) mbar_wait(smem_u32(&bars[BAR_FREE + ps]), (uint32_t)((u - 1) & 1));
      if (dbg_on) s_dbg[di * 8 + 1] = clock64();
      if (work) {
        const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(p.tmem_acols + (2 * ps + grp) * 32);
        if (!(p.skip & 1)) tmem_st32(ta, buf);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      if (!(p.skip & 16)) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (dbg_on) s_dbg[di * 8 + 2] = clock64();
      // one arrival per WARP (every lane has completed and fenced its own store): 256 per-thread arrivals on one
      // mbarrier serialise in the shared-memory atomic unit and were the floor of the whole pipeline
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars[BAR_FULL + ps]));
      // Refill this buffer (own iteration +2) only AFTER the arrive: tcgen05.wait::st / tcgen05.fence compile to
      // FENCE.VIEW.ASYNC, which also waits for every outstanding global load of the thread -- issued before the
      // fence, the gather's L2 latency was paid in full on every iteration instead of overlapping the other buffer.
      if (work) load(buf);
      if (dbg_on) s_dbg[di * 8 + 3] = clock64();
      P++;
      if (++ps == NP) { ps = 0; u++; }
    };
    load(va);
    load(vb);
    for (int j = 0; j < npairs; j += 2) {
      consume(va);
      if (j + 1 < npairs) consume(vb);
    }
   }  // GATHER == 0
  } else if (producer) {
    float4 v[8];
    int vsrc = -1;
    // gather of iteration i into registers (whole 32-channel slice of this thread's row)
    auto load_iter = [&](int a_idx, int kc) {
      const int o = s_list[a_idx];
      const int c0 = kc * TC_KC;
      const int kvalid = min(TC_KC, p.Cin - c0);
      vsrc = p.map ? map_s[o * TC_ROWS + r] : (row_ok ? my_row : -1);
      const float *rp = (vsrc >= 0) ? p.in + (size_t)vsrc * p.in_stride + p.in_off + c0 : nullptr;
      const bool fast = (vsrc >= 0) && vec_ok && kvalid == TC_KC;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fast) {
          v[q] = __ldg(reinterpret_cast<const float4 *>(rp) + q);
        } else if (vsrc >= 0 && 4 * q < kvalid) {
          float t[4];
#pragma unroll
          for (int e = 0; e < 4; e++) t[e] = (4 * q + e < kvalid) ? __ldg(rp + 4 * q + e) : 0.f;
          v[q] = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
    };
    // own iterations i = grp, grp+2, ... = slot grp of pair P; (a_idx, kc) advanced incrementally
    int i = grp;
    int kc = k0 + grp, a_idx = a0;
    while (kc >= nkc) { kc -= nkc; a_idx++; }
    int ps = 0, u = 0;
    if (i < total) load_iter(a_idx, kc);
    for (int P = 0; P < npairs; P++, i += 2) {
      const bool work = i < total;
      if (u >= 1) mbar_wait(smem_u32(&bars[BAR_FREE + ps]), (uint32_t)((u - 1) & 1));
      if (work) {
        const int c0 = kc * TC_KC;
        const int kvalid = min(TC_KC, p.Cin - c0);
        // ---- A: registers -> (BN+ReLU) -> hi / lo -> tensor memory (lane = row, column = channel) -----------
        const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(p.tmem_acols + (2 * ps + grp) * 32);
        // x = hi + lo with hi = fp16(x), lo = fp16(x - hi): |x - hi - lo| <= 2^-22 |x| (or 2^-25 absolute when lo is
        // subnormal); two halves per 32-bit TMEM column (channel 2c in the low half). 16 columns hi + 16 columns lo.
        uint32_t hv[16], lv[16];
#pragma unroll
        for (int q = 0; q < 8; q++) {
          float4 x = v[q];
          if (has_act && vsrc >= 0) {
            const float4 sc = *reinterpret_cast<const float4 *>(&s_scale[c0 + 4 * q]);
            const float4 sh = *reinterpret_cast<const float4 *>(&s_shift[c0 + 4 * q]);
            x.x = fmaxf(fmaf(x.x, sc.x, sh.x), 0.f);
            x.y = fmaxf(fmaf(x.y, sc.y, sh.y), 0.f);
            x.z = fmaxf(fmaf(x.z, sc.z, sh.z), 0.f);
            x.w = fmaxf(fmaf(x.w, sc.w, sh.w), 0.f);
            if (kvalid < TC_KC) {  // channels past Cin must stay exactly 0
              if (4 * q + 0 >= kvalid) x.x = 0.f;
              if (4 * q + 1 >= kvalid) x.y = 0.f;
              if (4 * q + 2 >= kvalid) x.z = 0.f;
              if (4 * q + 3 >= kvalid) x.w = 0.f;
            }
          }
          const __half2 h01 = f2h2_sat(x.x, x.y), h23 = f2h2_sat(x.z, x.w);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const __half2 l01 = f2h2_sat((x.x - f01.x) * kLoScale, (x.y - f01.y) * kLoScale),
                        l23 = f2h2_sat((x.z - f23.x) * kLoScale, (x.w - f23.y) * kLoScale);
          hv[2 * q] = *reinterpret_cast<const uint32_t *>(&h01);
          hv[2 * q + 1] = *reinterpret_cast<const uint32_t *>(&h23);
          lv[2 * q] = *reinterpret_cast<const uint32_t *>(&l01);
          lv[2 * q + 1] = *reinterpret_cast<const uint32_t *>(&l23);
        }
        tmem_st16(ta, hv);
        tmem_st16(ta + 16u, lv);
        kc += 2;
        while (kc >= nkc) { kc -= nkc; a_idx++; }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars[BAR_FULL + ps]));
      if (work && i + 2 < total) load_iter(a_idx, kc);  // after the fence (it would wait for these loads), see above
      if (++ps == NP) { ps = 0; u++; }
    }
  } else if (warp == 9) {
    // ---- weight loader: one elected lane streams the packed slices with TMA bulk copies (cp.async.bulk, async proxy:
    //      no generic->async fence needed); completion is signalled on the pair's mbarrier by complete_tx. Global
    //      layout [chunk q][hi|lo][N][16 B] == shared layout [q][hi nt | lo nt][16 B] when the CTA owns all N columns,
    //      so a whole slice is ONE bulk copy; with a column split it is one copy per (q, part), one lane each.
    if (p.prefetch) {
      // every CTA streams the weights in the same order, so without this they are cold for everybody at once
      const long long wbytes = (long long)p.K * nkc * N * 128;
      const int lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
      const int ncta = min((int)(gridDim.x * gridDim.y * gridDim.z), 256);
      if (lin < ncta) {
        const char *wb = reinterpret_cast<const char *>(p.Wp);
        for (long long off = ((long long)lin * 32 + lane) * 4096; off < wbytes; off += (long long)ncta * 32 * 4096) {
          const uint32_t sz = (uint32_t)min(4096ll, wbytes - off);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(wb + off), "r"(sz) : "memory");
        }
      }
    }
    int ps = 0, ub = 0, kc = k0, a_idx = a0;
    for (int P = 0; P < npairs; P++) {
      if (ub >= 1) mbar_wait(smem_u32(&bars[BAR_FREE + ps]), (uint32_t)((ub - 1) & 1));
      const int nit = min(2, total - 2 * P);
      const uint32_t bar = smem_u32(&bars[BAR_BFULL + ps]);
      // slot 0 / slot 1 of the pair: offset, slice, 16-channel k-steps
      const int o0 = s_list[a_idx], kc0 = kc;
      int kc1 = kc + 1, a1 = a_idx;
      if (kc1 == nkc) { kc1 = 0; a1++; }
      const int o1 = (nit > 1) ? s_list[a1] : 0;
      const int ks0 = (min(TC_KC, p.Cin - kc0 * TC_KC) + 15) >> 4;
      const int ks1 = (nit > 1) ? (min(TC_KC, p.Cin - kc1 * TC_KC) + 15) >> 4 : 0;
      const int nseg0 = 4 * ks0, nseg1 = 4 * ks1;  // (chunk, part) segments of nt*16 bytes
      if (p.skip & 8) {
        if (lane == 0) mbar_arrive(bar);
        if (++ps == NP) { ps = 0; ub++; }
        kc += 2;
        while (kc >= nkc) { kc -= nkc; a_idx++; }
        continue;
      }
      if (lane == 0) {
        const uint32_t bytes = (uint32_t)(nseg0 + nseg1) * (uint32_t)nt * 16u;
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
      }
      __syncwarp();
      const float4 *g0 = reinterpret_cast<const float4 *>(p.Wp) + ((size_t)o0 * nkc + kc0) * (size_t)N * 8;
      const float4 *g1 = reinterpret_cast<const float4 *>(p.Wp) + ((size_t)o1 * nkc + kc1) * (size_t)N * 8;
      const uint32_t bb0 = smem_u32(bring + (size_t)(2 * ps) * bstage_bytes), bb1 = bb0 + bstage_bytes;
      if (nt == N) {
        if (lane == 0)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(bb0), "l"(g0), "r"((uint32_t)nseg0 * (uint32_t)nt * 16u), "r"(bar) : "memory");
        if (lane == 1 && nit > 1)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(bb1), "l"(g1), "r"((uint32_t)nseg1 * (uint32_t)nt * 16u), "r"(bar) : "memory");
      } else if (lane < nseg0 + nseg1) {  // column split: one bulk copy per (chunk, part) segment
        const bool second = lane >= nseg0;
        const int sg = second ? lane - nseg0 : lane;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((second ? bb1 : bb0) + (uint32_t)(sg * nt) * 16u), "l"((second ? g1 : g0) + (size_t)sg * N + n0),
                     "r"((uint32_t)nt * 16u), "r"(bar)
                     : "memory");
      }
      if (++ps == NP) { ps = 0; ub++; }
      kc += 2;
      while (kc >= nkc) { kc -= nkc; a_idx++; }
    }
  } else if (warp_u == 8) {
    // ---- MMA issuer: per 16-channel k-step two instructions
    //        D[:, 0:2nt]  += A_hi * [B_hi | B_lo]     (N = 2nt)
    //        D[:, nt:2nt] += A_lo *  B_hi             (N = nt)
    //      so columns [0,nt) hold hi*hi and [nt,2nt) the two correction products (summed in the epilogue).
    //      The WHOLE warp runs this loop and one elected lane issues: with warp-uniform control flow and operands the
    //      compiler keeps descriptors in uniform registers and emits bare UTCHMMA; issued from a divergent
    //      `lane == 0` branch every instruction was wrapped in an R2UR + VOTEU/ELECT/BRA.U.ANY waterfall loop, and that
    //      single thread's instruction stream (not the tensor pipe) paced the whole kernel.
    // instruction descriptor: D = f32 (1 << 4), A = B = f16 (format 0), both K-major, N >> 3, M >> 4
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const int total_u = __shfl_sync(0xffffffffu, total, 0);
    const int npairs_u = (total_u + 1) >> 1;
    const int k0_u = __shfl_sync(0xffffffffu, k0, 0);
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * nt) >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
    const uint32_t idesc1 = (1u << 4) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
    const uint32_t b_lbo = (uint32_t)(2 * nt) * 16;
    const uint64_t b_step = (uint64_t)((2 * b_lbo) >> 4);
    const uint32_t bring_u = smem_u32(bring);
    const uint32_t bars_u = smem_u32(&bars[0]);
    const uint32_t acol0 = tmem_u + (uint32_t)p.tmem_acols;
    uint32_t first = 0u;  // 0 for the very first MMA (overwrite), then 1
    int ps = 0, kc = k0_u;
    uint32_t par = 0u;
    for (int P = 0; P < npairs_u; P++) {
      const bool dbg_on = GATHER == 0 && p.dbg && (p.skip & 32) && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0 && P < 64 && leader;
      if (dbg_on) s_dbg[P * 8 + 4] = clock64();
      mbar_wait(bars_u + 8u * (uint32_t)(BAR_BFULL + ps), par);
      if (dbg_on) s_dbg[P * 8 + 5] = clock64();
      mbar_wait(bars_u + 8u * (uint32_t)(BAR_FULL + ps), par);
      if (dbg_on) s_dbg[P * 8 + 6] = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int nit = min(2, total_u - 2 * P);
      for (int h = 0; h < nit; h++) {
        const int kvalid = min(TC_KC, p.Cin - kc * TC_KC);
        const int ksteps = (kvalid + 15) >> 4;
        const uint32_t sbt = bring_u + (uint32_t)(2 * ps + h) * bstage_bytes;
        uint64_t dbb = umma_desc(sbt, b_lbo, 128);
        uint32_t ah = acol0 + (uint32_t)((2 * ps + h) * 32), al = ah + 16u;
        for (int ks = 0; ks < ksteps; ks++) {
          if (leader && !(p.skip & 2)) {
            umma_f16_ts(tmem_u, ah, dbb, idesc2, first);
            umma_f16_ts(tmem_u + (uint32_t)nt, al, dbb, idesc1, 1u);
          }
          first = 1u;
          ah += 8u; al += 8u; dbb += b_step;
        }
        if (++kc == nkc) kc = 0;
      }
      if (leader) umma_commit(bars_u + 8u * (uint32_t)(BAR_FREE + ps));  // frees both A stages and both weight stages of the pair
      __syncwarp();
      if (dbg_on) s_dbg[P * 8 + 7] = clock64();
      if (++ps == NP) { ps = 0; par ^= 1u; }
    }
    if (leader && total_u > 0) umma_commit(bars_u + 8u * 8u);
    __syncwarp();
  }
  // ---- epilogue: TMEM lane = output row; warp w reads lanes 32*(w%4).., column half w/4 -----------------
  const bool split = S > 1;
  if (mark_on) p.dbg[64 * 8 + 4] = clock64();
  const uint32_t pbuf = smem_u32(bring);  // split-K partial tile [nt columns][128 rows] f32 (the weight ring is idle by now)
  if (split) {
    if (producer && z > 0) {
      if (total > 0) {
        mbar_wait(smem_u32(&bars[8]), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const int lane_grp = warp & 3, chalf = warp >> 2;
      const int cbeg = chalf * (nt >> 1), cend = cbeg + (nt >> 1);
      for (int cb = cbeg; cb < cend; cb += 8) {
        uint32_t v[8], c[8];
        if (total > 0) {
          uint32_t taddr = tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)cb;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                       : "r"(taddr));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7])
                       : "r"(taddr + (uint32_t)nt));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int e = 0; e < 8; e++) v[e] = c[e] = 0u;
        }
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const float x = fmaf(__uint_as_float(c[e]), kLoInv, __uint_as_float(v[e]));
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(pbuf + (uint32_t)(((cb + e) * TC_ROWS + lane_grp * 32 + lane) * 4)), "f"(x) : "memory");
        }
      }
    }
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (producer && (!split || z == 0)) {
    if (total > 0) {
      mbar_wait(smem_u32(&bars[8]), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (mark_on) p.dbg[64 * 8 + 5] = clock64();
    const int lane_grp = warp & 3, chalf = warp >> 2;
    const int row = row0 + lane_grp * 32 + lane;
    const int cbeg = chalf * (nt >> 1), cend = cbeg + (nt >> 1);
    for (int cb = cbeg; cb < cend; cb += 8) {
      uint32_t v[8], c[8];
      if (total > 0) {
        uint32_t taddr = tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)cb;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7])
                     : "r"(taddr + (uint32_t)nt));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = __float_as_uint(fmaf(__uint_as_float(c[e]), kLoInv, __uint_as_float(v[e])));
      } else {
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = 0u;
      }
      if (split) {  // add the other ranks' partial tiles in rank order (distributed shared memory)
        for (int zz = 1; zz < S; zz++) {
          uint32_t rbase;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbase) : "r"(pbuf), "r"(zz));
          float t[8];
#pragma unroll
          for (int e = 0; e < 8; e++)
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(t[e]) : "r"(rbase + (uint32_t)(((cb + e) * TC_ROWS + lane_grp * 32 + lane) * 4)) : "memory");
#pragma unroll
          for (int e = 0; e < 8; e++) v[e] = __float_as_uint(__uint_as_float(v[e]) + t[e]);
        }
      }
      if (row < p.Mout) {
        const int col = n0 + cb;
        float *op = p.out + (size_t)row * p.out_stride + p.out_off + col;
        const float *rp = p.residual ? p.residual + (size_t)row * p.res_stride + p.res_off + col : nullptr;
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = __uint_as_float(v[e]);
        const bool full = col + 8 <= p.Cout;
        if (p.bias) {
#pragma unroll
          for (int e = 0; e < 8; e++)
            if (col + e < p.Cout) x[e] += __ldg(&p.bias[col + e]);
        }
        if (rp) {
          if (full && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
            const float4 r0 = __ldg(reinterpret_cast<const float4 *>(rp)), r1 = __ldg(reinterpret_cast<const float4 *>(rp) + 1);
            x[0] += r0.x; x[1] += r0.y; x[2] += r0.z; x[3] += r0.w;
            x[4] += r1.x; x[5] += r1.y; x[6] += r1.z; x[7] += r1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; e++)
              if (col + e < p.Cout) x[e] += __ldg(rp + e);
          }
        }
        if (full && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
          reinterpret_cast<float4 *>(op)[0] = make_float4(x[0], x[1], x[2], x[3]);
          reinterpret_cast<float4 *>(op)[1] = make_float4(x[4], x[5], x[6], x[7]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; e++)
            if (col + e < p.Cout) op[e] = x[e];
        }
      }
    }
  }
  if (split) {  // ranks > 0 keep their shared memory alive until rank 0 has read it
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (mark_on) p.dbg[64 * 8 + 6] = clock64();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols) : "memory");
  }
  if (mark_on) p.dbg[64 * 8 + 7] = clock64();
  if constexpr (GATHER == 0) {
    if (mark_on && (p.skip & 32))
      for (int i = 0; i < 64 * 8; i++) p.dbg[i] = s_dbg[i];
  }
}

}  // namespace sgb

using namespace sgb;

namespace sgb {
// y = BatchNorm(eval)+ReLU(x) (or x when scale == nullptr), split into fp16 hi/lo and packed for the tensor-core
// kernel: per row and per 32-channel chunk, 16 words of hi pairs (channel 2c in the low half) then 16 words of lo
// pairs. One thread per (row, chunk, word pair); channels past C are zero.
__global__ void act_split_kernel(const float *__restrict__ x, int x_stride, int x_off, const float *__restrict__ scale,
                                 const float *__restrict__ shift, int relu, uint32_t *__restrict__ y, int M, int C, int Cpad) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int wpr = Cpad >> 1;  // word pairs (2 channels) per row
  if (t >= (long long)M * wpr) return;
  const int row = (int)(t / wpr), pr = (int)(t % wpr);
  const int c = 2 * pr;  // first channel of the pair
  float a = 0.f, b = 0.f;
  if (c < C) a = x[(size_t)row * x_stride + x_off + c];
  if (c + 1 < C) b = x[(size_t)row * x_stride + x_off + c + 1];
  if (scale) {
    if (c < C) a = fmaf(a, __ldg(&scale[c]), __ldg(&shift[c]));
    if (c + 1 < C) b = fmaf(b, __ldg(&scale[c + 1]), __ldg(&shift[c + 1]));
    if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
  }
  const __half2 h = f2h2_sat(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = f2h2_sat((a - hf.x) * kLoScale, (b - hf.y) * kLoScale);
  const int chunk = c >> 5, w = (c & 31) >> 1;  // word index inside the chunk's hi block
  uint32_t *yr = y + (size_t)row * Cpad + chunk * 32;
  yr[w] = *reinterpret_cast<const uint32_t *>(&h);
  yr[16 + w] = *reinterpret_cast<const uint32_t *>(&l);
}
}  // namespace sgb

static long long *g_tc_dbg = nullptr;
static int g_tc_prefetch = 0, g_tc_maxb = 3, g_tc_maxsplit = 8, g_tc_skip = 0, g_tc_split_policy = 0, g_tc_gather = 0;

extern "C" {

void sgb_test_set_tc_debug(long long *d_buf) { g_tc_dbg = d_buf; }
void sgb_test_set_tc_skip(int mask) { g_tc_skip = mask; }
void sgb_test_set_tc_split_policy(int policy) { g_tc_split_policy = policy; }
void sgb_test_set_tc_gather(int mode) { g_tc_gather = mode; }  // 0: per-lane gather (validated), 1: cooperative gather  // 0: validated heuristic, 1: wave-aware
// test/bench hook: weight prefetch on/off, most pair stages in the ring (1..3), largest split-K cluster (1 = off)
void sgb_test_set_tc_tuning(int prefetch, int max_pairs, int max_ksplit) {
  g_tc_prefetch = prefetch;
  g_tc_maxb = std::max(1, std::min(max_pairs, 3));
  g_tc_maxsplit = std::max(1, std::min(max_ksplit, 8));
}

int sgb_act_split(const float *d_x, int x_stride, int x_off, const float *d_scale, const float *d_shift, int relu,
                  float *d_y, int M, int C, void *stream) {
  if (M == 0 || C == 0) return SGB_OK;
  SGB_REQUIRE(d_x && d_y && M > 0 && C > 0 && (d_scale == nullptr) == (d_shift == nullptr), SGB_ERR_ARG, "act_split arguments");
  int Cpad = (C + 31) / 32 * 32;
  long long tot = (long long)M * (Cpad / 2);
  act_split_kernel<<<div_up(tot, 256), 256, 0, (cudaStream_t)stream>>>(d_x, x_stride, x_off, d_scale, d_shift, relu,
                                                                     (uint32_t *)d_y, M, C, Cpad);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}

// log2 of the factor the fp16 remainders (weights AND activations) are scaled by; the host packer must use the same.
int sgb_spconv_tc_lo_shift(void) { return sgb::kLoShift; }

// Packed weight size in floats for sgb_spconv_forward_tc: K * ceil(Cin/32) * 8 * N * 4 with N = Cout rounded to 16.
long long sgb_spconv_tc_packed_floats(int K, int Cin, int Cout) {
  int N = (Cout + 15) / 16 * 16;
  int nkc = (Cin + 31) / 32;
  return (long long)K * nkc * 4 * 2 * N * 4;  // fp16: 4 chunks of 8 halves, counted in 32-bit words
}

int sgb_spconv_forward_tc(const float *d_in, int in_stride, int in_off, const int32_t *d_map, int K, int Mout,
                          const float *d_Wp, int Cin, int Cout, const float *d_in_scale,
                          const float *d_in_shift, const float *d_residual, int res_stride, int res_off,
                          const float *d_bias, float *d_out, int out_stride, int out_off, int in_packed, void *stream) {
  if (Mout == 0 || Cout == 0) return SGB_OK;
  SGB_REQUIRE(!in_packed || (d_in_scale == nullptr && (in_stride & 31) == 0 && (in_off & 31) == 0), SGB_ERR_ARG,
              "packed input: no scale/shift, row stride and offset multiples of 32 words");
  SGB_REQUIRE(d_in && d_Wp && d_out && K >= 1 && Mout > 0 && Cin > 0 && Cout > 0, SGB_ERR_ARG,
              "spconv_forward_tc arguments");
  SGB_REQUIRE(d_map || K == 1, SGB_ERR_ARG, "identity map requires K == 1");
  SGB_REQUIRE((d_in_scale == nullptr) == (d_in_shift == nullptr), SGB_ERR_ARG, "scale/shift must come together");
  SGB_REQUIRE((in_packed || in_stride >= in_off + Cin) && out_stride >= out_off + Cout, SGB_ERR_ARG, "row strides");
  int N = (Cout + 15) / 16 * 16;
  SGB_REQUIRE(N <= 256 && Cin <= 512, SGB_ERR_RANGE, "spconv_forward_tc: Cout > 256 or Cin > 512 is not tiled");
  TcArgs p;
  p.in = d_in; p.in_stride = in_stride; p.in_off = in_off;
  p.map = d_map; p.K = K; p.Mout = Mout;
  p.Wp = d_Wp; p.Cin = Cin; p.N = N; p.Cout = Cout;
  p.in_scale = d_in_scale; p.in_shift = d_in_shift;
  p.residual = d_residual; p.res_stride = res_stride; p.res_off = res_off;
  p.bias = d_bias;
  p.out = d_out; p.out_stride = out_stride; p.out_off = out_off;
  p.dbg = g_tc_dbg;
  p.skip = g_tc_skip;
  p.in_packed = in_packed;
  // Column split: few row tiles (deep U-Net levels) would leave most SMs idle and make one CTA stream all the
  // weights, so N is cut into NT-column CTAs until the grid covers the machine (NT multiple of 16, >= 32).
  int tiles = div_up(Mout, TC_ROWS);
  int NT = std::min(N, 128);  // [B_hi | B_lo] is one operand with 2*NT <= 256 columns
  if (N > 128) NT = (N / 2 + 15) / 16 * 16;
  while (NT > 32 && tiles * div_up(N, NT) < kNumSMs) {
    int nxt = (NT / 2 + 15) / 16 * 16;
    if (nxt >= NT) break;
    NT = nxt;
  }
  p.NT = NT;
  size_t bstage = 2 * (size_t)NT * TC_KC * 2;      // weight ring stage (fp16 hi + lo) in shared memory
  size_t map_bytes = d_map ? (size_t)K * TC_ROWS * 4 : 0;
  // TMEM budget: accumulator [main | corrections] = 2*NT columns, then the A ring (32 columns per stage: hi + lo pairs;
  // an even number of stages because the pipeline moves in pairs). 256 columns (two CTAs per SM) when that leaves
  // >= 4 stages (two pairs, so gathers overlap the MMAs), else all 512. The weight ring in shared memory has the same number of stages.
  int dcols = 2 * NT;
  int acols0 = (dcols + 31) / 32 * 32;
  int ns = ((256 - acols0) / 32) & ~1;
  if (ns >= 4) { p.tmem_cols = 256; }
  else { p.tmem_cols = 512; ns = ((512 - acols0) / 32) & ~1; }
  ns = std::min(ns, std::min(6, 2 * g_tc_maxb));
  const size_t ring_cap = (p.tmem_cols == 256) ? 64 * 1024 : 100 * 1024;  // one CTA per SM when it owns all of TMEM
  while (ns > 2 && (size_t)ns * bstage > ring_cap) ns -= 2;
  p.nstages = ns;
  p.nbstages = ns;
  p.tmem_acols = acols0;
  p.prefetch = g_tc_prefetch;
  // split-K: when even the narrowest column split leaves most SMs idle, a cluster of S CTAs shares each tile's
  // iterations (S <= 8, the portable cluster size) and reduces through distributed shared memory.
  const int nkc = (Cin + TC_KC - 1) / TC_KC;
  int ctas = tiles * div_up(N, NT), S = 1;
  if (ctas * 2 <= kNumSMs && p.tmem_cols == 256 && (size_t)p.nbstages * bstage >= (size_t)NT * TC_ROWS * 4) {
    S = std::min({g_tc_maxsplit, 2 * kNumSMs / ctas, std::max(1, K * nkc / 4)});
    S = std::max(S, 1);
  } else if (g_tc_split_policy == 1 && p.tmem_cols == 256 && (size_t)p.nbstages * bstage >= (size_t)NT * TC_ROWS * 4) {
    // round-2 candidate (off by default): wave quantisation. With 2 CTAs per SM there are 296 slots; 162 CTAs (level 3)
    // leave half of them empty and 332 CTAs (level 2) run a second, almost empty wave. Pick the S in {1,2,4} with the
    // fewest (waves / S), i.e. the shortest critical path in units of whole-tile time.
    const int slots = 2 * kNumSMs;
    int best = 1;
    double best_cost = (double)div_up(ctas, slots);
    for (int cand = 2; cand <= std::min(4, g_tc_maxsplit); cand *= 2) {
      if (K * nkc / cand < 4) break;
      const double cost = (double)div_up(ctas * cand, slots) / cand;
      if (cost < best_cost * 0.9) { best = cand; best_cost = cost; }
    }
    S = best;
  }
  p.ksplit = S;
  size_t smem = bstage * p.nbstages + map_bytes + 1024;
  const bool coop = g_tc_gather == 1 && in_packed;  // round-2 candidate: cooperative gather through shared memory
  p.gather_off = 0;
  if (coop) {
    p.gather_off = (int)align_up(bstage * p.nbstages + map_bytes);
    smem = (size_t)p.gather_off + 8 * 8192 + 1024;
  }
  void (*kern)(TcArgs) = coop ? spconv_tc_kernel<1> : spconv_tc_kernel<0>;
  static bool attr_set = false;
  if (!attr_set) {
    SGB_CUDA_CHECK(cudaFuncSetAttribute(spconv_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(160 * 1024)));
    SGB_CUDA_CHECK(cudaFuncSetAttribute(spconv_tc_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    SGB_CUDA_CHECK(cudaFuncSetAttribute(spconv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    SGB_CUDA_CHECK(cudaFuncSetAttribute(spconv_tc_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set = true;
  }
  dim3 grid(tiles, div_up(N, NT), S);
  if (S == 1) {
    kern<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = S;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SGB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
  }
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}
}
