// spconv_ss.cu -- sparse convolution on tcgen05, persistent kernel with a shared-memory gather ring (sm_100a).
//
// Same arithmetic and data formats as spconv_tc.cu (packed fp16 hi/lo activation rows, error-compensated products
// hi*hi + hi*lo + lo*hi accumulated in fp32 in tensor memory), different data movement. Measurements that
// shaped it (profiles/r2_umma_rate2.txt, profiles/README.md round 2):
//   * the tensor pipe needs max(~48, N/2) cycles per M = 128, K = 16 instruction (r2_umma_rate3.txt: the minimum does not
//     overlap across issuing warps or CTAs) and holds the issuing thread until it takes the instruction; what that thread
//     does between stages leaves the pipe idle -> the issue loop is unrolled per stage (up to 8 UTCHMMA + 1 commit), with
//     descriptors advanced by adds;
//   * a gathered row slice held in registers gives a prefetch distance of one iteration (tcgen05.wait::st drains the
//     thread's outstanding loads), so the register-gather kernel's iteration time was the L2 latency -> rows are copied
//     with cp.async (16 B per lane, 8 lanes per 128-byte line, zero-fill for absent neighbours) straight into the
//     SWIZZLE_128B K-major tile the MMA reads, several stages in flight, no registers held;
//   * 256 per-thread arrivals on one mbarrier serialise in the shared-memory atomic unit -> ONE warp owns a whole stage
//     (32 copy instructions of 4 rows each), waits for its own copies (cp.async.wait_group), makes them visible to the
//     async proxy (fence.proxy.async) and arrives once.
// A ring stage holds a PAIR of iterations (2 x 16 KB of rows + 2 weight slices), so the MMA warp and the weight loader pay
// their per-stage costs (an mbarrier wait is ~90 cycles even when it is already complete) once per two iterations.
// One CTA per SM (all 512 TMEM columns, ~200 KB of shared memory) walks work items (128 output rows x NT columns)
// round-robin; accumulators are double-buffered in TMEM so the epilogue of item j overlaps the pipeline of item j+1, and
// the rulebook slice of item j+1 is staged by its own warp while item j runs.
// Warp roles (480 threads): 0-7 gather (warp w owns the iteration slots g with g % gw == w, gw = min(8, 2 S - 2)), 8 MMA
// issue, 9 weight loader (TMA bulk copies), 10 rulebook loader, 11-14 epilogue (TMEM lane group = warp % 4).
// What bounds it now is the turnover of the ring (slot free -> 32 copies issued -> landed -> MMA -> commit, ~2 500 cycles for
// 3-4 pair stages in 227 KB): tried without gain -- two MMA-issuing warps, the next pair's barrier wait between the two slots
// of a pair, per-slot release of the ring (profiles/README.md, ROUND2_NOTES.md).
#include <algorithm>
#include <cstdlib>

#include <cuda_fp16.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace sgb {

#ifdef SGB_SS_TIMELINE
#define TL_DECL long long tl_w0 = 0, tl_w1 = 0, tl_w2 = 0, tl_t = 0; const bool tl_on = p.dbg && blockIdx.x == 0 && lane == 0
#define TL_T0() do { if (tl_on) tl_t = clock64(); } while (0)
#define TL_ADD(v) do { if (tl_on) v += clock64() - tl_t; } while (0)
#define TL_STAMP(slot, idx) do { if (p.dbg && blockIdx.x == 0 && lane == 0 && (idx) < 256) p.dbg[64 + (slot) * 256 + (idx)] = clock64(); } while (0)
#else
#define TL_DECL
#define TL_T0()
#define TL_ADD(v)
#define TL_STAMP(slot, idx)
#endif

constexpr int S2_ROWS = 128;
constexpr int S2_KC = 32;                    // channels per iteration: one 128-byte packed line per row
constexpr int S2_GW = 8;                     // gather warps
constexpr int S2_THREADS = 480;              // 15 warps
constexpr int S2_MAXS = 12;                  // ring depth limit
constexpr int S2_A_BYTES = S2_ROWS * 128;    // 16 KB per stage
constexpr float kSsLoInv = 1.0f / 2048.0f;   // 2^-kLoShift (spconv_tc.cu: kLoShift = 11; checked by a static_assert there)
constexpr float kSsLoScale = 2048.0f;

struct SsArgs {
  const uint32_t *in; int in_stride;     // packed input rows [Min][in_stride words]
  const int32_t *map; int K, Mout;       // map [K][Mout] (nullptr: identity, K == 1)
  const float *Wp;                       // packed weights [K][nkc][4][2][N][8 halves]
  int Cin, N, Cout, NT, nparts;          // N = Cout rounded up to 16; column parts of NT (the last may be shorter)
  const float *residual; int res_stride, res_off;
  const float *bias;
  float *out; int out_stride, out_off;   // optional fp32 rows
  uint32_t *pk; int pk_stride, pk_coff;  // optional packed rows (the consumer's input)
  const float *pk_scale, *pk_shift; int pk_relu;
  int pk_fill;
  int *oflow;
  int S, items;
  long long *dbg;                        // development builds (SGB_SS_TIMELINE) only
  int dev_flags;                         // development builds: 1 = skip the weight copies, 2 = skip the row copies
};

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {  // K-major SWIZZLE_128B, 8-row groups 1024 B apart
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(S2_THREADS, 1) spconv_ss_kernel(SsArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar_full[S2_MAXS], bar_empty[S2_MAXS];
  __shared__ __align__(8) unsigned long long bar_accf[2], bar_acce[2], bar_mapf[2], bar_mape[2];
  __shared__ uint32_t s_tmem;
  __shared__ int s_list[2][32];
  __shared__ int s_nact[2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int S = p.S, K = p.K, NT = p.NT;
  // A ring stage holds a PAIR of consecutive iterations (slot 0 / slot 1): the per-stage costs of the MMA warp and the weight
  // loader (one mbarrier wait ~90 cycles, descriptor set-up, one commit / one expect_tx) are paid once per two iterations.
  const uint32_t b_slice = (uint32_t)NT * 128u;  // one iteration's weights: [hi | lo] x 4 chunks x NT x 16 B
  const uint32_t b_stage = 2u * b_slice;
  unsigned char *a_ring = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);  // SWIZZLE_128B tiles: 1024-byte aligned
  unsigned char *b_ring = a_ring + (size_t)S * (2 * S2_A_BYTES);
  int32_t *map_s = reinterpret_cast<int32_t *>(b_ring + (size_t)S * b_stage);  // [2][K][128]
  const int nkc = (p.Cin + S2_KC - 1) / S2_KC;
  const int ks_last = (p.Cin - (nkc - 1) * S2_KC + 15) >> 4;  // 16-channel k-steps of the last slice (1 or 2)

  if (tid == 0) {
    for (int s = 0; s < S; s++) {
      mbar_init(smem_u32(&bar_full[s]), 3);   // the gather warps of slot 0 and slot 1 + the weight loader (expect_tx)
      mbar_init(smem_u32(&bar_empty[s]), 1);  // tcgen05.commit
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(smem_u32(&bar_accf[b]), 1);   // tcgen05.commit (or a plain arrive for an item without active pairs)
      mbar_init(smem_u32(&bar_acce[b]), 4);   // one arrival per epilogue warp
      mbar_init(smem_u32(&bar_mapf[b]), 1);   // rulebook loader
      mbar_init(smem_u32(&bar_mape[b]), S2_GW + 6);  // gather warps, MMA warp, weight loader, 4 epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
#ifdef SGB_SS_TIMELINE
  if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[13] = clock64();
#endif
  const int acc_cols = (2 * NT + 31) / 32 * 32;  // columns of one accumulator buffer (<= 256)
  const int first = blockIdx.x, stride = gridDim.x;

  if (warp < S2_GW) {
    // =========================== gather: warp w fills ring uses g with g % 8 == w =================================
    TL_DECL;
    const uint32_t a_base_u = smem_u32(a_ring);
    const int sub = lane >> 3, ch = lane & 7;  // one instruction copies 4 rows x 8 chunks of 16 B
    // destination offset of (row 4 j + sub, chunk ch) inside a stage: rows are 128 B apart, chunk position ^ (row & 7)
    const uint32_t dst_even = (uint32_t)sub * 128u + (uint32_t)((ch ^ sub) << 4);        // j even: row & 7 = sub
    const uint32_t dst_odd = (uint32_t)sub * 128u + (uint32_t)((ch ^ (sub + 4)) << 4);   // j odd:  row & 7 = sub + 4
    const uint32_t *gcol0 = p.in + ch * 4;
    int n = 0, g = 0;
    // Only gw <= 2 S - 2 warps copy: a warp's previous slot use (g - gw, pair P' >= P - gw/2 - 1) has seen the MMA warp pass
    // pair P' - S, so the `empty` barrier of pair P is at most one phase behind when it is tested (its parity test would
    // alias otherwise).
    const int gw = max(1, min(S2_GW, 2 * S - 2));
    int pending_s = -1;  // stage of the copy group this warp has in flight (not yet signalled)
    int pending_g = 0;   // its ring use (timeline builds)
    auto flush = [&](int allow) {  // wait until at most `allow` groups are in flight, then signal the older one
      if (allow == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
      else asm volatile("cp.async.wait_group 1;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_full[pending_s]));
      TL_STAMP(2, pending_g);
    };
    for (int item = first; item < p.items; item += stride, n++) {
      const int buf = n & 1;
      if (pending_s >= 0 && !mbar_test(smem_u32(&bar_mapf[buf]), (uint32_t)((n >> 1) & 1))) {
        flush(0);  // never sit on an unsignalled stage while waiting for the next rulebook slice
        pending_s = -1;
      }
      TL_T0();
      mbar_wait(smem_u32(&bar_mapf[buf]), (uint32_t)((n >> 1) & 1));
      TL_ADD(tl_w0);
      const int total = s_nact[buf] * nkc;
      const int total2 = (total + 1) & ~1;  // an odd tail leaves slot 1 of the last pair empty: its owner only arrives
      const int32_t *ms = map_s + (size_t)buf * K * S2_ROWS;
      for (int i = (warp < gw) ? (warp + gw - g % gw) % gw : total2; i < total2; i += gw) {
        const int gi = g + i;
        const int P = gi >> 1, h = gi & 1;
        const int s = P % S, u = P / S;
        const uint32_t ebar = smem_u32(&bar_empty[s]);
        if (u >= 1 && !mbar_test(ebar, (uint32_t)((u - 1) & 1))) {
          // the ring slot is still being read: signal what is in flight first (the MMA warp may be waiting for exactly
          // that stage), then block
          if (pending_s >= 0) { flush(0); pending_s = -1; }
          TL_T0();
          mbar_wait(ebar, (uint32_t)((u - 1) & 1));
          TL_ADD(tl_w1);
        }
        TL_STAMP(0, gi);
        if (i < total) {
          const int a = i / nkc, kc = i - a * nkc;
          const int o = s_list[buf][a];
          const uint32_t stage = a_base_u + (uint32_t)(2 * s + h) * S2_A_BYTES;
          const int32_t *mrow = ms + o * S2_ROWS + sub;
          const uint32_t *gcol = gcol0 + kc * S2_KC;
          // all 32 rulebook entries first (independent shared-memory loads), then the copies: the "memory" clobber of the
          // copy instruction would otherwise order every entry load behind the previous copy (measured: 2 400 cycles per stage)
          int srcv[32];
#pragma unroll
          for (int j = 0; j < 32; j++) srcv[j] = mrow[4 * j];  // the 8 lanes of a quarter-warp read the same entry (broadcast)
#ifdef SGB_SS_TIMELINE
          if (!(p.dev_flags & 2))
#endif
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const int src = srcv[j];
            const uint32_t dst = stage + (uint32_t)j * 512u + ((j & 1) ? dst_odd : dst_even);
            const uint32_t *gp = gcol + (size_t)max(src, 0) * p.in_stride;
            const int nbytes = (src >= 0) ? 16 : 0;  // 0 source bytes = zero fill (absent neighbour)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gp), "r"(nbytes) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // (an empty group for the slot of an odd tail)
        TL_STAMP(1, gi);
        if (pending_s >= 0) {
          TL_T0();
          flush(1);
          TL_ADD(tl_w2);
        }
        pending_s = s;
        pending_g = gi;
      }
      g += total2;
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_mape[buf]));
    }
    if (pending_s >= 0) flush(0);
#ifdef SGB_SS_TIMELINE
    if (tl_on && warp == 0) { p.dbg[0] = tl_w0; p.dbg[1] = tl_w1; p.dbg[2] = tl_w2; p.dbg[3] = g; p.dbg[4] = n; }
#endif
  } else if (warp_u == 8) {
    // =========================== MMA issue (warp-uniform; one elected lane issues) ==============================
    TL_DECL;
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t full_u = __shfl_sync(0xffffffffu, smem_u32(&bar_full[0]), 0);
    const uint32_t empty_u = __shfl_sync(0xffffffffu, smem_u32(&bar_empty[0]), 0);
    const uint64_t ad0 = desc_sw128(__shfl_sync(0xffffffffu, smem_u32(a_ring), 0));
    const uint32_t b_base = __shfl_sync(0xffffffffu, smem_u32(b_ring), 0);
    const uint64_t a_inc = (uint64_t)((2 * S2_A_BYTES) >> 4), b_inc = (uint64_t)(b_stage >> 4);  // per ring stage (pair)
    const uint64_t a_slot = (uint64_t)(S2_A_BYTES >> 4), b_slot = (uint64_t)(b_slice >> 4);      // slot 1 of a pair
    int n = 0, s = 0, gtl = 0;
    uint32_t par = 0u;
    uint64_t ad = ad0, bd = 0;
    for (int item = first; item < p.items; item += stride, n++) {
      const int buf = n & 1, ab = n & 1;
      const int part = item % p.nparts;
      const int nt = min(NT, p.N - part * NT);
      TL_T0();
      mbar_wait(smem_u32(&bar_mapf[buf]), (uint32_t)((n >> 1) & 1));
      TL_ADD(tl_w0);
      const int total = __shfl_sync(0xffffffffu, s_nact[buf], 0) * nkc;
      __syncwarp();
      if (leader) mbar_arrive(smem_u32(&bar_mape[buf]));
      TL_T0();
      if (n >= 2) mbar_wait(smem_u32(&bar_acce[ab]), (uint32_t)(((n >> 1) - 1) & 1));
      TL_ADD(tl_w1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d = tmem_u + (uint32_t)(ab * acc_cols), dc = d + (uint32_t)nt;
      // instruction descriptor: D = f32 (1 << 4), A = B = f16, both K-major, N >> 3, M >> 4
      const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * nt) >> 3) << 17) | ((uint32_t)(S2_ROWS >> 4) << 24);
      const uint32_t idesc1 = (1u << 4) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(S2_ROWS >> 4) << 24);
      const uint32_t b_lbo = (uint32_t)(2 * nt) * 16u;             // distance between the two 8-channel chunks of a k-step
      const uint64_t b_step = (uint64_t)((2 * b_lbo) >> 4);        // next k-step
      const uint64_t bd0 = umma_desc(b_base, b_lbo, 128);
      bd = bd0 + b_inc * (uint64_t)s;
      uint32_t acc = 0u;
      int kc = 0;
      const int npairs = (total + 1) >> 1;
      // (Tried: waiting for the NEXT pair's barrier between the two slots of the current pair, to hide the >= 90 cycles of an
      // mbarrier wait behind the pipe. Slower -- GPU call 19: when the next pair is not ready yet the current pair's commit is
      // delayed and the ring, which is what bounds the kernel, turns over later.)
      for (int P = 0; P < npairs; P++) {
        TL_T0();
        mbar_wait(full_u + 8u * (uint32_t)s, par);
        TL_ADD(tl_w2);
        TL_STAMP(3, gtl);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool two0 = (kc + 1 < nkc) || ks_last == 2;
        if (++kc == nkc) kc = 0;
        const bool has1 = 2 * P + 1 < total;
        const bool two1 = (kc + 1 < nkc) || ks_last == 2;
        if (has1 && ++kc == nkc) kc = 0;
        if (leader) {
          umma_f16_ss(d, ad, bd, idesc2, acc);            // slot 0: A_hi [B_hi | B_lo]
          umma_f16_ss(dc, ad + 4u, bd, idesc1, 1u);       //         A_lo B_hi  (A_lo: +64 B inside the 128-byte row)
          if (two0) {
            umma_f16_ss(d, ad + 2u, bd + b_step, idesc2, 1u);   // channels 16..31: +32 B
            umma_f16_ss(dc, ad + 6u, bd + b_step, idesc1, 1u);
          }
          if (has1) {
            const uint64_t ad1 = ad + a_slot, bd1 = bd + b_slot;
            umma_f16_ss(d, ad1, bd1, idesc2, 1u);
            umma_f16_ss(dc, ad1 + 4u, bd1, idesc1, 1u);
            if (two1) {
              umma_f16_ss(d, ad1 + 2u, bd1 + b_step, idesc2, 1u);
              umma_f16_ss(dc, ad1 + 6u, bd1 + b_step, idesc1, 1u);
            }
          }
          umma_commit(empty_u + 8u * (uint32_t)s);
        }
        __syncwarp();
        TL_STAMP(4, gtl);
        gtl++;
        acc = 1u;
        ad += a_inc; bd += b_inc;
        if (++s == S) { s = 0; par ^= 1u; ad = ad0; bd = bd0; }
      }
      if (leader) {
        if (total > 0) umma_commit(smem_u32(&bar_accf[ab]));
        else mbar_arrive(smem_u32(&bar_accf[ab]));
      }
      __syncwarp();
    }
#ifdef SGB_SS_TIMELINE
    if (tl_on) { p.dbg[5] = tl_w0; p.dbg[6] = tl_w1; p.dbg[7] = tl_w2; }
#endif
  } else if (warp == 9) {
    // =========================== weight loader ================================================================
    TL_DECL;
    int n = 0, s = 0, u = 0, gtl = 0;
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t b_base = __shfl_sync(0xffffffffu, smem_u32(b_ring), 0);
    for (int item = first; item < p.items; item += stride, n++) {
      const int buf = n & 1;
      const int part = item % p.nparts;
      const int n0 = part * NT;
      const int nt = min(NT, p.N - n0);
      mbar_wait(smem_u32(&bar_mapf[buf]), (uint32_t)((n >> 1) & 1));
      const int nact = s_nact[buf];
      const int my_o = s_list[buf][lane < nact ? lane : 0];  // lane l keeps the l-th active offset
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_mape[buf]));
      const int total = nact * nkc;
      const int npairs = (total + 1) >> 1;
      const size_t slice_f4 = (size_t)p.N * 8;  // float4s of one (offset, slice) weight block
      int a = 0, kc = 0;
      for (int P = 0; P < npairs; P++) {
        // slot 0 / slot 1 of the pair: (offset list position, channel slice)
        const int o0 = __shfl_sync(0xffffffffu, my_o, a), kc0 = kc;
        int a1 = a, kc1 = kc + 1;
        if (kc1 == nkc) { kc1 = 0; a1++; }
        const bool has1 = 2 * P + 1 < total;
        const int o1 = __shfl_sync(0xffffffffu, my_o, has1 ? a1 : a);
        TL_T0();
        if (u >= 1) mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)((u - 1) & 1));
        TL_ADD(tl_w0);
        const uint32_t bar = smem_u32(&bar_full[s]);
        const int ks0 = (kc0 + 1 < nkc) ? 2 : ks_last, ks1 = has1 ? ((kc1 + 1 < nkc) ? 2 : ks_last) : 0;
        const uint32_t seg_bytes = (uint32_t)nt * 16u;  // one (chunk, hi|lo) segment
        const uint32_t bytes0 = 4u * (uint32_t)ks0 * seg_bytes, bytes1 = 4u * (uint32_t)ks1 * seg_bytes;
        const float4 *g0 = reinterpret_cast<const float4 *>(p.Wp) + ((size_t)o0 * nkc + kc0) * slice_f4;
        const float4 *g1 = reinterpret_cast<const float4 *>(p.Wp) + ((size_t)o1 * nkc + kc1) * slice_f4;
        const uint32_t dst0 = b_base + (uint32_t)s * b_stage, dst1 = dst0 + b_slice;
#ifdef SGB_SS_TIMELINE
        if (p.dev_flags & 1) {
          if (leader) mbar_arrive(bar);
        } else
#endif
        if (leader) {
          mbar_expect_tx(bar, bytes0 + bytes1);
          if (nt == p.N) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst0), "l"(g0), "r"(bytes0), "r"(bar) : "memory");
            if (has1)
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                           ::"r"(dst1), "l"(g1), "r"(bytes1), "r"(bar) : "memory");
          } else {
            for (int sg = 0; sg < 4 * ks0; sg++)
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                           ::"r"(dst0 + (uint32_t)sg * seg_bytes), "l"(g0 + (size_t)sg * p.N + n0), "r"(seg_bytes), "r"(bar) : "memory");
            for (int sg = 0; sg < 4 * ks1; sg++)
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                           ::"r"(dst1 + (uint32_t)sg * seg_bytes), "l"(g1 + (size_t)sg * p.N + n0), "r"(seg_bytes), "r"(bar) : "memory");
          }
        }
        __syncwarp();
        TL_STAMP(5, gtl);
        gtl++;
        kc = kc1; a = a1;
        if (has1 && ++kc == nkc) { kc = 0; a++; }
        if (++s == S) { s = 0; u++; }
      }
    }
#ifdef SGB_SS_TIMELINE
    if (tl_on) p.dbg[8] = tl_w0;
#endif
  } else if (warp == 10) {
    // =========================== rulebook loader: slice of the item -> shared memory, active offsets in order =====
    TL_DECL;
    int n = 0;
    for (int item = first; item < p.items; item += stride, n++) {
      const int buf = n & 1;
      TL_T0();
      if (n >= 2) mbar_wait(smem_u32(&bar_mape[buf]), (uint32_t)(((n >> 1) - 1) & 1));
      TL_ADD(tl_w0);
      const int tile = item / p.nparts;
      const int row0 = tile * S2_ROWS;
      int32_t *ms = map_s + (size_t)buf * K * S2_ROWS;
      unsigned int mask = 0u;
      if (p.map) {
        for (int ob = 0; ob < K; ob += 7) {  // 28 loads in flight
          int v[7][4];
#pragma unroll
          for (int j = 0; j < 7; j++)
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const int o = ob + j, row = row0 + q * 32 + lane;
              v[j][q] = (o < K && row < p.Mout) ? __ldg(&p.map[(size_t)o * p.Mout + row]) : -1;
            }
#pragma unroll
          for (int j = 0; j < 7; j++) {
            const int o = ob + j;
            if (o < K) {
              bool any = false;
#pragma unroll
              for (int q = 0; q < 4; q++) {
                ms[o * S2_ROWS + q * 32 + lane] = v[j][q];
                any |= v[j][q] >= 0;
              }
              if (__any_sync(0xffffffffu, any)) mask |= 1u << o;
            }
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int row = row0 + q * 32 + lane;
          ms[q * 32 + lane] = row < p.Mout ? row : -1;
        }
        mask = 1u;
      }
      if (lane < K && (mask >> lane & 1u)) s_list[buf][__popc(mask & ((1u << lane) - 1u))] = lane;
      if (lane == 0) s_nact[buf] = __popc(mask);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_mapf[buf]));
    }
#ifdef SGB_SS_TIMELINE
    if (tl_on) p.dbg[9] = tl_w0;
#endif
  } else {
    // =========================== epilogue (warps 11-14; TMEM lane group = warp % 4) ============================
    TL_DECL;
    const int lg = warp & 3;
    int n = 0;
    for (int item = first; item < p.items; item += stride, n++) {
      const int ab = n & 1;
      const int tile = item / p.nparts, part = item % p.nparts;
      const int n0 = part * NT;
      const int nt = min(NT, p.N - n0);
      const int row = tile * S2_ROWS + lg * 32 + lane;
      mbar_wait(smem_u32(&bar_mapf[ab]), (uint32_t)((n >> 1) & 1));
      const bool has_acc = __shfl_sync(0xffffffffu, s_nact[ab], 0) > 0;
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_mape[ab]));
      TL_T0();
      mbar_wait(smem_u32(&bar_accf[ab]), (uint32_t)((n >> 1) & 1));
      TL_ADD(tl_w0);
      TL_T0();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // an item without any active pair (possible for the strided / inverse maps) never touched its accumulator: zero rows
      const uint32_t tbase = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ab * acc_cols);
      for (int cb = 0; cb < nt; cb += 8) {
        uint32_t v[8], c[8];
        if (has_acc) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                       : "r"(tbase + (uint32_t)cb));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7])
                       : "r"(tbase + (uint32_t)(nt + cb)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int e = 0; e < 8; e++) v[e] = c[e] = 0u;
        }
        if (cb + 8 >= nt) {  // last read of this accumulator: hand it back before the global traffic of the last columns
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_acce[ab]));
        }
        if (row < p.Mout) {
          const int col = n0 + cb;
          float x[8];
#pragma unroll
          for (int e = 0; e < 8; e++) x[e] = fmaf(__uint_as_float(c[e]), kSsLoInv, __uint_as_float(v[e]));
          const bool full = col + 8 <= p.Cout;
          if (p.bias) {
#pragma unroll
            for (int e = 0; e < 8; e++)
              if (col + e < p.Cout) x[e] += __ldg(&p.bias[col + e]);
          }
          if (p.residual) {
            const float *rp = p.residual + (size_t)row * p.res_stride + p.res_off + col;
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
          if (p.out) {
            float *op = p.out + (size_t)row * p.out_stride + p.out_off + col;
            if (full && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
              reinterpret_cast<float4 *>(op)[0] = make_float4(x[0], x[1], x[2], x[3]);
              reinterpret_cast<float4 *>(op)[1] = make_float4(x[4], x[5], x[6], x[7]);
            } else {
#pragma unroll
              for (int e = 0; e < 8; e++)
                if (col + e < p.Cout) op[e] = x[e];
            }
          }
          if (p.pk) {
            // the consumer's BatchNorm(eval)+ReLU, then the fp16 hi/lo split: 8 channels -> 4 words hi + 4 words lo
            float y[8];
            bool big = false;
#pragma unroll
            for (int e = 0; e < 8; e++) {
              float t = x[e];
              if (col + e >= p.Cout) t = 0.f;  // padding channels stay exactly zero
              else if (p.pk_scale) t = fmaf(t, __ldg(&p.pk_scale[col + e]), __ldg(&p.pk_shift[col + e]));
              if (p.pk_relu) t = fmaxf(t, 0.f);
              big |= !(fabsf(t) <= 65504.f);
              y[e] = t;
            }
            if (big) *p.oflow = 1;
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const __half2 h = f2h2_sat(y[2 * q], y[2 * q + 1]);
              const float2 hf = __half22float2(h);
              const __half2 l = f2h2_sat((y[2 * q] - hf.x) * kSsLoScale, (y[2 * q + 1] - hf.y) * kSsLoScale);
              hw[q] = *reinterpret_cast<const uint32_t *>(&h);
              lw[q] = *reinterpret_cast<const uint32_t *>(&l);
            }
            const int chn = p.pk_coff + col;  // channel in the packed tensor (multiple of 8)
            uint32_t *dst = p.pk + (size_t)row * p.pk_stride + (chn >> 5) * 32 + ((chn & 31) >> 1);
            *reinterpret_cast<uint4 *>(dst) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4 *>(dst + 16) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            if (cb + 8 == nt && n0 + nt == p.N && (p.N & 31) && p.pk_fill) {
              // N = Cout rounded to 16 ends in the middle of a 32-channel chunk: the consumer reads whole chunks, so the
              // upper half must be zero (not stale memory: 0 * NaN would poison the sums)
              const int ch2 = p.pk_coff + p.N;
              uint32_t *z = p.pk + (size_t)row * p.pk_stride + (ch2 >> 5) * 32 + ((ch2 & 31) >> 1);
              const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
              reinterpret_cast<uint4 *>(z)[0] = zero; reinterpret_cast<uint4 *>(z)[1] = zero;
              reinterpret_cast<uint4 *>(z + 16)[0] = zero; reinterpret_cast<uint4 *>(z + 16)[1] = zero;
            }
          }
        }
      }
      TL_ADD(tl_w1);
    }
#ifdef SGB_SS_TIMELINE
    if (tl_on && warp == 11) { p.dbg[10] = tl_w0; p.dbg[11] = tl_w1; }
#endif
  }
#ifdef SGB_SS_TIMELINE
  if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[12] = clock64();
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

#ifdef SGB_SS_TIMELINE
long long *g_ss_dbg = nullptr;
extern "C" void sgb_dev_ss_timeline(long long *d_buf) { g_ss_dbg = d_buf; }  // development build only (scripts/ss_timeline.py)
#endif

// Tile configuration of the persistent kernel: out[0] = NT, [1] = column parts, [2] = ring stages, [3] = work items,
// [4] = CTAs. Returns false when the shape is not handled (the caller falls back to spconv_tc_kernel).
bool spconv_ss_plan(int K, int Mout, int Cin, int Cout, int sms, int *out) {
  const int N = (Cout + 15) / 16 * 16;
  if (N > 256 || Cin > 512 || K > 27) return false;
  const int tiles = div_up(Mout, S2_ROWS);
  int NT = std::min(N, 128);
  if (N > 128) NT = (div_up(N, div_up(N, 128)) + 15) / 16 * 16;
  const int nparts = div_up(N, NT);
  const size_t stage = 2 * ((size_t)S2_A_BYTES + (size_t)NT * 128);  // a pair of iterations: rows + weights
  const size_t map_bytes = 2 * (size_t)K * S2_ROWS * 4;
  const size_t budget = 227 * 1024 - 2560;  // static shared memory (barriers, lists) + the 1 KB alignment slack
  int S = (int)((budget - map_bytes - 1024) / stage);
  S = std::max(2, std::min(S, S2_MAXS));
  out[0] = NT; out[1] = nparts; out[2] = S; out[3] = tiles * nparts; out[4] = std::min(tiles * nparts, sms);
  return true;
}

int spconv_ss_launch(const float *d_in_pk, int in_stride, const int32_t *d_map, int K, int Mout, const float *d_Wp, int Cin,
                     int Cout, const float *d_residual, int res_stride, int res_off, const float *d_bias, float *d_out,
                     int out_stride, int out_off, float *d_pk_out, int pk_stride, int pk_coff, const float *d_pk_scale,
                     const float *d_pk_shift, int pk_relu, int pk_fill, int *d_oflow, int sms, bool *attr_set, cudaStream_t stream) {
  int plan[5];
  SGB_REQUIRE(spconv_ss_plan(K, Mout, Cin, Cout, sms, plan), SGB_ERR_RANGE, "spconv_ss: shape not tiled");
  SsArgs p;
  p.in = (const uint32_t *)d_in_pk; p.in_stride = in_stride;
  p.map = d_map; p.K = K; p.Mout = Mout;
  p.Wp = d_Wp; p.Cin = Cin; p.N = (Cout + 15) / 16 * 16; p.Cout = Cout;
  p.NT = plan[0]; p.nparts = plan[1]; p.S = plan[2]; p.items = plan[3];
  p.residual = d_residual; p.res_stride = res_stride; p.res_off = res_off;
  p.bias = d_bias;
  p.out = d_out; p.out_stride = out_stride; p.out_off = out_off;
  p.pk = (uint32_t *)d_pk_out; p.pk_stride = pk_stride; p.pk_coff = pk_coff;
  p.pk_scale = d_pk_scale; p.pk_shift = d_pk_shift; p.pk_relu = pk_relu; p.pk_fill = pk_fill;
  p.oflow = d_oflow;
#ifdef SGB_SS_TIMELINE
  p.dbg = g_ss_dbg;
  p.dev_flags = getenv("SGB_SS_FLAGS") ? atoi(getenv("SGB_SS_FLAGS")) : 0;
  if (getenv("SGB_SS_S")) p.S = std::max(2, std::min(p.S, atoi(getenv("SGB_SS_S"))));
#else
  p.dbg = nullptr;
  p.dev_flags = 0;
#endif
  const size_t smem = (size_t)p.S * 2 * (S2_A_BYTES + (size_t)p.NT * 128) + 2 * (size_t)K * S2_ROWS * 4 + 1024;
  if (!*attr_set) {
    SGB_CUDA_CHECK(cudaFuncSetAttribute(spconv_ss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024 - 1536)));
    *attr_set = true;
  }
  spconv_ss_kernel<<<plan[4], S2_THREADS, smem, stream>>>(p);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}

}  // namespace sgb
