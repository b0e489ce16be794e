// spconv_ss.cu -- sparse convolution, round-2 kernel: persistent CTAs, a deep cp.async gather ring, tcgen05 with both
// operands in shared memory, accumulators double-buffered in tensor memory, fused activation/split epilogue (sm_100a).
//
// Data: every activation tensor lives in HBM as PACKED rows -- per 32-channel chunk one 128-byte line
// [16 words of fp16 hi pairs | 16 words of fp16 lo pairs], x = hi + lo * 2^-kLoShift (error-compensated fp16 pair,
// fp32-grade products hi*hi + hi*lo + lo*hi accumulated in fp32). The convolution is output stationary:
// a work item is a tile of 128 output rows x NT output columns; for every kernel offset with an active pair in the
// tile and every 32-channel chunk of Cin ("iteration"):
//   * the 128 input row slices (128 B each, row indices from the rulebook slice in shared memory) are copied by
//     cp.async, 16 bytes per lane -- the 8 lanes of a quarter-warp fetch ONE whole line, a warp instruction 4 lines --
//     straight into the SWIZZLE_128B K-major tile tcgen05.mma reads (chunk position ^ (row & 7)); an absent neighbour
//     is a zero-fill copy (source size 0). Completion is tracked per stage by cp.async.mbarrier.arrive.noinc, so the
//     copies of S = 6..10 iterations are in flight per CTA without holding a register. Why: the round-1 kernel kept the
//     gathered rows in REGISTERS (two slices per producer thread), i.e. a prefetch distance of one iteration -- its
//     iteration time equalled the L2/DRAM latency (1457 cycles per pair measured) at a few per cent of the L1/L2 bandwidth.
//     TMA row gather (cp.async.bulk.tensor tile::gather4) was built and measured first: 20 cycles per 512-byte
//     instruction and SM (<= 31 B/clk/SM, scripts/tma_probe.cu; 2-3x slower per level, profiles/r2_*), so it is not used;
//   * the weight slice [B_hi | B_lo] (pre-split, pre-packed in core-matrix order by the host) arrives by
//     cp.async.bulk (TMA bulk copy) on the same stage barrier;
//   * one elected lane issues  D[:, 0:2nt] += A_hi [B_hi | B_lo],  D[:, nt:2nt] += A_lo B_hi  per 16-channel k-step
//     (tcgen05.mma.cta_group::1.kind::f16, M = 128, A and B from shared-memory descriptors) and commits the stage back.
// The CTA is persistent (grid = #SMs) and walks work items round-robin. Accumulators are double-buffered in TMEM: the
// epilogue warps drain tile j (tcgen05.ld -> + bias + residual -> optional fp32 rows, optional NEXT layer's
// BatchNorm+ReLU -> fp16 hi/lo split -> packed rows) while the pipeline already runs tile j+1; the rulebook slice of
// tile j+1 is prefetched into registers during tile j's gathers.
// Warp roles (448 threads): 0-7 rulebook + gather, 8 MMA issue, 9 weight loader, 10-13 epilogue.
#include <algorithm>

#include <cuda_fp16.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace sgb {

#ifdef SGB_SS_TIMELINE
#define TL_DECL long long tl_acc = 0, tl_t0 = 0; const bool tl_on = p.dbg && blockIdx.x == 0 && lane == 0
#define TL_BEGIN() do { if (tl_on) tl_t0 = clock64(); } while (0)
#define TL_END() do { if (tl_on) tl_acc += clock64() - tl_t0; } while (0)
#define TL_STORE(slot) do { if (tl_on) p.dbg[slot] = tl_acc; } while (0)
#else
#define TL_DECL
#define TL_BEGIN()
#define TL_END()
#define TL_STORE(slot)
#endif

constexpr int T2_ROWS = 128;
constexpr int T2_KC = 32;        // channels per iteration (one 128-byte packed line per row)
constexpr int T2_THREADS = 448;
constexpr int T2_PROD = 256;    // gather threads (8 warps)
constexpr int T2_MAXS = 12;      // ring depth limit (barrier arrays)
constexpr int T2_A_BYTES = T2_ROWS * 128;
constexpr int kLoShift2 = 11;    // lo = fp16((x - hi) * 2^11): no fp16 subnormals for |x| >= 2^-14
constexpr float kLoScale2 = (float)(1 << kLoShift2), kLoInv2 = 1.0f / kLoScale2;

struct Tc2Args {
  const uint32_t *in; int in_stride;     // packed input rows [Min][in_stride words]
  const int32_t *map; int K, Mout, Min;  // map [K][Mout] (nullptr: identity, K == 1)
  const float *Wp;                       // packed weights [K][nkc][4][2][N][8 halves]
  int Cin, N, Cout, NT, nparts;          // N = Cout rounded up to 16; column parts of NT (last may be shorter)
  const float *residual; int res_stride, res_off;
  const float *bias;
  float *out; int out_stride, out_off;   // optional fp32 rows
  uint32_t *pk; int pk_stride, pk_coff;  // optional packed rows: row stride in words, first channel (multiple of 8)
  const float *pk_scale, *pk_shift; int pk_relu;  // per OUTPUT channel of this conv (nullptr: identity)
  int pk_fill;                           // zero the rest of a half-written last 32-channel chunk
  int *oflow;                            // device flag: |value| > 65504 met while packing
  int S;                                 // ring stages
  int tiles, items;
  long long *dbg;                        // SGB_SS_TIMELINE builds only: per-role wait/busy cycle counters of CTA 0
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
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Shared-memory carve-up (dynamic, 1024-byte aligned): [A ring: S x 16 KB][B ring: S x NT*128 B][map: 2 x K x 128 int32]
__global__ void __launch_bounds__(T2_THREADS, 1) spconv_ss_kernel(Tc2Args p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar_full[T2_MAXS], bar_empty[T2_MAXS];
  __shared__ __align__(8) unsigned long long bar_accf[2], bar_acce[2], bar_mapf[2], bar_mape[2];
  __shared__ uint32_t s_tmem;
  __shared__ unsigned int s_mask[2];
  __shared__ int s_list[2][32];
  __shared__ int s_nact[2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int S = p.S, K = p.K, NT = p.NT;
  const uint32_t b_stage = (uint32_t)NT * 128u;  // [hi | lo] x 4 chunks x NT x 16 B
  // SWIZZLE_128B tiles need 1024-byte alignment: align at run time (the launch adds 1 KB of slack)
  unsigned char *a_ring = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  unsigned char *b_ring = a_ring + (size_t)S * T2_A_BYTES;
  int32_t *map_s = reinterpret_cast<int32_t *>(b_ring + (size_t)S * b_stage);  // [2][K][128]
  const int nkc = (p.Cin + T2_KC - 1) / T2_KC;

  if (tid == 0) {
    for (int s = 0; s < S; s++) {
      mbar_init(smem_u32(&bar_full[s]), T2_PROD + 1);  // every gather thread (cp.async ... arrive.noinc) + the weight loader
      mbar_init(smem_u32(&bar_empty[s]), 1);  // tcgen05.commit
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(smem_u32(&bar_accf[b]), 1);    // tcgen05.commit (or the MMA lane for an empty tile)
      mbar_init(smem_u32(&bar_acce[b]), 4);    // one arrival per epilogue warp
      mbar_init(smem_u32(&bar_mapf[b]), 1);    // gather thread 0
      mbar_init(smem_u32(&bar_mape[b]), 6);    // MMA warp, weight loader and the 4 epilogue warps have read the tile's list
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_mask[0] = s_mask[1] = 0u;
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const int acc_cols = (2 * NT + 31) / 32 * 32;  // columns of one accumulator buffer (<= 256)

  const int first = blockIdx.x, stride = gridDim.x;

  if (warp < 8) {
    // =========================== rulebook slices + cp.async gather ===========================================
    const int r = tid & (T2_ROWS - 1);  // tile row this thread reads the rulebook for (two threads per row: even / odd offsets)
    const int half = tid >> 7;
    int mreg[14];  // rulebook entries of row r for the offsets half, half + 2, ...
    auto load_map = [&](int item) {  // global -> registers (absent / out of range: -1)
      const int tile = item / p.nparts;
      const int row = tile * T2_ROWS + r;
      const bool ok = row < p.Mout;
      if (p.map) {
#pragma unroll
        for (int j = 0; j < 14; j++) {
          const int o = half + 2 * j;
          mreg[j] = (o < K && ok) ? __ldg(&p.map[(size_t)o * p.Mout + row]) : -1;
        }
      } else {
        mreg[0] = (ok && half == 0) ? row : -1;
      }
    };
    auto publish_map = [&](int buf, int n) {  // registers -> shared memory, active-offset list, signal
      if (n >= 2) mbar_wait(smem_u32(&bar_mape[buf]), (uint32_t)(((n >> 1) - 1) & 1));
      int32_t *ms = map_s + (size_t)buf * K * T2_ROWS;
      unsigned int flags = 0u;
#pragma unroll
      for (int j = 0; j < 14; j++) {
        const int o = half + 2 * j;
        if (o < K) {
          ms[o * T2_ROWS + r] = mreg[j];
          if (mreg[j] >= 0) flags |= 1u << o;
        }
      }
      flags = __reduce_or_sync(0xffffffffu, flags);
      if (lane == 0 && flags) atomicOr(&s_mask[buf], flags);
      named_bar_sync(1, T2_PROD);
      if (tid < 32) {
        const unsigned int m = s_mask[buf];
        if (tid < K && (m >> tid & 1u)) s_list[buf][__popc(m & ((1u << tid) - 1u))] = tid;
        if (tid == 0) s_nact[buf] = __popc(m);
      }
      named_bar_sync(1, T2_PROD);
      if (tid == 0) {
        s_mask[buf] = 0u;
        mbar_arrive(smem_u32(&bar_mapf[buf]));
      }
    };
    int n = 0;        // local tile counter
    int g = 0;        // global iteration counter of this CTA (ring position)
    TL_DECL;
#ifdef SGB_SS_TIMELINE
    long long tl_pub = 0, tl_start = clock64();
#endif
    const uint32_t a_base_u = smem_u32(a_ring);
    // copy geometry of this thread: warp w covers tile rows 16 w .. 16 w + 15 in four instructions of 4 rows x 8 chunks
    const int sub = lane >> 3, ch = lane & 7;
    if (first < p.items) {
      load_map(first);
      publish_map(0, 0);
    }
    for (int item = first; item < p.items; item += stride, n++) {
      const int buf = n & 1;
      const bool has_next = item + stride < p.items;
      if (has_next) load_map(item + stride);  // in flight during this tile's gathers
      const int nact = s_nact[buf];
      const int32_t *ms = map_s + (size_t)buf * K * T2_ROWS;
      const int total = nact * nkc;
      int a = 0, kc = 0;
      for (int i = 0; i < total; i++, g++) {
        const int s = g % S, u = g / S;
        TL_BEGIN();
        if (u >= 1) mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)((u - 1) & 1));
        TL_END();
        const uint32_t bar = smem_u32(&bar_full[s]);
        const int o = s_list[buf][a];
        const uint32_t stage = a_base_u + (uint32_t)s * T2_A_BYTES;
        const uint32_t *gcol = p.in + kc * T2_KC + ch * 4;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int row = warp * 16 + j * 4 + sub;
          const int src = ms[o * T2_ROWS + row];  // the 8 lanes of a quarter-warp read the same entry (broadcast)
          const uint32_t dst = stage + (uint32_t)row * 128u + (uint32_t)((ch ^ (row & 7)) << 4);
          const uint32_t *gp = gcol + (size_t)max(src, 0) * p.in_stride;
          const int nbytes = (src >= 0) ? 16 : 0;  // 0 source bytes = zero fill (absent neighbour)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gp), "r"(nbytes) : "memory");
        }
        // the barrier's pending count was initialised with this arrival: it fires when the copies above have landed
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
        if (++kc == nkc) { kc = 0; a++; }
      }
#ifdef SGB_SS_TIMELINE
      const long long tp0 = clock64();
#endif
      if (has_next) publish_map(buf ^ 1, n + 1);
#ifdef SGB_SS_TIMELINE
      tl_pub += clock64() - tp0;
#endif
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
#ifdef SGB_SS_TIMELINE
    if (tl_on && warp == 0) { p.dbg[0] = clock64() - tl_start; p.dbg[1] = tl_acc; p.dbg[2] = tl_pub; p.dbg[9] = g; p.dbg[10] = n; }
#endif
  } else if (warp_u == 8) {
    // =========================== MMA issue ====================================================================
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t a_base = smem_u32(a_ring), b_base = smem_u32(b_ring);
    int n = 0, g = 0;
    TL_DECL;
#ifdef SGB_SS_TIMELINE
    long long tl_acce = 0, tl_mapf = 0;
#endif
    for (int item = first; item < p.items; item += stride, n++) {
      const int buf = n & 1, ab = n & 1;
      const int part = item % p.nparts;
      const int nt = min(NT, p.N - part * NT);
#ifdef SGB_SS_TIMELINE
      long long tq = clock64();
#endif
      mbar_wait(smem_u32(&bar_mapf[buf]), (uint32_t)((n >> 1) & 1));
#ifdef SGB_SS_TIMELINE
      tl_mapf += clock64() - tq;
#endif
      const int total = __shfl_sync(0xffffffffu, s_nact[buf], 0) * nkc;
      __syncwarp();
      if (leader) mbar_arrive(smem_u32(&bar_mape[buf]));
#ifdef SGB_SS_TIMELINE
      tq = clock64();
#endif
      if (n >= 2) mbar_wait(smem_u32(&bar_acce[ab]), (uint32_t)(((n >> 1) - 1) & 1));
#ifdef SGB_SS_TIMELINE
      tl_acce += clock64() - tq;
#endif
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d = tmem + (uint32_t)(ab * acc_cols);
      const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * nt) >> 3) << 17) | ((uint32_t)(T2_ROWS >> 4) << 24);
      const uint32_t idesc1 = (1u << 4) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(T2_ROWS >> 4) << 24);
      const uint32_t b_lbo = (uint32_t)(2 * nt) * 16u;
      const uint64_t b_step = (uint64_t)((2 * b_lbo) >> 4);
      uint32_t acc = 0u;
      int kc = 0;
      for (int i = 0; i < total; i++, g++) {
        const int s = g % S;
        TL_BEGIN();
        mbar_wait(smem_u32(&bar_full[s]), (uint32_t)((g / S) & 1));
        TL_END();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // cp.async wrote the A stage through the generic proxy
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int ksteps = (min(T2_KC, p.Cin - kc * T2_KC) + 15) >> 4;
        uint64_t ad = desc_sw128(a_base + (uint32_t)s * T2_A_BYTES);
        uint64_t bd = umma_desc(b_base + (uint32_t)s * b_stage, b_lbo, 128);
        for (int ks = 0; ks < ksteps; ks++) {
          if (leader) {
            umma_f16_ss(d, ad, bd, idesc2, acc);                      // A_hi [B_hi | B_lo]
            umma_f16_ss(d + (uint32_t)nt, ad + 4u, bd, idesc1, 1u);   // A_lo B_hi   (+64 B on the A start address)
          }
          acc = 1u;
          ad += 2u;  // +32 B: next 16 channels inside the 128-byte swizzle atom
          bd += b_step;
        }
        if (leader) umma_commit(smem_u32(&bar_empty[s]));
        __syncwarp();
        if (++kc == nkc) kc = 0;
      }
      if (leader) {
        if (total > 0) umma_commit(smem_u32(&bar_accf[ab]));
        else mbar_arrive(smem_u32(&bar_accf[ab]));
      }
      __syncwarp();
    }
#ifdef SGB_SS_TIMELINE
    if (tl_on) { p.dbg[3] = tl_acc; p.dbg[4] = tl_acce; p.dbg[5] = tl_mapf; }
#endif
  } else if (warp == 9) {
    // =========================== weight loader ================================================================
    int n = 0, g = 0;
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
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
      int a = 0, kc = 0;
      for (int i = 0; i < total; i++, g++) {
        const int s = g % S, u = g / S;
        const int o = __shfl_sync(0xffffffffu, my_o, a);
        if (u >= 1) mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)((u - 1) & 1));
        const uint32_t bar = smem_u32(&bar_full[s]);
        const int ks = (min(T2_KC, p.Cin - kc * T2_KC) + 15) >> 4;
        const int nseg = 4 * ks;  // (chunk, hi|lo) segments of nt * 16 bytes
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)nseg * (uint32_t)nt * 16u);
        __syncwarp();
        const float4 *gsrc = reinterpret_cast<const float4 *>(p.Wp) + ((size_t)o * nkc + kc) * (size_t)p.N * 8;
        const uint32_t dst = __shfl_sync(0xffffffffu, smem_u32(b_ring), 0) + (uint32_t)s * b_stage;
        // uniform operands + one elected lane: no per-lane issue waterfall (see the gather warps)
        if (nt == p.N) {
          if (leader)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(gsrc), "r"((uint32_t)nseg * (uint32_t)nt * 16u), "r"(bar) : "memory");
        } else {
          for (int sg = 0; sg < nseg; sg++)
            if (leader)
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                           ::"r"(dst + (uint32_t)(sg * nt) * 16u), "l"(gsrc + (size_t)sg * p.N + n0), "r"((uint32_t)nt * 16u), "r"(bar)
                           : "memory");
        }
        if (++kc == nkc) { kc = 0; a++; }
      }
    }
  } else {
    // =========================== epilogue (warps 10-13; TMEM lane group = warp % 4) ============================
    const int lg = warp & 3;
    int n = 0;
    TL_DECL;
#ifdef SGB_SS_TIMELINE
    long long tl_busy = 0;
#endif
    for (int item = first; item < p.items; item += stride, n++) {
      const int ab = n & 1;
      const int tile = item / p.nparts, part = item % p.nparts;
      const int n0 = part * NT;
      const int nt = min(NT, p.N - n0);
      const int row = tile * T2_ROWS + lg * 32 + lane;
      TL_BEGIN();
      mbar_wait(smem_u32(&bar_accf[ab]), (uint32_t)((n >> 1) & 1));
      TL_END();
#ifdef SGB_SS_TIMELINE
      const long long te0 = clock64();
#endif
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // a tile without any active pair (possible for the strided / inverse maps) never touched its accumulator: zero rows
      const uint32_t tbase = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ab * acc_cols);
      const bool has_acc = __shfl_sync(0xffffffffu, s_nact[n & 1], 0) > 0;
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_mape[n & 1]));
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
          for (int e = 0; e < 8; e++) x[e] = fmaf(__uint_as_float(c[e]), kLoInv2, __uint_as_float(v[e]));
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
              const __half2 l = f2h2_sat((y[2 * q] - hf.x) * kLoScale2, (y[2 * q + 1] - hf.y) * kLoScale2);
              hw[q] = *reinterpret_cast<const uint32_t *>(&h);
              lw[q] = *reinterpret_cast<const uint32_t *>(&l);
            }
            const int ch = p.pk_coff + col;  // channel in the packed tensor (multiple of 8)
            uint32_t *dst = p.pk + (size_t)row * p.pk_stride + (ch >> 5) * 32 + ((ch & 31) >> 1);
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
#ifdef SGB_SS_TIMELINE
      tl_busy += clock64() - te0;
#endif
    }
#ifdef SGB_SS_TIMELINE
    if (tl_on && warp == 10) { p.dbg[7] = tl_acc; p.dbg[8] = tl_busy; }
#endif
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// y = BatchNorm(eval)+ReLU(x) (or x when scale == nullptr) -> packed rows (see the header of this file). One thread per
// (row, chunk, word pair); channels past C are zero. Used where a tensor has no producing convolution to fuse into
// (network input, concat halves written by different producers with one BatchNorm over both, gathered point rows).
__global__ void act_pack_kernel(const float *__restrict__ x, int x_stride, int x_off, const float *__restrict__ scale,
                                const float *__restrict__ shift, int relu, uint32_t *__restrict__ y, int y_stride, int y_coff,
                                int M, int C, int Cpad, int *__restrict__ oflow) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int wpr = Cpad >> 1;  // word pairs (2 channels) per row
  if (t >= (long long)M * wpr) return;
  const int row = (int)(t / wpr), pr = (int)(t % wpr);
  const int c = 2 * pr;
  float a = 0.f, b = 0.f;
  if (c < C) a = x[(size_t)row * x_stride + x_off + c];
  if (c + 1 < C) b = x[(size_t)row * x_stride + x_off + c + 1];
  if (scale) {
    if (c < C) a = fmaf(a, __ldg(&scale[c]), __ldg(&shift[c]));
    if (c + 1 < C) b = fmaf(b, __ldg(&scale[c + 1]), __ldg(&shift[c + 1]));
  }
  if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
  if (!(fabsf(a) <= 65504.f) || !(fabsf(b) <= 65504.f)) *oflow = 1;
  const __half2 h = f2h2_sat(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = f2h2_sat((a - hf.x) * kLoScale2, (b - hf.y) * kLoScale2);
  const int ch = y_coff + c;
  uint32_t *yr = y + (size_t)row * y_stride + (ch >> 5) * 32 + ((ch & 31) >> 1);
  yr[0] = *reinterpret_cast<const uint32_t *>(&h);
  yr[16] = *reinterpret_cast<const uint32_t *>(&l);
}

static long long *g_ss_dbg = nullptr;  // only ever set by the SGB_SS_TIMELINE development build
static int *g_oflow = nullptr;  // device flag shared by every launch of this process (per current device at first use)

}  // namespace sgb

using namespace sgb;

extern "C" {

#ifdef SGB_SS_TIMELINE
void sgb_dev_ss_timeline(long long *d_buf) { g_ss_dbg = d_buf; }
#endif

int sgb_spconv_lo_shift(void) { return kLoShift2; }

// Reads (and clears) the overflow flag raised by the packing code paths: blocking 4-byte read on `stream`.
int sgb_spconv_overflow(int *h_flag, void *stream) {
  SGB_REQUIRE(h_flag, SGB_ERR_ARG, "null flag");
  *h_flag = 0;
  if (!g_oflow) return SGB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  SGB_CUDA_CHECK(cudaMemcpyAsync(h_flag, g_oflow, 4, cudaMemcpyDeviceToHost, st));
  SGB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (*h_flag) SGB_CUDA_CHECK(cudaMemsetAsync(g_oflow, 0, 4, st));
  return SGB_OK;
}

static int ensure_oflow() {
  if (!g_oflow) {
    SGB_CUDA_CHECK(cudaMalloc(&g_oflow, 4));
    SGB_CUDA_CHECK(cudaMemset(g_oflow, 0, 4));
  }
  return SGB_OK;
}

int sgb_act_pack(const float *d_x, int x_stride, int x_off, const float *d_scale, const float *d_shift, int relu,
                 float *d_pk, int pk_stride, int pk_coff, int M, int C, int Cfill, void *stream) {
  if (M == 0 || C == 0) return SGB_OK;
  SGB_REQUIRE(d_x && d_pk && M > 0 && C > 0 && (d_scale == nullptr) == (d_shift == nullptr), SGB_ERR_ARG, "act_pack arguments");
  SGB_REQUIRE((pk_stride & 31) == 0 && (pk_coff & 1) == 0 && (Cfill & 1) == 0 && Cfill >= C && pk_stride >= pk_coff + Cfill,
              SGB_ERR_ARG, "act_pack: row stride multiple of 32 words, channel offset and fill width even, fill inside the row");
  int rc = ensure_oflow();
  if (rc) return rc;
  long long tot = (long long)M * (Cfill / 2);
  act_pack_kernel<<<div_up(tot, 256), 256, 0, (cudaStream_t)stream>>>(d_x, x_stride, x_off, d_scale, d_shift, relu,
                                                                    (uint32_t *)d_pk, pk_stride, pk_coff, M, C, Cfill, g_oflow);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}

int sgb_spconv_forward_ss(const float *d_in_pk, int in_stride, int Min, const int32_t *d_map, int K, int Mout,
                           const float *d_Wp, int Cin, int Cout, const float *d_residual, int res_stride, int res_off,
                           const float *d_bias, float *d_out, int out_stride, int out_off, float *d_pk_out, int pk_stride,
                           int pk_coff, const float *d_pk_scale, const float *d_pk_shift, int pk_relu, int pk_fill, void *stream) {
  if (Mout == 0 || Cout == 0) return SGB_OK;
  SGB_REQUIRE(d_in_pk && d_Wp && (d_out || d_pk_out) && K >= 1 && K <= 27 && Mout > 0 && Min > 0 && Cin > 0 && Cout > 0, SGB_ERR_ARG,
              "spconv_forward_ss arguments");
  SGB_REQUIRE(d_map || (K == 1 && Min >= Mout), SGB_ERR_ARG, "identity map requires K == 1");
  SGB_REQUIRE((in_stride & 31) == 0 && in_stride >= (Cin + 31) / 32 * 32, SGB_ERR_ARG, "packed input row stride (words, multiple of 32)");
  SGB_REQUIRE((((uintptr_t)d_in_pk) & 15) == 0, SGB_ERR_ARG, "packed input must be 16-byte aligned");
  SGB_REQUIRE(!d_out || out_stride >= out_off + Cout, SGB_ERR_ARG, "fp32 output row stride");
  SGB_REQUIRE(!d_pk_out || ((pk_stride & 31) == 0 && (pk_coff & 7) == 0 && pk_stride >= pk_coff + (Cout + 15) / 16 * 16), SGB_ERR_ARG,
              "packed output: row stride multiple of 32 words, channel offset multiple of 8");
  SGB_REQUIRE((d_pk_scale == nullptr) == (d_pk_shift == nullptr), SGB_ERR_ARG, "scale/shift must come together");
  const int N = (Cout + 15) / 16 * 16;
  SGB_REQUIRE(N <= 256 && Cin <= 512, SGB_ERR_RANGE, "spconv_forward_ss: Cout > 256 or Cin > 512 is not tiled");
  int rc = ensure_oflow();
  if (rc) return rc;
  Tc2Args p;
  p.in = (const uint32_t *)d_in_pk; p.in_stride = in_stride;
  p.map = d_map; p.K = K; p.Mout = Mout; p.Min = Min;
  p.Wp = d_Wp; p.Cin = Cin; p.N = N; p.Cout = Cout;
  p.residual = d_residual; p.res_stride = res_stride; p.res_off = res_off;
  p.bias = d_bias;
  p.out = d_out; p.out_stride = out_stride; p.out_off = out_off;
  p.pk = (uint32_t *)d_pk_out; p.pk_stride = pk_stride; p.pk_coff = pk_coff;
  p.pk_scale = d_pk_scale; p.pk_shift = d_pk_shift; p.pk_relu = pk_relu; p.pk_fill = pk_fill;
  p.oflow = g_oflow;
  p.dbg = g_ss_dbg;
  // Column parts: [B_hi | B_lo] is one operand of 2*NT <= 256 columns; few row tiles (deep levels) are cut further so the
  // work items cover the SMs. Every extra part re-gathers the tile's input rows, so stop at one item per SM.
  int sms = kNumSMs;
  const int tiles = div_up(Mout, T2_ROWS);
  int NT = std::min(N, 128);
  if (N > 128) NT = (div_up(N, div_up(N, 128)) + 15) / 16 * 16;
  while (NT > 16 && tiles * div_up(N, NT) < sms) {
    int nxt = (NT / 2 + 15) / 16 * 16;
    if (nxt >= NT) break;
    NT = nxt;
  }
  p.NT = NT;
  p.nparts = div_up(N, NT);
  p.tiles = tiles;
  p.items = tiles * p.nparts;
  const size_t b_stage = (size_t)NT * 128;
  const size_t map_bytes = 2 * (size_t)K * T2_ROWS * 4;
  const size_t budget = 216 * 1024;  // + ~1.2 KB of static shared memory (barriers, lists) must stay within 227 KB per CTA
  int S = (int)((budget - map_bytes - 1024) / (T2_A_BYTES + b_stage));
  S = std::max(2, std::min(S, T2_MAXS));
  p.S = S;
  const size_t smem = (size_t)S * (T2_A_BYTES + b_stage) + map_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    SGB_CUDA_CHECK(cudaFuncSetAttribute(spconv_ss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    attr_set = true;
  }
  const int grid = std::min(p.items, sms);
  spconv_ss_kernel<<<grid, T2_THREADS, smem, (cudaStream_t)stream>>>(p);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}
}
