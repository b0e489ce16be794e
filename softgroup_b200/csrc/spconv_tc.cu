// spconv_tc.cu -- sparse convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// One CTA owns a tile of 128 output rows x N = Cout columns. For every kernel offset with at least one active
// pair in the tile and every 32-channel slice of Cin:
//   * the producer threads gather one input row slice each (eval-BatchNorm + ReLU folded in), split every value into
//     a TF32-exact high part and an fp32 remainder, and write both straight into TENSOR MEMORY with tcgen05.st
//     (A operand from TMEM: lane = row, column = channel) -- the shared-memory pipe only carries the weights;
//   * the pre-packed weight slice (same split, same core-matrix order, done once on the host side) is copied in;
//   * one thread issues the three error-compensated products hi*hi + hi*lo + lo*hi as
//     tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=Cout, K=8 per instruction) accumulating fp32 in TMEM, then
//     tcgen05.commit's the shared-memory stage back to the gather threads through an mbarrier.
// Stages form a ring so gathers of the next slice overlap the MMAs of the previous one. The epilogue reads the
// accumulator with tcgen05.ld (32 lanes x 32 bit x 16 columns per instruction), adds bias / residual and
// writes strided rows (the U-Net concat buffer). Dropping the lo*lo term and the TF32 truncation of the
// remainders leaves a relative error of ~2^-21 per product, i.e. fp32-grade (the north star's 1e-4 over ~40
// sequential convolutions rules out plain TF32).
#include <algorithm>

#include <cuda_fp16.h>

#include "common.cuh"

namespace sgb {

constexpr int TC_ROWS = 128;
constexpr int TC_KC = 32;  // channels per stage
constexpr int TC_THREADS = 320;  // 8 producer warps + 1 MMA warp + 1 weight-loader warp

struct TcArgs {
  const float *in; int in_stride, in_off;
  const int32_t *map; int K, Mout;
  const float *Wp;  // packed [K][nkc][8 chunks][2 (hi,lo)][N][4]
  int Cin, N, Cout;        // N = Cout rounded up to 16
  int NT;                  // columns per CTA (multiple of 16); gridDim.y = ceil(N / NT)
  const float *in_scale, *in_shift;
  const float *residual; int res_stride, res_off;
  const float *bias;
  float *out; int out_stride, out_off;
  int nstages, nbstages, tmem_cols, tmem_acols;  // A ring depth (TMEM), weight ring depth (smem), TMEM columns, first A column
  int in_packed;   // input rows are already activated + split: per 32-channel chunk [16 words hi pairs | 16 words lo pairs]
  long long *dbg;  // optional timeline buffer (test hook)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t *v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
      "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
      "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor: core matrix = 8 rows x 16 B (128 contiguous bytes);
// LBO = byte distance between the two 16-byte K chunks of one MMA, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ __half2 f2h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));  // upper half <- first source
  return *reinterpret_cast<__half2 *>(&r);
}
__device__ __forceinline__ float4 tf32_hi(float4 v) {
  float4 h;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  return h;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}

// Warp roles: warps 0-3 = producer group 0, warps 4-7 = producer group 1 (one output row per thread; the groups
// take alternate pipeline iterations), warp 8 = MMA issuer (one elected lane). Producers keep the NEXT iteration's
// gathered row slice in registers while the current one is being split and stored, so the L2 latency of the gather
// overlaps the stores, the weight copy (cp.async) and the other group's work. full[s]: 128 producer arrivals,
// free[s]: tcgen05.commit, done: accumulator complete.
__global__ void __launch_bounds__(TC_THREADS, 2) spconv_tc_kernel(TcArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[20];  // [0..3] A full, [4..7] A free, [8] done, [10..14] B full, [15..19] B free
  __shared__ uint32_t s_tmem;
  __shared__ unsigned int s_mask;
  __shared__ int s_list[32];
  __shared__ int s_nact;
  __shared__ __align__(16) float s_scale[512], s_shift[512];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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
  const int NS = p.nstages, NSB = p.nbstages;   // NS: A stages in TMEM (64 columns each), NSB: weight stages in smem
  unsigned char *bring = smem;
  int32_t *map_s = reinterpret_cast<int32_t *>(bring + (size_t)NSB * bstage_bytes);  // [K][128] (only when p.map)

  if (tid == 0) {
    for (int i = 0; i < NS; i++) {
      mbar_init(smem_u32(&bars[i]), TC_ROWS);
      mbar_init(smem_u32(&bars[4 + i]), 1);
    }
    mbar_init(smem_u32(&bars[8]), 1);
    for (int i = 0; i < NSB; i++) {
      mbar_init(smem_u32(&bars[10 + i]), 1);
      mbar_init(smem_u32(&bars[15 + i]), 1);
    }
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

  if (has_act) {
    for (int c = tid; c < 512; c += TC_THREADS) {
      s_scale[c] = (c < p.Cin) ? __ldg(&p.in_scale[c]) : 0.f;
      s_shift[c] = (c < p.Cin) ? __ldg(&p.in_shift[c]) : 0.f;
    }
  }
  // ---- rulebook slice of this tile -> shared memory; which kernel offsets have any active pair ------------
  if (p.map) {
    unsigned int flags = 0u;
    if (producer) {
      for (int o = grp; o < p.K; o += 2) {
        int src = row_ok ? __ldg(&p.map[(size_t)o * p.Mout + my_row]) : -1;
        map_s[o * TC_ROWS + r] = src;
        if (src >= 0) flags |= 1u << o;
      }
    }
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0 && flags) atomicOr(&s_mask, flags);
  }
  __syncthreads();
  if (tid == 0) {
    int n = 0;
    if (p.map) {
      unsigned int m = s_mask;
      for (int o = 0; o < p.K; o++)
        if (m >> o & 1u) s_list[n++] = o;
    } else {
      s_list[n++] = 0;
    }
    s_nact = n;
  }
  __syncthreads();
  const int nkc = (p.Cin + TC_KC - 1) / TC_KC;
  const int total = s_nact * nkc;

  if (producer && p.in_packed) {
    // ---- packed input (activated + split once by sgb_act_split): the gather is pure data movement, so the registers
    //      freed by the missing transform hold TWO future iterations of this thread's row (4 iterations ahead of the
    //      MMA warp counting both groups) -- the L2 latency of the gather is covered without shared memory.
    float4 va[8], vb[8];
    int ia = 0, ikc = grp;  // issue cursor: (offset list position, channel slice) of the next own iteration to load
    while (ikc >= nkc) { ikc -= nkc; ia++; }
    int nload = grp;        // global iteration index of the next load
    auto load = [&](float4 (&buf)[8]) {
      if (nload < total) {
        const int o = s_list[ia];
        const int src = p.map ? map_s[o * TC_ROWS + r] : (row_ok ? my_row : -1);
        if (src >= 0) {
          const float4 *rp = reinterpret_cast<const float4 *>(p.in + (size_t)src * p.in_stride + p.in_off + ikc * TC_KC);
#pragma unroll
          for (int q = 0; q < 8; q++) buf[q] = __ldg(rp + q);
        } else {
#pragma unroll
          for (int q = 0; q < 8; q++) buf[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      nload += 2;
      ikc += 2;
      while (ikc >= nkc) { ikc -= nkc; ia++; }
    };
    int s = grp % NS, u = grp / NS;
    int dbg_i = grp;
    auto consume = [&](float4 (&buf)[8]) {
      const bool dbg_on = p.dbg && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && r == 0 && dbg_i < 64;
      if (dbg_on) p.dbg[dbg_i * 8 + 0] = clock64();
      if (u >= 1) mbar_wait(smem_u32(&bars[4 + s]), (uint32_t)((u - 1) & 1));
      if (dbg_on) p.dbg[dbg_i * 8 + 1] = clock64();
      const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(p.tmem_acols + s * 32);
      uint32_t w[32];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        w[4 * q + 0] = __float_as_uint(buf[q].x); w[4 * q + 1] = __float_as_uint(buf[q].y);
        w[4 * q + 2] = __float_as_uint(buf[q].z); w[4 * q + 3] = __float_as_uint(buf[q].w);
      }
      tmem_st32(ta, w);
      const int s_done = s;
      s += 2;
      if (s >= NS) { s -= NS; u++; }
      if (dbg_on) p.dbg[dbg_i * 8 + 2] = clock64();
      load(buf);  // refill this buffer with own-iteration +2 while the store drains
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (dbg_on) p.dbg[dbg_i * 8 + 3] = clock64();
      mbar_arrive(smem_u32(&bars[s_done]));
      dbg_i += 2;
    };
    load(va);
    load(vb);
    for (int i = grp; i < total; i += 4) {
      consume(va);
      if (i + 2 < total) consume(vb);
    }
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
      const bool fast = (vsrc >= 0) && vec_ok && (kvalid == TC_KC || p.in_packed);
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
    // iteration i = grp, grp+2, ...: (a_idx, kc) = (i / nkc, i % nkc), stage s = i % NS, use u = i / NS -- all advanced
    // incrementally (no integer division in the loop)
    int i = grp;
    int kc = grp, a_idx = 0;
    while (kc >= nkc) { kc -= nkc; a_idx++; }
    int s = grp % NS, u = grp / NS;
    if (i < total) load_iter(a_idx, kc);
    for (; i < total; i += 2) {
      const bool dbg_on = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && r == 0 && i < 64;
      if (dbg_on) p.dbg[i * 8 + 0] = clock64();
      if (u >= 1) mbar_wait(smem_u32(&bars[4 + s]), (uint32_t)((u - 1) & 1));
      if (dbg_on) p.dbg[i * 8 + 1] = clock64();
      const int c0 = kc * TC_KC;
      const int kvalid = min(TC_KC, p.Cin - c0);
      // ---- A: registers -> (BN+ReLU) -> hi / lo -> tensor memory (lane = row, column = channel) -----------
      const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(p.tmem_acols + s * 32);
      {
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
          const __half2 l01 = f2h2_sat(x.x - f01.x, x.y - f01.y), l23 = f2h2_sat(x.z - f23.x, x.w - f23.y);
          hv[2 * q] = *reinterpret_cast<const uint32_t *>(&h01);
          hv[2 * q + 1] = *reinterpret_cast<const uint32_t *>(&h23);
          lv[2 * q] = *reinterpret_cast<const uint32_t *>(&l01);
          lv[2 * q + 1] = *reinterpret_cast<const uint32_t *>(&l23);
        }
        tmem_st16(ta, hv);
        tmem_st16(ta + 16u, lv);
      }
      const int s_done = s;
      kc += 2;
      while (kc >= nkc) { kc -= nkc; a_idx++; }
      s += 2;
      if (s >= NS) { s -= NS; u++; }
      if (dbg_on) p.dbg[i * 8 + 2] = clock64();
      if (i + 2 < total) load_iter(a_idx, kc);  // next gather of this group is in flight during the waits below
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (dbg_on) p.dbg[i * 8 + 3] = clock64();
      mbar_arrive(smem_u32(&bars[s_done]));
    }
  } else if (warp == 9) {
    // ---- weight loader: one elected lane streams the packed slices with TMA bulk copies (cp.async.bulk, async proxy:
    //      no generic->async fence needed) into a deeper ring; completion is signalled on the stage's mbarrier by
    //      complete_tx. Global layout [chunk q][hi|lo][N][16 B] == shared layout [q][hi nt | lo nt][16 B] when the CTA
    //      owns all N columns, so a whole slice is ONE bulk copy; with a column split it is one copy per (q, part).
    {
      int sb = 0, ub = 0, kc = 0, a_idx = 0;
      for (int i = 0; i < total; i++) {
        if (ub >= 1) mbar_wait(smem_u32(&bars[15 + sb]), (uint32_t)((ub - 1) & 1));
        const int o = s_list[a_idx];
        const int kvalid = min(TC_KC, p.Cin - kc * TC_KC);
        const int ksteps = (kvalid + 15) >> 4;  // K = 16 halves per tcgen05.mma
        const float4 *g = reinterpret_cast<const float4 *>(p.Wp) + ((size_t)o * nkc + kc) * (size_t)N * 8;
        const uint32_t bb = smem_u32(bring + (size_t)sb * bstage_bytes);
        const uint32_t bar = smem_u32(&bars[10 + sb]);
        const int nseg = 4 * ksteps;  // (chunk, part) segments of nt*16 bytes
        const uint32_t bytes = (uint32_t)nseg * (uint32_t)nt * 16u;
        if (lane == 0)
          asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
        __syncwarp();
        if (nt == N) {
          if (lane == 0)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(bb), "l"(g), "r"(bytes), "r"(bar) : "memory");
        } else if (lane < nseg) {  // column split: one bulk copy per (chunk, part) segment, one lane each
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(bb + (uint32_t)(lane * nt) * 16u), "l"(g + (size_t)lane * N + n0), "r"((uint32_t)nt * 16u), "r"(bar)
                       : "memory");
        }
        if (++sb == NSB) { sb = 0; ub++; }
        if (++kc == nkc) { kc = 0; a_idx++; }
      }
    }
  } else if (warp == 8 && lane == 0) {
    // ---- MMA issuer: per 8-channel k-step two instructions
    //        D[:, 0:2nt]  += A_hi * [B_hi | B_lo]     (N = 2nt)
    //        D[:, nt:2nt] += A_lo *  B_hi             (N = nt)
    //      so columns [0,nt) hold hi*hi and [nt,2nt) the two correction products (summed in the epilogue).
    //      Descriptors are advanced by adding constants to their low word; the issue loop is kept minimal because
    //      a single thread's instruction stream paces the tensor pipe (measured ~50 cycles / tcgen05.mma at N<=64).
    // instruction descriptor: D = f32 (1 << 4), A = B = f16 (format 0), both K-major, N >> 3, M >> 4
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * nt) >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
    const uint32_t idesc1 = (1u << 4) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
    const uint32_t b_lbo = (uint32_t)(2 * nt) * 16;
    const uint64_t b_step = (uint64_t)((2 * b_lbo) >> 4);
    uint32_t first = 0u;  // 0 for the very first MMA (overwrite), then 1
    int s = 0, sb = 0, kc = 0;
    uint32_t pa = 0u, pb = 0u;
    for (int i = 0; i < total; i++) {
      const bool dbg_on = p.dbg && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && i < 64;
      if (dbg_on) p.dbg[i * 8 + 4] = clock64();
      mbar_wait(smem_u32(&bars[10 + sb]), pb);
      if (dbg_on) p.dbg[i * 8 + 5] = clock64();
      mbar_wait(smem_u32(&bars[s]), pa);
      if (dbg_on) p.dbg[i * 8 + 6] = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int kvalid = min(TC_KC, p.Cin - kc * TC_KC);
      const int ksteps = (kvalid + 15) >> 4;
      const uint32_t sbt = smem_u32(bring + (size_t)sb * bstage_bytes);
      uint64_t dbb = umma_desc(sbt, b_lbo, 128);
      uint32_t ah = tmem + (uint32_t)(p.tmem_acols + s * 32), al = ah + 16u;
      for (int ks = 0; ks < ksteps; ks++) {
        umma_tf32_ts(tmem, ah, dbb, idesc2, (ks == 0) ? first : 1u);
        umma_tf32_ts(tmem + (uint32_t)nt, al, dbb, idesc1, 1u);
        ah += 8u; al += 8u; dbb += b_step;
      }
      first = 1u;
      umma_commit(smem_u32(&bars[4 + s]));    // frees the A stage once the MMAs above have read it
      umma_commit(smem_u32(&bars[15 + sb]));  // ... and the weight stage
      if (dbg_on) p.dbg[i * 8 + 7] = clock64();
      if (++s == NS) { s = 0; pa ^= 1u; }
      if (++sb == NSB) { sb = 0; pb ^= 1u; }
      if (++kc == nkc) kc = 0;
    }
    if (total > 0) umma_commit(smem_u32(&bars[8]));
  }
  // ---- epilogue: TMEM lane = output row; warp w reads lanes 32*(w%4).., column half w/4 -----------------
  if (producer) {
    if (total > 0) {
      mbar_wait(smem_u32(&bars[8]), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
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
        for (int e = 0; e < 8; e++) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(c[e]));
      } else {
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = 0u;
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
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols) : "memory");
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
  const __half2 l = f2h2_sat(a - hf.x, b - hf.y);
  const int chunk = c >> 5, w = (c & 31) >> 1;  // word index inside the chunk's hi block
  uint32_t *yr = y + (size_t)row * Cpad + chunk * 32;
  yr[w] = *reinterpret_cast<const uint32_t *>(&h);
  yr[16 + w] = *reinterpret_cast<const uint32_t *>(&l);
}
}  // namespace sgb

static long long *g_tc_dbg = nullptr;

extern "C" {

void sgb_test_set_tc_debug(long long *d_buf) { g_tc_dbg = d_buf; }

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
  // TMEM budget: accumulator [main | corrections] = 2*NT columns, then the A ring (64 columns per stage: hi + lo).
  // 256 columns (two CTAs per SM) when that leaves >= 2 A stages, else all 512.
  int dcols = 2 * NT;
  int acols0 = (dcols + 31) / 32 * 32;
  if (acols0 + 2 * 32 <= 256) { p.tmem_cols = 256; p.nstages = (256 - acols0) / 32; }
  else { p.tmem_cols = 512; p.nstages = (512 - acols0) / 32; }
  p.nstages = std::min(p.nstages, 4);
  p.tmem_acols = acols0;
  p.nbstages = (int)std::max<size_t>(2, std::min<size_t>(4, (96 * 1024 - map_bytes) / bstage));
  size_t smem = bstage * p.nbstages + map_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    SGB_CUDA_CHECK(cudaFuncSetAttribute(spconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    SGB_CUDA_CHECK(cudaFuncSetAttribute(spconv_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set = true;
  }
  dim3 grid(tiles, div_up(N, NT));
  spconv_tc_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(p);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}
}

// ---- micro-benchmark hook: cost of back-to-back tcgen05.mma kind::tf32 (M=128, N, K=8) into one accumulator ----
namespace sgb {
__global__ void umma_rate_kernel(int N, int reps, int per_commit, long long *out, int a_in_tmem) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 256 * 32 * 4) / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = 0.f;
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    uint64_t da = umma_desc(smem_u32(smem), 2048, 128);
    uint64_t db = umma_desc(smem_u32(smem + 16384), (uint32_t)N * 16, 128);
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int r = 0; r < reps; r += per_commit) {
      if (a_in_tmem > 0) {
        for (int k = 0; k < per_commit; k++) umma_tf32_ts(tmem + 64 * (k % a_in_tmem), tmem + 256 + 8 * (k & 7), db, idesc, 1u);
      } else {
        for (int k = 0; k < per_commit; k++) umma_tf32(tmem + 64 * (k % (-a_in_tmem + 1)), da, db, idesc, 1u);
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
    }
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
}  // namespace sgb

extern "C" int sgb_test_umma_rate(int N, int reps, int per_commit, long long *d_out, void *stream, int a_in_tmem) {
  SGB_CUDA_CHECK(cudaFuncSetAttribute(sgb::umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  sgb::umma_rate_kernel<<<1, 128, 16384 + 256 * 32 * 4, (cudaStream_t)stream>>>(N, reps, per_commit, d_out, a_in_tmem);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}
