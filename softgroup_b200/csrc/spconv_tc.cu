// spconv_tc.cu -- sparse convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// One CTA owns a tile of 128 output rows x N = Cout columns. For every kernel offset with at least one active
// pair in the tile and every 32-channel slice of Cin:
//   * the 128 threads gather one input row slice each (eval-BatchNorm + ReLU folded in), split every value into
//     a TF32-exact high part and an fp32 remainder, and store both as UMMA "K-major, no swizzle" core matrices
//     (8 rows x 16 B, contiguous 128 B) in shared memory;
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

#include "common.cuh"

namespace sgb {

constexpr int TC_ROWS = 128;
constexpr int TC_KC = 32;  // channels per stage
constexpr int TC_THREADS = 256;

struct TcArgs {
  const float *in; int in_stride, in_off;
  const int32_t *map; int K, Mout;
  const float *Whi, *Wlo;  // packed [K][nkc][8][N][4]
  int Cin, N, Cout;        // N = Cout rounded up to 16
  int NT;                  // columns per CTA (multiple of 16); gridDim.y = ceil(N / NT)
  const float *in_scale, *in_shift;
  const float *residual; int res_stride, res_off;
  const float *bias;
  float *out; int out_stride, out_off;
  int nstages, tmem_cols;
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
// K-major, SWIZZLE_NONE shared-memory matrix descriptor: core matrix = 8 rows x 16 B (128 contiguous bytes);
// LBO = byte distance between the two 16-byte K chunks of one MMA, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float4 tf32_hi(float4 v) {
  float4 h;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  return h;
}

__global__ void __launch_bounds__(TC_THREADS) spconv_tc_kernel(TcArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[8];  // [0..3] stage free, [4] accumulator done
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & (TC_ROWS - 1), half = tid >> 7;  // two threads per row: each gathers 16 of the 32 channels
  const int row0 = blockIdx.x * TC_ROWS;
  const int my_row = row0 + r;
  const bool row_ok = my_row < p.Mout;
  const bool has_act = p.in_scale != nullptr;
  const bool vec_ok = ((p.in_stride & 3) == 0) && ((p.in_off & 3) == 0) && ((((uintptr_t)p.in) & 15) == 0);
  const int N = p.N, NT = p.NT;
  const int n0 = blockIdx.y * NT;
  const int nt = min(NT, N - n0);                       // columns of this CTA (multiple of 16)
  const uint32_t a_bytes = TC_ROWS * TC_KC * 4;         // 16 KB
  const uint32_t b_bytes = (uint32_t)NT * TC_KC * 4;    // NT * 128 B
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  const int NS = p.nstages;

  if (tid == 0) {
    for (int i = 0; i < NS; i++) mbar_init(smem_u32(&bars[i]), 1);
    mbar_init(smem_u32(&bars[4]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
  // instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3, M>>4
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);

  const int nkc = (p.Cin + TC_KC - 1) / TC_KC;
  int it = 0;
  int src_next = -1;
  if (row_ok) src_next = p.map ? __ldg(&p.map[my_row]) : my_row;
  for (int o = 0; o < p.K; o++) {
    const int src = src_next;
    if (o + 1 < p.K && row_ok) src_next = __ldg(&p.map[(size_t)(o + 1) * p.Mout + my_row]);  // prefetch
    if (!__syncthreads_or(src >= 0)) continue;
    for (int kc = 0; kc < nkc; kc++, it++) {
      const int s = it % NS;
      if (it >= NS) mbar_wait(smem_u32(&bars[s]), (uint32_t)((it / NS - 1) & 1));
      unsigned char *st = smem + (size_t)s * stage_bytes;
      float4 *Ahi = reinterpret_cast<float4 *>(st);
      float4 *Alo = reinterpret_cast<float4 *>(st + a_bytes);
      float4 *Bhi = reinterpret_cast<float4 *>(st + 2 * a_bytes);
      float4 *Blo = reinterpret_cast<float4 *>(st + 2 * a_bytes + b_bytes);
      const int c0 = kc * TC_KC;
      const int kvalid = min(TC_KC, p.Cin - c0);  // channels of this slice that exist
      const int ksteps = (kvalid + 7) >> 3;
      // ---- B: packed weight slice (already split, core-matrix order) via cp.async -----------------------
      {
        const size_t slice = ((size_t)o * nkc + kc) * (size_t)N * 8;  // float4 units: 8 chunks x N rows
        const float4 *gh = reinterpret_cast<const float4 *>(p.Whi) + slice + n0;
        const float4 *gl = reinterpret_cast<const float4 *>(p.Wlo) + slice + n0;
        const int nvec = 2 * ksteps * nt;
        const uint32_t bh = smem_u32(Bhi), bl = smem_u32(Blo);
        for (int t = tid; t < nvec; t += TC_THREADS) {
          int q = t / nt, n = t - q * nt;
          cp_async16(bh + (uint32_t)(q * nt + n) * 16, gh + (size_t)q * N + n);
          cp_async16(bl + (uint32_t)(q * nt + n) * 16, gl + (size_t)q * N + n);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      // ---- A: gathered row slice (this thread: chunks half*4 .. half*4+3) -> hi / lo core matrices -----
      {
        float4 v[4];
        const float *rp = (src >= 0) ? p.in + (size_t)src * p.in_stride + p.in_off + c0 : nullptr;
        const bool fast = (src >= 0) && vec_ok && (kvalid == TC_KC);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int q = half * 4 + j;
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (fast) {
            v[j] = __ldg(reinterpret_cast<const float4 *>(rp) + q);
          } else if (src >= 0 && 4 * q < kvalid) {
            float t[4];
#pragma unroll
            for (int e = 0; e < 4; e++) t[e] = (4 * q + e < kvalid) ? __ldg(rp + 4 * q + e) : 0.f;
            v[j] = make_float4(t[0], t[1], t[2], t[3]);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int q = half * 4 + j;
          if (q < 2 * ksteps) {
            float4 x = v[j];
            if (has_act && src >= 0) {
              float sc[4], sh[4];
#pragma unroll
              for (int e = 0; e < 4; e++) {
                bool ok = 4 * q + e < kvalid;
                sc[e] = ok ? __ldg(&p.in_scale[c0 + 4 * q + e]) : 0.f;
                sh[e] = ok ? __ldg(&p.in_shift[c0 + 4 * q + e]) : 0.f;
              }
              x.x = fmaxf(fmaf(x.x, sc[0], sh[0]), 0.f);
              x.y = fmaxf(fmaf(x.y, sc[1], sh[1]), 0.f);
              x.z = fmaxf(fmaf(x.z, sc[2], sh[2]), 0.f);
              x.w = fmaxf(fmaf(x.w, sc[3], sh[3]), 0.f);
            }
            float4 h = tf32_hi(x);
            float4 l = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
            Ahi[q * TC_ROWS + r] = h;  // chunk q, row r: byte offset q*2048 + r*16
            Alo[q * TC_ROWS + r] = l;
          }
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to UMMA
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(Ahi), a_lo = smem_u32(Alo), b_hi = smem_u32(Bhi), b_lo = smem_u32(Blo);
        const uint32_t a_lbo = TC_ROWS * 16, b_lbo = (uint32_t)nt * 16;
        for (int ks = 0; ks < ksteps; ks++) {
          uint64_t dah = umma_desc(a_hi + ks * 2 * a_lbo, a_lbo, 128);
          uint64_t dal = umma_desc(a_lo + ks * 2 * a_lbo, a_lbo, 128);
          uint64_t dbh = umma_desc(b_hi + ks * 2 * b_lbo, b_lbo, 128);
          uint64_t dbl = umma_desc(b_lo + ks * 2 * b_lbo, b_lbo, 128);
          umma_tf32(tmem, dah, dbh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          umma_tf32(tmem, dah, dbl, idesc, 1u);
          umma_tf32(tmem, dal, dbh, idesc, 1u);
        }
        umma_commit(smem_u32(&bars[s]));  // frees this stage once the MMAs above have read it
      }
    }
  }
  if (it > 0) {
    if (tid == 0) umma_commit(smem_u32(&bars[4]));
    mbar_wait(smem_u32(&bars[4]), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // ---- epilogue: TMEM lane = output row; warp w reads lanes 32*(w%4).., column half w/4 -----------------
  {
    const int lane_grp = warp & 3, chalf = warp >> 2;
    const int row = row0 + lane_grp * 32 + (tid & 31);
    const int cbeg = chalf * (nt >> 1), cend = cbeg + (nt >> 1);
    for (int cb = cbeg; cb < cend; cb += 8) {
      uint32_t v[8];
      if (it > 0) {
        uint32_t taddr = tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)cb;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = 0u;
      }
      if (row < p.Mout) {
        const int col = n0 + cb;
        float *op = p.out + (size_t)row * p.out_stride + p.out_off + col;
        const float *rp = p.residual ? p.residual + (size_t)row * p.res_stride + p.res_off + col : nullptr;
#pragma unroll
        for (int e = 0; e < 8; e++) {
          if (col + e < p.Cout) {
            float x = __uint_as_float(v[e]);
            if (p.bias) x += __ldg(&p.bias[col + e]);
            if (rp) x += __ldg(rp + e);
            op[e] = x;
          }
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

extern "C" {

// Packed weight size in floats for sgb_spconv_forward_tc: K * ceil(Cin/32) * 8 * N * 4 with N = Cout rounded to 16.
long long sgb_spconv_tc_packed_floats(int K, int Cin, int Cout) {
  int N = (Cout + 15) / 16 * 16;
  int nkc = (Cin + 31) / 32;
  return (long long)K * nkc * 8 * N * 4;
}

int sgb_spconv_forward_tc(const float *d_in, int in_stride, int in_off, const int32_t *d_map, int K, int Mout,
                          const float *d_Whi, const float *d_Wlo, int Cin, int Cout, const float *d_in_scale,
                          const float *d_in_shift, const float *d_residual, int res_stride, int res_off,
                          const float *d_bias, float *d_out, int out_stride, int out_off, void *stream) {
  if (Mout == 0 || Cout == 0) return SGB_OK;
  SGB_REQUIRE(d_in && d_Whi && d_Wlo && d_out && K >= 1 && Mout > 0 && Cin > 0 && Cout > 0, SGB_ERR_ARG,
              "spconv_forward_tc arguments");
  SGB_REQUIRE(d_map || K == 1, SGB_ERR_ARG, "identity map requires K == 1");
  SGB_REQUIRE((d_in_scale == nullptr) == (d_in_shift == nullptr), SGB_ERR_ARG, "scale/shift must come together");
  SGB_REQUIRE(in_stride >= in_off + Cin && out_stride >= out_off + Cout, SGB_ERR_ARG, "row strides");
  int N = (Cout + 15) / 16 * 16;
  SGB_REQUIRE(N <= 256, SGB_ERR_RANGE, "spconv_forward_tc: Cout > 256 is not tiled");
  TcArgs p;
  p.in = d_in; p.in_stride = in_stride; p.in_off = in_off;
  p.map = d_map; p.K = K; p.Mout = Mout;
  p.Whi = d_Whi; p.Wlo = d_Wlo; p.Cin = Cin; p.N = N; p.Cout = Cout;
  p.in_scale = d_in_scale; p.in_shift = d_in_shift;
  p.residual = d_residual; p.res_stride = res_stride; p.res_off = res_off;
  p.bias = d_bias;
  p.out = d_out; p.out_stride = out_stride; p.out_off = out_off;
  // Column split: few row tiles (deep U-Net levels) would leave most SMs idle and make one CTA stream all the
  // weights, so N is cut into NT-column CTAs until the grid covers the machine (NT multiple of 16, >= 32).
  int tiles = div_up(Mout, TC_ROWS);
  int NT = N;
  while (NT > 32 && tiles * div_up(N, NT) < kNumSMs) {
    int nxt = (NT / 2 + 15) / 16 * 16;
    if (nxt >= NT) break;
    NT = nxt;
  }
  p.NT = NT;
  size_t stage = 2 * (size_t)TC_ROWS * TC_KC * 4 + 2 * (size_t)NT * TC_KC * 4;
  p.nstages = (stage * 3 <= 110 * 1024) ? 3 : 2;
  int cols = 32;
  while (cols < NT) cols <<= 1;
  p.tmem_cols = cols;
  size_t smem = stage * p.nstages + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    SGB_CUDA_CHECK(cudaFuncSetAttribute(spconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    attr_set = true;
  }
  dim3 grid(tiles, div_up(N, NT));
  spconv_tc_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(p);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}
}
