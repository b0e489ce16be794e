// tma_probe.cu -- development probe (never part of libsgb200.so): what the round-2 conv kernel needs to know about
// TMA row gather on sm_100a before it is built on it.
//   A. layout: cp.async.bulk.tensor.2d ... tile::gather4 with SWIZZLE_128B into a 1024-byte aligned tile -- where does
//      each (row, 16-byte chunk) land, what happens to out-of-bounds row indices (-1, >= rows)?
//   B. MMA: tcgen05.mma kind::f16, A from that tile through a K-major SWIZZLE_128B shared-memory descriptor (SBO 1024,
//      K advance = +32 B on the start address: hi k0, hi k1, lo k0, lo k1 of a [hi 32ch | lo 32ch] row), B from the
//      product's no-swizzle core-matrix layout; result vs the host.
//   C. throughput: persistent CTAs stream 128-row x 128-byte stages through a ring (P producer warps issue 32 gather4
//      per stage, a consumer warp frees the stage) -- GB/s vs ring depth, producer warps, CTAs per SM, row pattern.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/bin/tma_probe scripts/tma_probe.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); exit(1); }
  return (EncodeTiledFn)fn;
}

// rows x cols 32-bit words, row stride = cols * 4 bytes; box = 32 words (128 B) x 1 row, 128-byte swizzle
static CUtensorMap make_map(void *base, uint64_t rows, uint64_t cols) {
  static EncodeTiledFn enc = get_encode();
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 4};
  cuuint32_t box[2] = {32, 1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return m;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap *tm, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// A. layout dump
// ---------------------------------------------------------------------------------------------------------------
__global__ void layout_kernel(const __grid_constant__ CUtensorMap tm, const int *rows, int ngather, int col, uint32_t *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x;
  for (int i = tid; i < ngather * 128; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0xDEADBEEFu;
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    mbar_expect(smem_u32(&bar), (uint32_t)ngather * 512u);
    for (int g = 0; g < ngather; g++)
      gather4(smem_u32(smem) + g * 512, &tm, col, rows[4 * g], rows[4 * g + 1], rows[4 * g + 2], rows[4 * g + 3], smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  for (int i = tid; i < ngather * 128; i += blockDim.x) out[i] = reinterpret_cast<uint32_t *>(smem)[i];
}

// ---------------------------------------------------------------------------------------------------------------
// B. SS-form MMA from the gathered tile
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {  // K-major, SWIZZLE_128B, SBO = 1024, LBO = 0
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_none(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__global__ void mma_kernel(const __grid_constant__ CUtensorMap tm, const int *rows, const __half *Bg, float *out) {
  // A: 128 rows x 128 B (gathered), B: [2 chunks][32 n][8 halves] = 1024 B, D: 4 results x 32 columns
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar, done;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char *As = smem, *Bs = smem + 16384;
  for (int i = tid; i < 512; i += blockDim.x) reinterpret_cast<__half *>(Bs)[i] = Bg[i];
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (warp == 1) {
    if (lane == 0) mbar_expect(smem_u32(&bar), 16384u);
    __syncwarp();
    gather4(smem_u32(As) + lane * 512, &tm, 0, rows[4 * lane], rows[4 * lane + 1], rows[4 * lane + 2], rows[4 * lane + 3], smem_u32(&bar));
  }
  if (tid == 0) {
    mbar_wait(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bd = desc_none(smem_u32(Bs), 32 * 16, 128);
    for (int v = 0; v < 4; v++) {  // hi k0, hi k1, lo k0, lo k1: +32 bytes each on the start address
      const uint64_t ad = desc_sw128(smem_u32(As) + v * 32);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem + v * 32), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
  }
  __syncthreads();
  mbar_wait(smem_u32(&done), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    for (int c = 0; c < 128; c += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int e = 0; e < 8; e++) out[(warp * 32 + lane) * 128 + c + e] = __uint_as_float(v[e]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// C. throughput
// ---------------------------------------------------------------------------------------------------------------
// idx: [iters_total][128] row indices. CTA b handles iterations b, b + grid, ... Warps 0..P-1 produce, warp P consumes.
__global__ void __launch_bounds__(192) stream_kernel(const __grid_constant__ CUtensorMap tm, const int *__restrict__ idx, int iters_total,
                                                     int S, int P, int ncol) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[16], empty[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(smem_u32(&full[s]), P); mbar_init(smem_u32(&empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int per = 32 / P;  // gather4 per warp per stage
  if (warp < P) {
    int s = 0, u = 0;
    const bool act = lane < per;
    const int g = warp * per + lane;
    int4 nxt = make_int4(0, 0, 0, 0);
    int it = blockIdx.x;
    if (it < iters_total && act) nxt = __ldg(reinterpret_cast<const int4 *>(idx + (size_t)it * 128) + g);
    int n = 0;
    for (; it < iters_total; it += gridDim.x, n++) {
      const int4 cur = nxt;
      const int it2 = it + gridDim.x;
      if (it2 < iters_total && act) nxt = __ldg(reinterpret_cast<const int4 *>(idx + (size_t)it2 * 128) + g);
      if (u >= 1) mbar_wait(smem_u32(&empty[s]), (uint32_t)((u - 1) & 1));
      if (lane == 0) mbar_expect(smem_u32(&full[s]), (uint32_t)per * 512u);
      __syncwarp();
      if (act) gather4(smem_u32(smem) + s * 16384 + g * 512, &tm, (n % ncol) * 32, cur.x, cur.y, cur.z, cur.w, smem_u32(&full[s]));
      if (++s == S) { s = 0; u++; }
    }
  } else if (warp == P) {
    int s = 0;
    uint32_t par = 0;
    for (int it = blockIdx.x; it < iters_total; it += gridDim.x) {
      mbar_wait(smem_u32(&full[s]), par);
      if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
      __syncwarp();
      if (++s == S) { s = 0; par ^= 1; }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// D. issue-cost variants. mode 0: every gather4 issued from warp-uniform operands (row ids broadcast with __shfl_sync,
//    unrolled: no ELECT/R2UR.BROADCAST/BRA.U.ANY waterfall per lane); mode 1: cp.async 16 B x 8 lanes per row written
//    straight into the SWIZZLE_128B positions (4 rows per warp instruction), completion by cp.async.mbarrier.arrive.
//    P producer warps, warp P consumes.
// ---------------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(544) stream2_kernel(const __grid_constant__ CUtensorMap tm, const uint32_t *__restrict__ table, int rowwords,
                                                      int nrows, const int *__restrict__ idx, int iters_total, int S, int P, int ncol) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[16], empty[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(smem_u32(&full[s]), MODE == 0 ? P : P * 32); mbar_init(smem_u32(&empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int per = 32 / P;  // row quads per warp per stage
  const uint32_t base = smem_u32(smem);
  if (warp < P) {
    int s = 0, u = 0, n = 0;
    for (int it = blockIdx.x; it < iters_total; it += gridDim.x, n++) {
      const int col = (n % ncol) * 32;
      if (MODE == 0) {
        // lane j < per keeps the 4 row ids of quad warp*per + j
        int4 cur = make_int4(0, 0, 0, 0);
        if (lane < per) cur = __ldg(reinterpret_cast<const int4 *>(idx + (size_t)it * 128) + warp * per + lane);
        if (u >= 1) mbar_wait(smem_u32(&empty[s]), (uint32_t)((u - 1) & 1));
        const uint32_t bar = smem_u32(&full[s]);
        if (lane == 0) mbar_expect(bar, (uint32_t)per * 512u);
        __syncwarp();
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
#pragma unroll 8
        for (int j = 0; j < per; j++) {
          const int r0 = __shfl_sync(0xffffffffu, cur.x, j), r1 = __shfl_sync(0xffffffffu, cur.y, j);
          const int r2 = __shfl_sync(0xffffffffu, cur.z, j), r3 = __shfl_sync(0xffffffffu, cur.w, j);
          const uint32_t dst = base + (uint32_t)s * 16384u + (uint32_t)(warp * per + j) * 512u;
          if (leader) gather4(dst, &tm, col, r0, r1, r2, r3, bar);
        }
      } else {
        int rid[8];
#pragma unroll
        for (int j = 0; j < 8; j++) rid[j] = (j < per) ? __ldg(idx + (size_t)it * 128 + (warp * per + j) * 4 + (lane >> 3)) : 0;
        if (u >= 1) mbar_wait(smem_u32(&empty[s]), (uint32_t)((u - 1) & 1));
        const uint32_t bar = smem_u32(&full[s]);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (j < per) {
            const int row = (warp * per + j) * 4 + (lane >> 3);
            const int ch = lane & 7;
            const uint32_t dst = base + (uint32_t)s * 16384u + (uint32_t)row * 128u + (uint32_t)((ch ^ (row & 7)) << 4);
            const bool ok = rid[j] >= 0 && rid[j] < nrows;
            const uint32_t *g = table + (size_t)(ok ? rid[j] : 0) * rowwords + col + ch * 4;
            const int nbytes = ok ? 16 : 0;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(g), "r"(nbytes) : "memory");
          }
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
      }
      if (++s == S) { s = 0; u++; }
    }
    if (MODE == 1) asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == P) {
    int s = 0;
    uint32_t par = 0;
    for (int it = blockIdx.x; it < iters_total; it += gridDim.x) {
      mbar_wait(smem_u32(&full[s]), par);
      if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
      __syncwarp();
      if (++s == S) { s = 0; par ^= 1; }
    }
  }
}

int main() {
  const int M = 137000, C = 64;  // level-0-like table: 137k rows, 2 chunks of 32 words (256 B per row)
  std::vector<uint32_t> h((size_t)(M + 1) * C);
  // every 32-bit word = two fp16: value pattern small integers
  auto val = [](int r, int hidx) { return (float)(((r * 7 + hidx * 3) % 17) - 8); };
  for (int r = 0; r < M; r++)
    for (int w = 0; w < C; w++) {
      __half a = __float2half(val(r, 2 * w)), b = __float2half(val(r, 2 * w + 1));
      uint16_t ua, ub;
      memcpy(&ua, &a, 2); memcpy(&ub, &b, 2);
      h[(size_t)r * C + w] = (uint32_t)ua | ((uint32_t)ub << 16);
    }
  for (int w = 0; w < C; w++) h[(size_t)M * C + w] = 0;  // zero row
  uint32_t *d_tab;
  CK(cudaMalloc(&d_tab, h.size() * 4));
  CK(cudaMemcpy(d_tab, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap tm = make_map(d_tab, M + 1, C);

  // ---- A ----------------------------------------------------------------------------------------------------
  {
    int hr[8] = {5, 77, 123456, 9, M, -1, M + 5, 3};
    int *d_rows; uint32_t *d_out;
    CK(cudaMalloc(&d_rows, sizeof(hr))); CK(cudaMalloc(&d_out, 2 * 512));
    CK(cudaMemcpy(d_rows, hr, sizeof(hr), cudaMemcpyHostToDevice));
    layout_kernel<<<1, 128, 2048>>>(tm, d_rows, 2, 32, d_out);
    CK(cudaDeviceSynchronize());
    uint32_t ho[256];
    CK(cudaMemcpy(ho, d_out, sizeof(ho), cudaMemcpyDeviceToHost));
    int ok = 1;
    for (int rr = 0; rr < 8; rr++) {
      printf("A row slot %d (src %d):", rr, hr[rr]);
      for (int ch = 0; ch < 8; ch++) {
        // which source chunk sits in physical chunk ch of row slot rr?
        uint32_t w0 = ho[rr * 32 + ch * 4];
        int found = -1;
        if (hr[rr] >= 0 && hr[rr] <= M)
          for (int sc = 0; sc < 8; sc++)
            if (w0 == h[(size_t)hr[rr] * C + 32 + sc * 4]) { found = sc; break; }
        printf(" %d", found);
        if (hr[rr] >= 0 && hr[rr] < M && found != (ch ^ (rr & 7))) ok = 0;
      }
      printf("  first word 0x%08x\n", ho[rr * 32]);
    }
    printf("A: swizzle (chunk ^ (row&7)) as expected: %s\n", ok ? "YES" : "NO");
  }
  // ---- B ----------------------------------------------------------------------------------------------------
  {
    std::vector<int> hr(128);
    for (int i = 0; i < 128; i++) hr[i] = (i * 9973 + 17) % M;
    hr[5] = M; hr[77] = M;  // zero rows
    std::vector<__half> hb(512);
    auto bval = [](int k, int n) { return (float)(((k * 5 + n * 3) % 7) - 3); };
    for (int q = 0; q < 2; q++)
      for (int n = 0; n < 32; n++)
        for (int e = 0; e < 8; e++) hb[(q * 32 + n) * 8 + e] = __float2half(bval(q * 8 + e, n));
    int *d_rows; __half *d_b; float *d_out;
    CK(cudaMalloc(&d_rows, 512)); CK(cudaMalloc(&d_b, 1024)); CK(cudaMalloc(&d_out, 128 * 128 * 4));
    CK(cudaMemcpy(d_rows, hr.data(), 512, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_b, hb.data(), 1024, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480));
    mma_kernel<<<1, 128, 16384 + 1024 + 1024>>>(tm, d_rows, d_b, d_out);
    CK(cudaDeviceSynchronize());
    std::vector<float> ho(128 * 128);
    CK(cudaMemcpy(ho.data(), d_out, ho.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int r = 0; r < 128; r++)
      for (int v = 0; v < 4; v++)
        for (int n = 0; n < 32; n++) {
          double want = 0;
          for (int k = 0; k < 16; k++) want += (hr[r] < M ? val(hr[r], v * 16 + k) : 0.f) * bval(k, n);
          double e = fabs(want - ho[r * 128 + v * 32 + n]);
          if (e > maxerr) maxerr = e;
        }
    printf("B: SS-form MMA from the gathered SWIZZLE_128B tile, 4 K offsets: max |err| = %g (%s)\n", maxerr, maxerr == 0 ? "EXACT" : "MISMATCH");
  }
  // ---- C ----------------------------------------------------------------------------------------------------
  {
    const int iters_total = 148 * 400;
    std::vector<int> hi((size_t)iters_total * 128);
    int *d_idx;
    CK(cudaMalloc(&d_idx, hi.size() * 4));
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int pattern = 0; pattern < 3; pattern++) {
      uint64_t st = 88172645463325252ull;
      auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
      for (int it = 0; it < iters_total; it++) {
        const int base = (int)(rnd() % (M - 4096));
        for (int r = 0; r < 128; r++) {
          int v;
          if (pattern == 0) v = (int)(rnd() % M);                       // uniform random rows (L2-resident table)
          else if (pattern == 1) v = base + (int)(rnd() % 2048);        // local window of 2048 rows
          else v = (rnd() % 100 < 50) ? M : base + (int)(rnd() % 2048); // 50% absent neighbours (zero row)
          hi[(size_t)it * 128 + r] = v;
        }
      }
      CK(cudaMemcpy(d_idx, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice));
      const char *pn[3] = {"random", "local2048", "local+50%zero-row"};
      for (int occ = 1; occ <= 2; occ++)
        for (int P : {1, 2, 4})
          for (int S : {6}) {
            size_t smem = (size_t)S * 16384 + 1024;
            if (smem * occ > 220 * 1024) continue;
            const int grid = 148 * occ;
            stream_kernel<<<grid, 32 * (P + 1), smem>>>(tm, d_idx, iters_total, S, P, 2);  // warm-up
            CK(cudaEventRecord(e0));
            stream_kernel<<<grid, 32 * (P + 1), smem>>>(tm, d_idx, iters_total, S, P, 2);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            double gb = (double)iters_total * 16384 / 1e9;
            printf("C: %-18s occ %d P %d S %2d: %.3f ms  %.0f GB/s  (%.1f B/clk/SM at 1.965 GHz; %.0f clk per 16 KB stage per SM)\n", pn[pattern],
                   occ, P, S, ms, gb / (ms * 1e-3), gb * 1e9 / (ms * 1e-3) / 148 / 1.965e9, ms * 1e-3 * 1.965e9 / (iters_total / 148.0));
          }
    }
  }
  // ---- D ----------------------------------------------------------------------------------------------------
  {
    const int iters_total = 148 * 400;
    std::vector<int> hi((size_t)iters_total * 128);
    int *d_idx;
    CK(cudaMalloc(&d_idx, hi.size() * 4));
    CK(cudaFuncSetAttribute(stream2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(stream2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CUtensorMap tm2 = make_map(d_tab, M, C);  // M rows: index M is out of bounds (zero fill, no memory access)
    for (int pattern = 0; pattern < 2; pattern++) {
      uint64_t st = 88172645463325252ull;
      auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
      for (int it = 0; it < iters_total; it++) {
        const int base = (int)(rnd() % (M - 4096));
        for (int r = 0; r < 128; r++)
          hi[(size_t)it * 128 + r] = (pattern == 1 && rnd() % 100 < 50) ? M : base + (int)(rnd() % 2048);
      }
      CK(cudaMemcpy(d_idx, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice));
      const char *pn[2] = {"local2048", "local+50%OOB"};
      for (int mode = 0; mode < 2; mode++)
        for (int occ = 1; occ <= 2; occ++)
          for (int P : {2, 4, 8, 16})
            for (int S : {4, 6}) {
              size_t smem = (size_t)S * 16384 + 1024;
              if (smem * occ > 220 * 1024) continue;
              if (32 * (P + 1) * occ > 2048) continue;
              const int grid = 148 * occ;
              for (int rep = 0; rep < 2; rep++) {
                if (rep == 1) CK(cudaEventRecord(e0));
                if (mode == 0) stream2_kernel<0><<<grid, 32 * (P + 1), smem>>>(tm2, d_tab, C, M, d_idx, iters_total, S, P, 2);
                else stream2_kernel<1><<<grid, 32 * (P + 1), smem>>>(tm2, d_tab, C, M, d_idx, iters_total, S, P, 2);
              }
              CK(cudaEventRecord(e1));
              CK(cudaDeviceSynchronize());
              float ms;
              CK(cudaEventElapsedTime(&ms, e0, e1));
              double gb = (double)iters_total * 16384 / 1e9;
              printf("D: %-14s %-22s occ %d P %2d S %d: %.3f ms  %.0f GB/s  (%.1f B/clk/SM; %.0f clk per 16 KB stage per SM)\n", pn[pattern],
                     mode == 0 ? "gather4 uniform issue" : "cp.async 16B swizzled", occ, P, S, ms, gb / (ms * 1e-3),
                     gb * 1e9 / (ms * 1e-3) / 148 / 1.965e9, ms * 1e-3 * 1.965e9 / (iters_total / 148.0));
            }
    }
  }
  printf("done\n");
  return 0;
}
