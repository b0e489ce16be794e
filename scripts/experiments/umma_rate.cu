// microbench.cu -- test hooks only (never on the product path): issue-rate micro-benchmarks of tcgen05.mma used to
// choose the conv kernel's instruction shapes (scripts/umma_rate.py; numbers in DESIGN.md 3.2).
#include "common.cuh"
#include "tcgen05.cuh"

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
      if (a_in_tmem >= 100) {
        // kind::f16 issue patterns of the conv kernel, nt = N: 100 = [N=2nt at D | N=nt at D+nt] (overlapping accumulators),
        // 101 = three N=nt products into ONE accumulator, 102 = [N=2nt at D | N=nt at a disjoint accumulator]
        const uint32_t i1 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t i2 = (1u << 4) | ((uint32_t)((2 * N) >> 3) << 17) | (8u << 24);
        for (int k = 0; k < per_commit; k++) {
          const uint32_t ta = tmem + 384 + 8 * (k & 7);
          if (a_in_tmem == 100) {
            umma_f16_ts(tmem, ta, db, i2, 1u);
            umma_f16_ts(tmem + N, ta + 16, db, i1, 1u);
          } else if (a_in_tmem == 101) {
            umma_f16_ts(tmem, ta, db, i1, 1u);
            umma_f16_ts(tmem, ta + 16, db, i1, 1u);
            umma_f16_ts(tmem, ta, db, i1, 1u);
          } else {
            umma_f16_ts(tmem, ta, db, i2, 1u);
            umma_f16_ts(tmem + 256, ta + 16, db, i1, 1u);
          }
        }
      } else if (a_in_tmem > 0) {
        for (int k = 0; k < per_commit; k++) umma_f16_ts(tmem + 64 * (k % a_in_tmem), tmem + 256 + 8 * (k & 7), db, idesc, 1u);
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
