// umma_rate3.cu -- follow-up of umma_rate2.cu: where does the ~110-cycle minimum per tcgen05.mma k-step come from, and does
// it overlap (a) between two CTAs resident on one SM, (b) between two issuing warps of one CTA, (c) when the B descriptor
// does not change? kind::f16, M = 128, K = 16, SS form, pair pattern [N = 2 nt | N = nt] of the conv kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate3 umma_rate3.cu && ./umma_rate3
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ uint64_t desc_none(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// issuers: 1 or 2 warps of the CTA issue independent streams (own accumulator, own barrier). same_b: every k-step uses the
// same B descriptor. tmem_cols: 256 (two CTAs fit on an SM) or 512. Every issuing warp runs `reps` k-steps of the pair
// pattern, `per_commit` k-steps per commit+wait. out[block * 2 + issuer] = cycles.
__global__ void rate_kernel(int N, int issuers, int same_b, int tmem_cols, int reps, int per_commit, long long *out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar[2];
  __shared__ uint32_t s_tmem;
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (48 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
  if (tid == 0) {
    mbar_init(smem_u32(&bar[0]), 1);
    mbar_init(smem_u32(&bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"((uint32_t)tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (warp >= 1 && warp <= issuers) {
    const int me = warp - 1;
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t a_smem = smem_u32(smem), b_smem = smem_u32(smem) + 16 * 1024;
    const uint32_t idescN = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t idesc2N = (1u << 4) | ((uint32_t)((2 * N) >> 3) << 17) | (8u << 24);
    const uint64_t da = desc_sw128(a_smem);
    const uint64_t db = desc_none(b_smem, (uint32_t)(2 * N) * 16, 128);
    const uint32_t d = tmem + (uint32_t)(me * 2 * N);
    const uint32_t mybar = smem_u32(&bar[me]);
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int r = 0; r < reps; r += per_commit) {
#pragma unroll 4
      for (int k = 0; k < per_commit; k++) {
        const uint64_t ka = da + 2u * (uint64_t)(k & 1);
        const uint64_t kb = same_b ? db : db + (uint64_t)((2u * (uint32_t)(2 * N) * 16u) >> 4) * (uint64_t)(k & 1);
        if (leader) {
          mma_ss(d, ka, kb, idesc2N);
          mma_ss(d + (uint32_t)N, ka + 4u, kb, idescN);
        }
      }
      if (leader) umma_commit(mybar);
      __syncwarp();
      mbar_wait(mybar, phase);
      phase ^= 1;
    }
    long long t1 = clock64();
    if (leader) out[blockIdx.x * 2 + me] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)tmem_cols) : "memory");
}

int main() {
  const int maxb = 296;
  long long *d_out, h[2 * maxb];
  cudaMalloc(&d_out, sizeof(h));
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
  const int reps = 4096;
  for (int N : {32, 64, 96, 128})
    for (int same_b : {0, 1})
      for (int cfg = 0; cfg < 4; cfg++) {
        // cfg 0: one CTA on the GPU, one issuer; 1: one CTA, two issuing warps; 2: 148 CTAs (one per SM); 3: 296 CTAs (two per SM)
        const int issuers = cfg == 1 ? 2 : 1;
        const int grid = cfg == 2 ? 148 : cfg == 3 ? 296 : 1;
        if (issuers * 2 * N > 256) continue;
        for (int pc : {4, 32}) {
          cudaMemset(d_out, 0, sizeof(h));
          rate_kernel<<<grid, 96, 50 * 1024>>>(N, issuers, same_b, 256, reps, pc, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int i = 0; i < 2 * grid; i++) mx = h[i] > mx ? h[i] : mx;
          const char *names[4] = {"1 CTA, 1 issuer ", "1 CTA, 2 issuers", "148 CTAs (1/SM) ", "296 CTAs (2/SM) "};
          printf("nt %3d  %s  B desc %s  per_commit %2d : %7.1f cycles per k-step per issuer (slowest)\n", N, names[cfg],
                 same_b ? "fixed  " : "changes", pc, (double)mx / reps);
        }
      }
  return 0;
}
