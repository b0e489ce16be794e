// umma_rate2.cu -- stand-alone issue-rate micro-benchmark of tcgen05.mma kind::f16 (M = 128, K = 16) instruction patterns
// (round 2, session 2). Not part of the product. Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_rate2 scripts/experiments/umma_rate2.cu && /tmp/umma_rate2
// Question answered: what does ONE small-N MMA cost when it accumulates into the same TMEM columns as its predecessor,
// and does rotating over independent accumulators (several row tiles per CTA) remove that cost?
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
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc) : "memory");
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

// mode: 0 = TS (A in TMEM), B no-swizzle; 1 = SS, A and B no-swizzle; 2 = SS, A and B SWIZZLE_128B; 3 = TS, B SWIZZLE_128B
// pattern: 0 = one instruction per step, N columns, accumulator rotates over `rot` disjoint ranges
//          1 = production pair: [N = 2 nt at D | N = nt at D + nt], accumulator pair rotates over `rot` disjoint ranges
__global__ void rate_kernel(int mode, int pattern, int N, int rot, int reps, int per_commit, long long *out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t s_tmem;
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
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
  if (warp == 1) {  // whole warp runs the loop, one elected lane issues (warp-uniform operands -> bare UTCHMMA)
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t a_smem = smem_u32(smem), b_smem = smem_u32(smem) + 32 * 1024;
    const int acc_cols = (pattern == 1) ? 2 * N : N;  // columns of one accumulator (set)
    const uint32_t idescN = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t idesc2N = (1u << 4) | ((uint32_t)((2 * N) >> 3) << 17) | (8u << 24);
    const bool sw = (mode == 2 || mode == 3);
    const uint64_t da = sw ? desc_sw128(a_smem) : desc_none(a_smem, 2048, 128);
    const uint64_t db = sw ? desc_sw128(b_smem) : desc_none(b_smem, (uint32_t)acc_cols * 16, 128);
    const uint32_t a_t = tmem + 448;  // A operand columns (TS): 4 k-steps x (8 hi + 8 lo) = 64 columns
    long long t0 = clock64();
    uint32_t phase = 0;
    int acc = 0;
    for (int r = 0; r < reps; r += per_commit) {
      for (int k = 0; k < per_commit; k++) {
        const uint32_t d = tmem + (uint32_t)(acc * acc_cols);
        const uint32_t ak = a_t + 8u * (uint32_t)(k & 3);
        const uint64_t ka = da + (uint64_t)((sw ? 32u : 4096u) >> 4) * (uint64_t)(k & 3);
        const uint64_t kb = db + (uint64_t)((sw ? 32u : 2u * (uint32_t)acc_cols * 16u) >> 4) * (uint64_t)(k & 3);
        if (leader) {
          if (pattern == 0) {
            if (mode == 0 || mode == 3) mma_ts(d, ak, kb, idescN); else mma_ss(d, ka, kb, idescN);
          } else {
            if (mode == 0 || mode == 3) { mma_ts(d, ak, kb, idesc2N); mma_ts(d + (uint32_t)N, ak + 32u, kb, idescN); }
            else { mma_ss(d, ka, kb, idesc2N); mma_ss(d + (uint32_t)N, ka + (8192u >> 4), kb, idescN); }
          }
        }
        if (++acc == rot) acc = 0;
      }
      if (leader) umma_commit(smem_u32(&bar));
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
    }
    long long t1 = clock64();
    if (leader) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long *d_out, h;
  cudaMalloc(&d_out, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  const char *mname[4] = {"TS  B:none ", "SS  AB:none", "SS  AB:sw128", "TS  B:sw128"};
  const int reps = 4096;
  for (int pattern = 0; pattern < 2; pattern++)
    for (int mode = 0; mode < 4; mode++)
      for (int N : {32, 64, 96, 128, 192, 256}) {
        const int acc_cols = pattern ? 2 * N : N;
        if (acc_cols > 256 && pattern == 1) continue;
        for (int rot : {1, 2, 4}) {
          if (rot * acc_cols > 448) continue;
          for (int pc : {8, 64}) {
            rate_kernel<<<1, 64, 66 * 1024>>>(mode, pattern, N, rot, reps, pc, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
            printf("%s %s N %3d rot %d per_commit %2d : %7.1f cycles per %s\n", pattern ? "pair[2N|N]" : "single    ", mname[mode], N, rot,
                   pc, (double)h / reps, pattern ? "k-step (2 MMAs)" : "MMA");
          }
        }
      }
  return 0;
}
