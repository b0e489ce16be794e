// dataprep.cu -- the test-time transform of the reference dataloader on the GPU.
// Replaces the numpy `np.matmul(xyz, m)` of CustomDataset.dataAugment with augmentation off
// (softgroup/data/custom.py:87-107: m = eye(3) @ rot_z(0.35 pi), float64) -- the first step of transform_test (:162-168).
// The product is evaluated in float64 like numpy does for a float32 [N,3] x float64 [3,3] matmul; the accumulation order
// is the k-ascending fused chain of the BLAS dgemm micro-kernels  acc = fma(a_k, b_kj, acc), acc_0 = a_0 * b_0j.
#include "common.cuh"

namespace sgb {
__global__ void affine3_f64_kernel(const float *__restrict__ xyz, double m00, double m01, double m02, double m10, double m11,
                                   double m12, double m20, double m21, double m22, double *__restrict__ out, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double x = (double)xyz[3 * (size_t)i], y = (double)xyz[3 * (size_t)i + 1], z = (double)xyz[3 * (size_t)i + 2];
  out[3 * (size_t)i + 0] = __fma_rn(z, m20, __fma_rn(y, m10, __dmul_rn(x, m00)));
  out[3 * (size_t)i + 1] = __fma_rn(z, m21, __fma_rn(y, m11, __dmul_rn(x, m01)));
  out[3 * (size_t)i + 2] = __fma_rn(z, m22, __fma_rn(y, m12, __dmul_rn(x, m02)));
}
}  // namespace sgb

extern "C" int sgb_affine3_f64(const float *d_xyz, const double *h_m9, double *d_out, int N, void *stream) {
  if (N == 0) return SGB_OK;
  SGB_REQUIRE(d_xyz && h_m9 && d_out && N > 0, SGB_ERR_ARG, "affine3_f64 arguments");
  sgb::affine3_f64_kernel<<<sgb::div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(d_xyz, h_m9[0], h_m9[1], h_m9[2], h_m9[3], h_m9[4],
                                                                              h_m9[5], h_m9[6], h_m9[7], h_m9[8], d_out, N);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}
