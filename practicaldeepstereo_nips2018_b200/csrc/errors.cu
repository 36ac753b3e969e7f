// f4 (evaluation half) -- errors.compute_absolute_error / compute_n_pixels_error (reference
// errors.py:9-74), the two metrics Trainer._test reports after every forward, in ONE pass over the
// disparity map: the pixel-wise maps and the three sums the means are made of.  Locations whose
// ground truth is +-inf contribute nothing and show 0 in the maps; |e - g| > n compares false for
// NaN (torch.gt).  HBM-bound: 8 bytes in, up to 8 bytes out per pixel.
#include "pds_common.cuh"

namespace pds {
namespace {

__global__ void __launch_bounds__(256)
disparity_errors_kernel(const float* __restrict__ est, const float* __restrict__ gt,
                        float* __restrict__ abs_map, float* __restrict__ bad_map, size_t count, float n,
                        double* __restrict__ sums) {
  double s_abs = 0.0;
  unsigned long long n_valid = 0, n_bad = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    const float g = __ldg(gt + i), e = __ldg(est + i);
    const bool valid = !isinf(g);
    const float d = fabsf(e - g);
    const bool bad = d > n;
    if (valid) { s_abs += (double)d; ++n_valid; n_bad += bad ? 1 : 0; }
    if (abs_map) abs_map[i] = valid ? d : 0.f;
    if (bad_map) bad_map[i] = (valid && bad) ? 1.f : 0.f;
  }
  __shared__ double rs[8];
  __shared__ unsigned long long rv[8], rb[8];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    s_abs += __shfl_xor_sync(0xffffffffu, s_abs, o);
    n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
    n_bad += __shfl_xor_sync(0xffffffffu, n_bad, o);
  }
  if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s_abs; rv[threadIdx.x >> 5] = n_valid; rb[threadIdx.x >> 5] = n_bad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { s_abs += rs[w]; n_valid += rv[w]; n_bad += rb[w]; }
    atomicAdd(sums, s_abs);
    atomicAdd(sums + 1, (double)n_valid);     // exact: counts stay far below 2^53
    atomicAdd(sums + 2, (double)n_bad);
  }
}

}  // namespace
}  // namespace pds

extern "C" int pds_disparity_errors(const float* estimated, const float* ground_truth, float* pixelwise_abs,
                                    float* pixelwise_n_pixels, size_t count, float n, double* sums,
                                    void* stream) {
  using namespace pds;
  PDS_CHECK_ARG(sums && (count == 0 || (estimated && ground_truth)), "pds_disparity_errors: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  PDS_CUDA(cudaMemsetAsync(sums, 0, 3 * sizeof(double), st));
  if (count == 0) return PDS_OK;
  PDS_KERNEL("disparity_errors", st);
  PDS_KERNEL_WORK(0, (double)count * (8.0 + (pixelwise_abs ? 4.0 : 0.0) + (pixelwise_n_pixels ? 4.0 : 0.0)));
  size_t blocks = (count + 255) / 256;
  const size_t cap = (size_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  disparity_errors_kernel<<<(unsigned)blocks, 256, 0, st>>>(estimated, ground_truth, pixelwise_abs,
                                                          pixelwise_n_pixels, count, n, sums);
  PDS_LAUNCH_CHECK("disparity_errors_kernel");
  return PDS_OK;
}
