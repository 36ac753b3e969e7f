// f4 (training): [LeakyReLU ->] InstanceNorm(affine) forward and backward on fp32 NCHW / NCDHW
// tensors -- the non-GEMM two thirds of every block of the reference (network_blocks.py:47-58,
// 88-131: Conv -> LeakyReLU(0.1, inplace) -> InstanceNorm{2,3}d(affine)) under autograd.
//
// ATen runs nn.InstanceNorm as a batch norm over a (1, N*C, ...) view on cuDNN's bn_fw_tr_1C11 /
// bn_bw_1C11 kernels (one CTA row per channel): 16.6 + 43.7 ms of a 206 ms training step at
// 960x540, md = 255 (tools/train_step_profile.py), for ~25 GB of algorithmic traffic (4 ms at the
// HBM roofline).  Here every pass is a flat HBM stream:
//
//   forward   pass 1: z = lrelu(x);  per (n, c): sum z, sum z^2  (fp32 per thread, double from the
//                     warp on; one double atomic pair per CTA)
//             pass 2: y = (z - mean) * (rstd * gamma) + beta;  mean / rstd saved for the backward
//   backward  pass 1: per (n, c): sum dy, sum dy * zhat
//             pass 2: dz = gamma * rstd * (dy - mean(dy) - zhat * mean(dy * zhat));  dx = dz * lrelu'(x)
//   (d gamma, d beta are the per-channel sums of the pass-1 results over n: done by the caller.)
//
// The activation is fused (slope != 1): the block's LeakyReLU output is never written or saved.
#include "pds_common.cuh"

namespace pds {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : slope * v; }

// block-wide sum of two values in double; result valid in thread 0
__device__ __forceinline__ void block_sum2(double& a, double& b) {
  __shared__ double red[2][kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, o);
    b += __shfl_down_sync(0xffffffffu, b, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = a; red[1][warp] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = 0.0; b = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) { a += red[0][w]; b += red[1][w]; }
  }
}

// the chunk [begin, end) of row `row` this CTA owns (multiples of 4 elements except the row's end)
__device__ __forceinline__ void chunk_of(size_t L, size_t& begin, size_t& end) {
  const size_t per = ((L + gridDim.x - 1) / gridDim.x + 3) & ~(size_t)3;
  begin = min((size_t)blockIdx.x * per, L);
  end = min(begin + per, L);
}

// MODE 0: sums of z and z^2 (forward);  MODE 1: sums of dy and dy * zhat (backward)
template <int MODE, bool VEC>
__global__ void __launch_bounds__(kThreads)
in_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean_rstd,
                 double* __restrict__ sums, size_t L, float slope) {
  const size_t row = blockIdx.y;
  size_t begin, end;
  chunk_of(L, begin, end);
  const float* xr = x + row * L;
  const float* gr = MODE == 1 ? dy + row * L : nullptr;
  float mean = 0.f, rstd = 1.f;
  if (MODE == 1) { mean = mean_rstd[2 * row]; rstd = mean_rstd[2 * row + 1]; }
  float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;      // two independent chains per moment
  auto take = [&](float xv, float gv, float& a, float& b) {
    const float z = lrelu(xv, slope);
    if (MODE == 0) { a += z; b = fmaf(z, z, b); }
    else { a += gv; b = fmaf(gv, (z - mean) * rstd, b); }
  };
  if (VEC) {
    const float4* x4 = reinterpret_cast<const float4*>(xr);
    const float4* g4 = reinterpret_cast<const float4*>(gr);
    const size_t v0 = begin / 4, v1 = end / 4;          // VEC: L % 4 == 0, so chunk ends are multiples of 4
    size_t i = v0 + threadIdx.x;
    for (; i + kThreads < v1; i += 2 * kThreads) {
      const float4 a = ldg_stream(x4 + i), b = ldg_stream(x4 + i + kThreads);
      float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga;
      if (MODE == 1) { ga = ldg_stream(g4 + i); gb = ldg_stream(g4 + i + kThreads); }
      take(a.x, ga.x, s0, s1); take(a.y, ga.y, s0, s1); take(a.z, ga.z, s0, s1); take(a.w, ga.w, s0, s1);
      take(b.x, gb.x, t0, t1); take(b.y, gb.y, t0, t1); take(b.z, gb.z, t0, t1); take(b.w, gb.w, t0, t1);
    }
    if (i < v1) {
      const float4 a = ldg_stream(x4 + i);
      float4 ga = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 1) ga = ldg_stream(g4 + i);
      take(a.x, ga.x, s0, s1); take(a.y, ga.y, s0, s1); take(a.z, ga.z, s0, s1); take(a.w, ga.w, s0, s1);
    }
  } else {
    for (size_t i = begin + threadIdx.x; i < end; i += kThreads) take(__ldg(xr + i), MODE == 1 ? __ldg(gr + i) : 0.f, s0, s1);
  }
  double a = (double)s0 + (double)t0, b = (double)s1 + (double)t1;
  block_sum2(a, b);
  if (threadIdx.x == 0) { atomicAdd(sums + 2 * row, a); atomicAdd(sums + 2 * row + 1, b); }
}

// MODE 0: y (writes mean / rstd of the row once);  MODE 1: dx
template <int MODE, bool VEC>
__global__ void __launch_bounds__(kThreads)
in_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                const float* __restrict__ beta, const double* __restrict__ sums, float* __restrict__ mean_rstd,
                float* __restrict__ out, int C, size_t L, float eps, float slope) {
  const size_t row = blockIdx.y;
  const int c = (int)(row % C);
  size_t begin, end;
  chunk_of(L, begin, end);
  const float g = gamma ? gamma[c] : 1.f;
  float mean, rstd, k0, k1, k2;     // MODE 0: y = (z - mean) * k0 + k1;  MODE 1: dz = k0 * ((dy - k1) - zhat * k2)
  if (MODE == 0) {
    const double m = sums[2 * row] / (double)L;
    double var = sums[2 * row + 1] / (double)L - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)eps));
    if (blockIdx.x == 0 && threadIdx.x == 0) { mean_rstd[2 * row] = mean; mean_rstd[2 * row + 1] = rstd; }
    k0 = rstd * g;
    k1 = beta ? beta[c] : 0.f;
    k2 = 0.f;
  } else {
    mean = mean_rstd[2 * row]; rstd = mean_rstd[2 * row + 1];
    const float m1 = (float)(sums[2 * row] / (double)L), m2 = (float)(sums[2 * row + 1] / (double)L);
    k0 = g * rstd; k1 = m1; k2 = m2;
  }
  const float* xr = x + row * L;
  const float* gr = MODE == 1 ? dy + row * L : nullptr;
  float* o = out + row * L;
  auto value = [&](float xv, float gv) {
    const float z = lrelu(xv, slope);
    if (MODE == 0) return fmaf(z - mean, k0, k1);       // the centred form: no cancellation between z * k0 and mean * k0
    const float dz = k0 * fmaf(-(z - mean) * rstd, k2, gv - k1);
    return xv > 0.f ? dz : slope * dz;
  };
  if (VEC) {
    const float4* x4 = reinterpret_cast<const float4*>(xr);
    const float4* g4 = reinterpret_cast<const float4*>(gr);
    float4* o4 = reinterpret_cast<float4*>(o);
    const size_t v0 = begin / 4, v1 = end / 4;
    size_t i = v0 + threadIdx.x;
    for (; i + kThreads < v1; i += 2 * kThreads) {
      const float4 a = ldg_stream(x4 + i), b = ldg_stream(x4 + i + kThreads);
      float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga;
      if (MODE == 1) { ga = ldg_stream(g4 + i); gb = ldg_stream(g4 + i + kThreads); }
      stg_stream(o4 + i, make_float4(value(a.x, ga.x), value(a.y, ga.y), value(a.z, ga.z), value(a.w, ga.w)));
      stg_stream(o4 + i + kThreads, make_float4(value(b.x, gb.x), value(b.y, gb.y), value(b.z, gb.z), value(b.w, gb.w)));
    }
    if (i < v1) {
      const float4 a = ldg_stream(x4 + i);
      float4 ga = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 1) ga = ldg_stream(g4 + i);
      stg_stream(o4 + i, make_float4(value(a.x, ga.x), value(a.y, ga.y), value(a.z, ga.z), value(a.w, ga.w)));
    }
  } else {
    for (size_t i = begin + threadIdx.x; i < end; i += kThreads) o[i] = value(__ldg(xr + i), MODE == 1 ? __ldg(gr + i) : 0.f);
  }
}

dim3 grid_for(size_t rows, size_t L) {
  // ~8 CTAs per SM over all rows, at least 1024 elements per CTA
  size_t chunks = ((size_t)num_sms() * 8 + rows - 1) / rows;
  const size_t most = (L + 1023) / 1024;
  if (chunks > most) chunks = most;
  if (chunks < 1) chunks = 1;
  return dim3((unsigned)chunks, (unsigned)rows);
}

bool vec_ok(const void* a, const void* b, const void* c, size_t L) {
  return L % 4 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

}  // namespace
}  // namespace pds

extern "C" int pds_instance_norm_forward(const float* x, const float* gamma, const float* beta, float* y,
                                         float* mean_rstd, double* sums, int N, int C, long long L, float eps,
                                         float slope, void* stream) {
  using namespace pds;
  PDS_CHECK_ARG(x && y && mean_rstd && sums, "pds_instance_norm_forward: null pointer");
  PDS_CHECK_ARG(N >= 0 && C >= 1 && L >= 0, "pds_instance_norm_forward: bad shape");
  const size_t rows = (size_t)N * C;
  PDS_CHECK_ARG(rows <= 65535, "pds_instance_norm_forward: more than 65535 (sample, channel) rows");
  if (rows == 0 || L == 0) return PDS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  PDS_CUDA(cudaMemsetAsync(sums, 0, rows * 2 * sizeof(double), st));
  const dim3 grid = grid_for(rows, (size_t)L);
  const bool vec = vec_ok(x, y, x, (size_t)L);
  {
    PDS_KERNEL("instance_norm_train[sums]", st);
    PDS_KERNEL_WORK(0, 4.0 * rows * L);
    if (vec) in_reduce_kernel<0, true><<<grid, kThreads, 0, st>>>(x, nullptr, nullptr, sums, (size_t)L, slope);
    else in_reduce_kernel<0, false><<<grid, kThreads, 0, st>>>(x, nullptr, nullptr, sums, (size_t)L, slope);
    PDS_LAUNCH_CHECK("in_reduce_kernel");
  }
  {
    PDS_KERNEL("instance_norm_train[apply]", st);
    PDS_KERNEL_WORK(0, 8.0 * rows * L);
    if (vec) in_apply_kernel<0, true><<<grid, kThreads, 0, st>>>(x, nullptr, gamma, beta, sums, mean_rstd, y, C, (size_t)L, eps, slope);
    else in_apply_kernel<0, false><<<grid, kThreads, 0, st>>>(x, nullptr, gamma, beta, sums, mean_rstd, y, C, (size_t)L, eps, slope);
    PDS_LAUNCH_CHECK("in_apply_kernel");
  }
  return PDS_OK;
}

extern "C" int pds_instance_norm_backward(const float* x, const float* dy, const float* gamma,
                                          const float* mean_rstd, float* dx, double* sums, int N, int C,
                                          long long L, float slope, void* stream) {
  using namespace pds;
  PDS_CHECK_ARG(x && dy && mean_rstd && dx && sums, "pds_instance_norm_backward: null pointer");
  PDS_CHECK_ARG(N >= 0 && C >= 1 && L >= 0, "pds_instance_norm_backward: bad shape");
  const size_t rows = (size_t)N * C;
  PDS_CHECK_ARG(rows <= 65535, "pds_instance_norm_backward: more than 65535 (sample, channel) rows");
  if (rows == 0 || L == 0) return PDS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  PDS_CUDA(cudaMemsetAsync(sums, 0, rows * 2 * sizeof(double), st));
  const dim3 grid = grid_for(rows, (size_t)L);
  const bool vec = vec_ok(x, dy, dx, (size_t)L);
  {
    PDS_KERNEL("instance_norm_train[backward sums]", st);
    PDS_KERNEL_WORK(0, 8.0 * rows * L);
    if (vec) in_reduce_kernel<1, true><<<grid, kThreads, 0, st>>>(x, dy, mean_rstd, sums, (size_t)L, slope);
    else in_reduce_kernel<1, false><<<grid, kThreads, 0, st>>>(x, dy, mean_rstd, sums, (size_t)L, slope);
    PDS_LAUNCH_CHECK("in_reduce_kernel");
  }
  {
    PDS_KERNEL("instance_norm_train[backward apply]", st);
    PDS_KERNEL_WORK(0, 12.0 * rows * L);
    if (vec) in_apply_kernel<1, true><<<grid, kThreads, 0, st>>>(x, dy, gamma, nullptr, sums, const_cast<float*>(mean_rstd), dx, C, (size_t)L, 0.f, slope);
    else in_apply_kernel<1, false><<<grid, kThreads, 0, st>>>(x, dy, gamma, nullptr, sums, const_cast<float*>(mean_rstd), dx, C, (size_t)L, 0.f, slope);
    PDS_LAUNCH_CHECK("in_apply_kernel");
  }
  return PDS_OK;
}
