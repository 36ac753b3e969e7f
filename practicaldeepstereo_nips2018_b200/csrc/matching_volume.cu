// a1 -- Matching.forward data movement (reference matching.py:50-63) for an
// arbitrary `operation`: one kernel builds the disparity-stacked operation
// input, one kernel re-stacks a batched operation output.
//
//   volume[b, d, c,     y, x] = left [b, c, y, x]                     c <  C
//   volume[b, d, C + c, y, x] = x >= d ? right[b, c, y, x - d] : 0    (matching.py:56-60)
//
// HBM-bound and write-dominated (C2: 17.7 MB read, 849 MB written).  One CTA
// owns one (b, channel, y) row: the source row is staged in shared memory once
// and re-used for the whole disparity sweep, every store is a coalesced
// 128-bit streaming store.  (With MatchingOperation the volume is never built:
// the fused path in matching_op.cu reads the descriptors directly.)
#include "pds_common.cuh"

namespace pds {
namespace {

constexpr int kRowThreads = 128;

template <typename T>
__global__ void __launch_bounds__(kRowThreads)
matching_concat_kernel(const T* __restrict__ left, const T* __restrict__ right,
                       T* __restrict__ volume, int C, int H, int W, int D) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* row = reinterpret_cast<T*>(smem_raw);          // [D-1 zeros][W values], 16B slack
  const int y = blockIdx.x, c2 = blockIdx.y, b = blockIdx.z;
  const bool is_left = c2 < C;
  const T* src = (is_left ? left + ((size_t)(b * C + c2) * H + y) * W
                          : right + ((size_t)(b * C + (c2 - C)) * H + y) * W);
  const int pad = is_left ? 0 : D - 1;
  for (int i = threadIdx.x; i < pad; i += kRowThreads) row[i] = T(0.f);
  for (int i = threadIdx.x; i < W; i += kRowThreads) row[pad + i] = src[i];
  __syncthreads();
  const size_t dstride = (size_t)2 * C * H * W;
  T* dst = volume + (((size_t)b * D) * 2 * C + c2) * H * W + (size_t)y * W;
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = (W % VEC == 0) && ((reinterpret_cast<uintptr_t>(volume) & 15) == 0);
  if (vec_ok) {
    const int nv = W / VEC;
    for (int i = threadIdx.x; i < nv * D; i += kRowThreads) {
      const int d = i / nv, v = i - d * nv;
      const T* s = row + pad - (is_left ? 0 : d) + v * VEC;   // shifted[x] = right[x-d]
      union { T e[VEC]; float4 f; } u;
#pragma unroll
      for (int k = 0; k < VEC; ++k) u.e[k] = s[k];
      stg_stream(reinterpret_cast<float4*>(dst + (size_t)d * dstride) + v, u.f);
    }
  } else {
    for (int i = threadIdx.x; i < W * D; i += kRowThreads) {
      const int d = i / W, x = i - d * W;
      dst[(size_t)d * dstride + x] = row[pad - (is_left ? 0 : d) + x];
    }
  }
}

// Backward of the volume (f4, training): the adjoint of the gather above,
//   grad_left [b, c, y, x] = sum_d g[b, d, c,     y, x]
//   grad_right[b, c, y, x] = sum_{d : x + d < W} g[b, d, C + c, y, x + d]
// (what autograd accumulates through the reference's D cat / pad / slice nodes, matching.py:53-60).
// Same ownership as the forward: one CTA per (b, channel, y) row, a thread per column walking d in
// ascending order (fixed summation order: deterministic), four independent partial sums in flight.
// HBM-bound, read-dominated: the volume gradient is read exactly once.
template <typename T>
__global__ void __launch_bounds__(kRowThreads)
matching_concat_backward_kernel(const T* __restrict__ gvolume, T* __restrict__ gleft, T* __restrict__ gright,
                                int C, int H, int W, int D) {
  const int y = blockIdx.x, c2 = blockIdx.y, b = blockIdx.z;
  const bool is_left = c2 < C;
  const size_t dstride = (size_t)2 * C * H * W;
  const T* src = gvolume + (((size_t)b * D) * 2 * C + c2) * H * W + (size_t)y * W;
  T* dst = is_left ? gleft + ((size_t)(b * C + c2) * H + y) * W : gright + ((size_t)(b * C + (c2 - C)) * H + y) * W;
  for (int x = threadIdx.x; x < W; x += kRowThreads) {
    const int nd = is_left ? D : min(D, W - x);        // right: only disparities whose shifted column exists
    const T* s = src + x;
    const size_t step = is_left ? dstride : dstride + 1;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int d = 0;
    for (; d + 4 <= nd; d += 4) {
      const float v0 = (float)__ldg(s + (size_t)d * step), v1 = (float)__ldg(s + (size_t)(d + 1) * step);
      const float v2 = (float)__ldg(s + (size_t)(d + 2) * step), v3 = (float)__ldg(s + (size_t)(d + 3) * step);
      a0 += v0; a1 += v1; a2 += v2; a3 += v3;
    }
    for (; d < nd; ++d) a0 += (float)__ldg(s + (size_t)d * step);
    dst[x] = T((a0 + a1) + (a2 + a3));
  }
}

// in (B*D, F, H, W) -> out (B, F, D, H, W): plane-granular permutation (INVERSE: the other way,
// which is also its backward).
template <typename T, bool INVERSE>
__global__ void matching_stack_kernel(const T* __restrict__ in, T* __restrict__ out, int F,
                                      int D, size_t plane) {
  const int d = blockIdx.y % D, f = blockIdx.y / D, b = blockIdx.z;
  const size_t i_bdf = (((size_t)b * D + d) * F + f) * plane, i_bfd = (((size_t)b * F + f) * D + d) * plane;
  const T* s = in + (INVERSE ? i_bfd : i_bdf);
  T* o = out + (INVERSE ? i_bdf : i_bfd);
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = (plane % VEC == 0) && (((reinterpret_cast<uintptr_t>(in) |
                                               reinterpret_cast<uintptr_t>(out)) & 15) == 0);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {
    const float4* s4 = reinterpret_cast<const float4*>(s);
    float4* o4 = reinterpret_cast<float4*>(o);
    for (size_t i = t; i < plane / VEC; i += stride) stg_stream(o4 + i, ldg_stream(s4 + i));
  } else {
    for (size_t i = t; i < plane; i += stride) o[i] = s[i];
  }
}

}  // namespace
}  // namespace pds

extern "C" int pds_matching_concat(const void* left, const void* right, void* volume, int B,
                                   int C, int H, int W, int D, int dtype, void* stream) {
  using namespace pds;
  PDS_CHECK_ARG(left && right && volume, "pds_matching_concat: null pointer");
  PDS_CHECK_ARG(B >= 0 && C >= 1 && H >= 0 && W >= 0 && D >= 1, "pds_matching_concat: bad shape");
  PDS_CHECK_ARG(dtype == PDS_F32 || dtype == PDS_BF16, "pds_matching_concat: bad dtype");
  PDS_CHECK_ARG(2 * C <= 65535 && B <= 65535, "pds_matching_concat: C or B too large");
  if (B == 0 || H == 0 || W == 0) return PDS_OK;
  const size_t esz = dtype == PDS_F32 ? 4 : 2;
  const size_t smem = align_up((size_t)(W + D - 1) * esz, 16) + 16;
  PDS_CHECK_ARG(smem <= 200 * 1024, "pds_matching_concat: row too long for shared memory");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)H, (unsigned)(2 * C), (unsigned)B);
  PDS_KERNEL("matching_concat", st);
  PDS_KERNEL_WORK(0, (double)B * H * W * esz * (2.0 * C + 2.0 * C * D));
  if (dtype == PDS_F32) {
    if (smem > 48 * 1024)
      PDS_CUDA(cudaFuncSetAttribute(matching_concat_kernel<float>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    matching_concat_kernel<float><<<grid, kRowThreads, smem, st>>>(
        (const float*)left, (const float*)right, (float*)volume, C, H, W, D);
  } else {
    if (smem > 48 * 1024)
      PDS_CUDA(cudaFuncSetAttribute(matching_concat_kernel<__nv_bfloat16>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    matching_concat_kernel<__nv_bfloat16><<<grid, kRowThreads, smem, st>>>(
        (const __nv_bfloat16*)left, (const __nv_bfloat16*)right, (__nv_bfloat16*)volume, C, H, W, D);
  }
  PDS_LAUNCH_CHECK("matching_concat_kernel");
  return PDS_OK;
}

namespace pds {
namespace {
int launch_stack(const char* what, const void* in, void* out, int B, int F, int D, int H, int W, int dtype,
                 bool inverse, void* stream) {
  PDS_CHECK_ARG(in && out, "%s: null pointer", what);
  PDS_CHECK_ARG(B >= 0 && F >= 1 && D >= 1 && H >= 0 && W >= 0, "%s: bad shape", what);
  PDS_CHECK_ARG(dtype == PDS_F32 || dtype == PDS_BF16, "%s: bad dtype", what);
  PDS_CHECK_ARG((size_t)F * D <= 65535 && B <= 65535, "%s: F*D or B too large", what);
  const size_t plane = (size_t)H * W;
  if (B == 0 || plane == 0) return PDS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned gx = (unsigned)((plane / 4 + 255) / 256 > 8 ? 8 : (plane / 4 + 255) / 256);
  dim3 grid(gx ? gx : 1, (unsigned)(F * D), (unsigned)B);
  PDS_KERNEL(inverse ? "matching_unstack" : "matching_stack", st);
  PDS_KERNEL_WORK(0, 2.0 * B * F * D * plane * (dtype == PDS_F32 ? 4 : 2));
  if (dtype == PDS_F32) {
    if (inverse) matching_stack_kernel<float, true><<<grid, 256, 0, st>>>((const float*)in, (float*)out, F, D, plane);
    else matching_stack_kernel<float, false><<<grid, 256, 0, st>>>((const float*)in, (float*)out, F, D, plane);
  } else {
    if (inverse)
      matching_stack_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, F, D, plane);
    else
      matching_stack_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, F, D, plane);
  }
  PDS_LAUNCH_CHECK("matching_stack_kernel");
  return PDS_OK;
}
}  // namespace
}  // namespace pds

extern "C" int pds_matching_stack(const void* in, void* out, int B, int F, int D, int H, int W,
                                  int dtype, void* stream) {
  return pds::launch_stack("pds_matching_stack", in, out, B, F, D, H, W, dtype, false, stream);
}

extern "C" int pds_matching_unstack(const void* in, void* out, int B, int F, int D, int H, int W,
                                    int dtype, void* stream) {
  return pds::launch_stack("pds_matching_unstack", in, out, B, F, D, H, W, dtype, true, stream);
}

extern "C" int pds_matching_concat_backward(const void* grad_volume, void* grad_left, void* grad_right, int B,
                                            int C, int H, int W, int D, int dtype, void* stream) {
  using namespace pds;
  PDS_CHECK_ARG(grad_volume && grad_left && grad_right, "pds_matching_concat_backward: null pointer");
  PDS_CHECK_ARG(B >= 0 && C >= 1 && H >= 0 && W >= 0 && D >= 1, "pds_matching_concat_backward: bad shape");
  PDS_CHECK_ARG(dtype == PDS_F32 || dtype == PDS_BF16, "pds_matching_concat_backward: bad dtype");
  PDS_CHECK_ARG(2 * C <= 65535 && B <= 65535, "pds_matching_concat_backward: C or B too large");
  if (B == 0 || H == 0 || W == 0) return PDS_OK;
  const size_t esz = dtype == PDS_F32 ? 4 : 2;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)H, (unsigned)(2 * C), (unsigned)B);
  PDS_KERNEL("matching_concat_backward", st);
  PDS_KERNEL_WORK(0, (double)B * H * W * esz * (2.0 * C + 2.0 * C * D));
  if (dtype == PDS_F32)
    matching_concat_backward_kernel<float><<<grid, kRowThreads, 0, st>>>(
        (const float*)grad_volume, (float*)grad_left, (float*)grad_right, C, H, W, D);
  else
    matching_concat_backward_kernel<__nv_bfloat16><<<grid, kRowThreads, 0, st>>>(
        (const __nv_bfloat16*)grad_volume, (__nv_bfloat16*)grad_left, (__nv_bfloat16*)grad_right, C, H, W, D);
  PDS_LAUNCH_CHECK("matching_concat_backward_kernel");
  return PDS_OK;
}
