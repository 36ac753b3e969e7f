// a4 -- SubpixelMap.__call__ (reference estimator.py:45-91) as ONE streaming
// pass over the cost volume.
//
// cost is (B, D, H, W): the disparity axis is the OUTER (stride H*W) axis, so
// the bandwidth-optimal mapping is threads along W with 128-bit loads and a
// serial scan along D.  The reference needs ~70 kernels and 10 host syncs
// (boolean-mask index_put_, estimator.py:74,77); here each thread keeps, in
// registers, the running maximum with its index and the R values on either
// side of it:
//   - the R values BEFORE the maximum come from a rolling history,
//   - the R values AFTER it are captured as the scan passes idx+1 .. idx+R,
// so every cost element is read from HBM exactly once and nothing is re-read.
// Algorithmic bytes: D*H'*W*sizeof(T) read + H'*W'*4 written (H', W' after the
// fused SizeAdapter.unpad crop, size_adapter.py:51-52): cropped rows are never
// loaded.
#include <stdlib.h>

#include <type_traits>

#include "pds_common.cuh"

namespace pds {
namespace {

template <typename T, int V>
struct Vec;
template <>
struct Vec<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    float4 r = ldg_stream(reinterpret_cast<const float4*>(p));
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
};
template <>
struct Vec<float, 2> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[2]) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    v[0] = r.x; v[1] = r.y;
  }
};
template <>
struct Vec<float, 1> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
};
template <>
struct Vec<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 r = ldg_stream(reinterpret_cast<const uint2*>(p));
    v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
  }
};
template <>
struct Vec<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[1]) {
    v[0] = __bfloat162float(*p);
  }
};

// th.max semantics: strictly-greater keeps the LOWEST index on
// ties; a NaN beats any number and the FIRST NaN is kept.
__device__ __forceinline__ bool takes_over(float v, float best) {
  return (v > best) || ((v != v) && (best == best));
}

// One thread = V consecutive pixels of one row.  R = half_support_window/step.
template <typename T, int V, int R, int UNROLL, int BS>
__global__ void __launch_bounds__(BS)
subpixel_map_kernel(const T* __restrict__ cost, float* __restrict__ disparity,
                    int64_t* __restrict__ argmax, int D, int H, int W, int step,
                    int crop_top, int crop_left, int quads_per_row) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int Hc = H - crop_top;
  const int b = blockIdx.y;
  if (q >= quads_per_row * Hc) return;
  const int y = crop_top + q / quads_per_row;
  const int x0 = (q % quads_per_row) * V;
  const size_t plane = (size_t)H * W;
  const T* p = cost + (size_t)b * D * plane + (size_t)y * W + x0;

  float best[V], before[V][R], after[V][R], prev[V][R];
  int idx[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    idx[i] = 0; best[i] = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) { before[i][r] = 0.f; after[i][r] = 0.f; prev[i][r] = 0.f; }
  }
  // The scan works on chunks of planes held in registers; per element it only runs the
  // chunk-local first-maximum (compare + two selects).  The window around the maximum is
  // captured once per chunk with compile-time register indices (select chains on the local
  // arg-max j): prev[] carries the last R values of the previous chunk, after-values that lie
  // beyond the chunk are filled in at the start of the next one.  The final state equals the
  // element-by-element scan of the reference semantics (lowest index on ties, first NaN wins).
  auto process = [&](auto uc, int d0, const float (&v)[decltype(uc)::value][V]) {
    constexpr int U = decltype(uc)::value;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      // (1) after-values of the current maximum that fall into this chunk
      const int off0 = d0 - idx[i];                      // >= 1 (ignored in the first chunk)
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int u = 0; u <= r && u < U; ++u)
          if (off0 == r + 1 - u) after[i][r] = v[u][i];
      // (2) first maximum of the chunk
      float m = v[0][i];
      int j = 0;
#pragma unroll
      for (int u = 1; u < U; ++u) {
        const bool t = takes_over(v[u][i], m);
        m = t ? v[u][i] : m;
        j = t ? u : j;
      }
      // (3) does it take over?  (4) capture its window
      if (d0 == 0 || takes_over(m, best[i])) {
        best[i] = m; idx[i] = d0 + j;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (j == u) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
              // x_{j-1-r}: inside the chunk or from the previous chunk's tail
              before[i][r] = (u - 1 - r >= 0) ? v[(u - 1 - r >= 0) ? u - 1 - r : 0][i]
                                              : prev[i][(r - u >= 0 && r - u < R) ? r - u : 0];
              if (u + 1 + r < U) after[i][r] = v[(u + 1 + r < U) ? u + 1 + r : 0][i];
            }
          }
        }
      }
      // (5) tail of this chunk for the next one
#pragma unroll
      for (int r = R - 1; r >= 0; --r)      // descending: a chunk shorter than R shifts the old tail up
        prev[i][r] = (U - 1 - r >= 0) ? v[(U - 1 - r >= 0) ? U - 1 - r : 0][i] : prev[i][(r - U >= 0) ? r - U : 0];
    }
  };
  int d = 0;
  for (; d + UNROLL <= D; d += UNROLL) {
    float v[UNROLL][V];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) Vec<T, V>::load(p + (size_t)(d + u) * plane, v[u]);
    process(std::integral_constant<int, UNROLL>(), d, v);
  }
  for (; d < D; ++d) {
    float v[1][V];
    Vec<T, V>::load(p + (size_t)d * plane, v[0]);
    process(std::integral_constant<int, 1>(), d, v);
  }

  const int Wc = W - crop_left;
  const size_t orow = ((size_t)b * Hc + (y - crop_top)) * Wc;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int x = x0 + i;
    if (x < crop_left || x >= W) continue;
    // softmax over the window, max-subtracted (the centre IS the maximum):
    // estimator.py:88-90.  Out-of-range taps: similarity -inf -> weight 0,
    // disparity 0 (estimator.py:71-83).  Summation order = shift order.
    float e[2 * R + 1], sum = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * R + 1; ++k) {
      const int j = idx[i] + k - R;
      float s = k < R ? before[i][R - 1 - k] : (k == R ? best[i] : after[i][k - R - 1]);
      const bool valid = (j >= 0) && (j < D);
      e[k] = valid ? expf(s - best[i]) : 0.f;
      sum += e[k];
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * R + 1; ++k) {
      const int j = idx[i] + k - R;
      const bool valid = (j >= 0) && (j < D);
      float pk = e[k] / sum;
      if (sizeof(T) == 2) pk = __bfloat162float(__float2bfloat16_rn(pk));  // bf16 softmax
      acc += pk * (valid ? (float)(step * j) : 0.f);
    }
    disparity[orow + (x - crop_left)] = acc;
    if (argmax) argmax[orow + (x - crop_left)] = idx[i];
  }
}

// Generic window radius (R > 4): arg-max pass, then gather the taps.
template <typename T>
__global__ void subpixel_map_generic_kernel(const T* __restrict__ cost,
                                            float* __restrict__ disparity,
                                            int64_t* __restrict__ argmax, int D, int H,
                                            int W, int R, int step, int crop_top,
                                            int crop_left) {
  const int Hc = H - crop_top, Wc = W - crop_left;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= Hc * Wc) return;
  const int y = crop_top + i / Wc, x = crop_left + i % Wc;
  const size_t plane = (size_t)H * W;
  const T* p = cost + (size_t)b * D * plane + (size_t)y * W + x;
  float best = (float)p[0];
  int idx = 0;
  for (int d = 1; d < D; ++d) {
    const float v = (float)p[(size_t)d * plane];
    if (takes_over(v, best)) { best = v; idx = d; }
  }
  float sum = 0.f;
  for (int k = -R; k <= R; ++k) {
    const int j = idx + k;
    if (j >= 0 && j < D) sum += expf((float)p[(size_t)j * plane] - best);
  }
  float acc = 0.f;
  for (int k = -R; k <= R; ++k) {
    const int j = idx + k;
    if (j < 0 || j >= D) continue;
    float pk = expf((float)p[(size_t)j * plane] - best) / sum;
    if (sizeof(T) == 2) pk = __bfloat162float(__float2bfloat16_rn(pk));
    acc += pk * (float)(step * j);
  }
  disparity[(size_t)b * Hc * Wc + i] = acc;
  if (argmax) argmax[(size_t)b * Hc * Wc + i] = idx;
}

template <typename T, int V>
int launch(const T* cost, float* disparity, int64_t* argmax, int B, int D, int H, int W,
           int R, int step, int crop_top, int crop_left, cudaStream_t st) {
  const int Hc = H - crop_top;
  const int quads = (W + V - 1) / V;
  PDS_KERNEL("subpixel_map", st);
  // rows above crop_top are never loaded; every other cost element is read once
  PDS_KERNEL_WORK(0, (double)B * Hc * ((double)D * W * sizeof(T) + (double)(W - crop_left) * 4));
  // 128 threads, chunks of 6 planes: measured best at C2 (50 us = 4.0 TB/s; chunks of 4 / 8 / 12 /
  // 16: 57 / 71 / 63 / 61 us; the element-by-element scan was instruction-bound at 77-112 us; with the
  // next chunk's loads issued before the current chunk is scanned -- twice the bytes in flight per
  // thread, 24 more registers -- 81 us for chunks of 6 and 72 us for chunks of 4)
  dim3 grid((unsigned)(((size_t)quads * Hc + 127) / 128), (unsigned)B);
#define PDS_EST(RR)                                                                  \
  subpixel_map_kernel<T, V, RR, 6, 128><<<grid, 128, 0, st>>>(cost, disparity, argmax, D, H, W, \
                                                              step, crop_top, crop_left, quads)
  switch (R) {
    case 1: PDS_EST(1); break;
    case 2: PDS_EST(2); break;
    case 3: PDS_EST(3); break;
    case 4: PDS_EST(4); break;
    default: {
      dim3 g((unsigned)(((size_t)Hc * (W - crop_left) + 127) / 128), (unsigned)B);
      subpixel_map_generic_kernel<T><<<g, 128, 0, st>>>(cost, disparity, argmax, D, H, W, R,
                                                       step, crop_top, crop_left);
    }
  }
#undef PDS_EST
  PDS_LAUNCH_CHECK("subpixel_map_kernel");
  return PDS_OK;
}

}  // namespace
}  // namespace pds

extern "C" int pds_subpixel_map(const void* cost, float* disparity, int64_t* argmax, int B,
                                int D, int H, int W, int half_support_window,
                                int disparity_step, int crop_top, int crop_left, int dtype,
                                void* stream) {
  using namespace pds;
  // estimator.py:34-41
  PDS_CHECK_ARG(disparity_step >= 1, "\"disparity_step\" should be positive integer.");
  PDS_CHECK_ARG(half_support_window >= 1, "\"half_support_window\" should be positive integer.");
  PDS_CHECK_ARG(half_support_window % disparity_step == 0,
                "\"half_support_window\" should be multiple of the\"disparity_step\"");
  PDS_CHECK_ARG(B >= 0 && D >= 1 && H >= 0 && W >= 0, "pds_subpixel_map: bad shape");
  PDS_CHECK_ARG((cost && disparity) || B == 0 || H == 0 || W == 0, "pds_subpixel_map: null pointer");
  PDS_CHECK_ARG(crop_top >= 0 && crop_top <= H && crop_left >= 0 && crop_left <= W,
                "pds_subpixel_map: crop outside the image");
  PDS_CHECK_ARG(dtype == PDS_F32 || dtype == PDS_BF16, "pds_subpixel_map: bad dtype");
  if (B == 0 || H - crop_top == 0 || W - crop_left == 0) return PDS_OK;
  PDS_CHECK_ARG(B <= 65535, "pds_subpixel_map: batch > 65535");
  cudaStream_t st = (cudaStream_t)stream;
  const int R = half_support_window / disparity_step;
  const size_t esz = dtype == PDS_F32 ? 4 : 2;
  const bool vec = (W % 4 == 0) && (((uintptr_t)cost) % (4 * esz) == 0);
  if (dtype == PDS_F32) {
    const float* c = (const float*)cost;
    return vec ? launch<float, 4>(c, disparity, argmax, B, D, H, W, R, disparity_step, crop_top, crop_left, st)
               : launch<float, 1>(c, disparity, argmax, B, D, H, W, R, disparity_step, crop_top, crop_left, st);
  }
  const __nv_bfloat16* c = (const __nv_bfloat16*)cost;
  return vec ? launch<__nv_bfloat16, 4>(c, disparity, argmax, B, D, H, W, R, disparity_step, crop_top, crop_left, st)
             : launch<__nv_bfloat16, 1>(c, disparity, argmax, B, D, H, W, R, disparity_step, crop_top, crop_left, st);
}
