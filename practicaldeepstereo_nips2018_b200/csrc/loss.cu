// f4 (training half) -- SubpixelCrossEntropy (reference loss.py:16-78) as two streaming kernels.
//
// The reference loops over the disparity axis in Python: log_softmax of the whole (B, D, H, W)
// volume, then per disparity index an exp / abs / divide pass and two accumulations -- about 6 D
// elementwise launches and a full-volume temporary.  Per location (b, y, x) with ground truth g:
//     P_t(d) = exp(-|g - step d| / diversity) / (2 diversity)          (un-normalised Laplace)
//     entropy = - sum_d P_t(d) (s_d - lse(s)) / sum_d P_t(d)
//     loss = sum w entropy / (sum w + 1e-15)   over locations with finite g   (weights given)
//          = mean of entropy over those locations                            (no weights)
// Forward: ONE pass over the volume -- online log-sum-exp (running maximum + rescaled sum) next to
// sum P_t and sum P_t s_d -- writes per location the entropy, lse and sum P_t and accumulates the two
// global sums in double.  Backward: ONE pass,
//     d loss / d s_d = upstream * c * (softmax(s)_d - P_t(d) / sum P_t),   c = w / (sum w + 1e-15) or 1 / N
// and d loss / d w = upstream * (entropy - loss) / (sum w + 1e-15).
// The volume is (B, D, H, W) with the disparity axis outermost per sample: threads along W, 128-bit
// accesses, serial walk along D.  HBM-bound: 4 D bytes read per location forward, 8 D backward.
#include "pds_common.cuh"

namespace pds {
namespace {

__device__ __forceinline__ float laplace(float g, float disparity, float inv_div, float norm) {
  return expf(-fabsf(g - disparity) * inv_div) * norm;     // loss.py:12-13
}

template <int V>
__global__ void __launch_bounds__(128)
sce_forward_kernel(const float* __restrict__ sim, const float* __restrict__ gt, const float* __restrict__ weights,
                   float* __restrict__ entropy, float* __restrict__ lse, float* __restrict__ sum_pt,
                   double* __restrict__ sums, int D, size_t HW, size_t groups_per_sample, float step, float inv_div,
                   float norm) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  double acc_num = 0.0, acc_den = 0.0;
  if (q < groups_per_sample) {
    const size_t pix = q * V;
    const float* s = sim + (size_t)b * D * HW + pix;
    float g[V], m[V], e[V], spt[V], sps[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      g[i] = pix + i < HW ? gt[(size_t)b * HW + pix + i] : INFINITY;
      m[i] = -INFINITY; e[i] = 0.f; spt[i] = 0.f; sps[i] = 0.f;
    }
    for (int d = 0; d < D; ++d) {
      float v[V];
      if (V == 4) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(s + (size_t)d * HW));
        v[0] = r.x; v[1 % V] = r.y; v[2 % V] = r.z; v[3 % V] = r.w;
      } else {
        v[0] = __ldg(s + (size_t)d * HW);
      }
      const float disparity = step * (float)d;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float nm = fmaxf(m[i], v[i]);
        e[i] = e[i] * expf(m[i] - nm) + expf(v[i] - nm);       // online log-sum-exp
        m[i] = nm;
        const float pt = laplace(g[i], disparity, inv_div, norm);
        spt[i] += pt;
        sps[i] = fmaf(pt, v[i], sps[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      if (pix + i >= HW) continue;
      const size_t o = (size_t)b * HW + pix + i;
      const bool known = !isinf(g[i]);                           // loss.py:52-53
      const float l = m[i] + logf(e[i]);
      const float ent = known ? -(sps[i] - l * spt[i]) / spt[i] : 0.f;
      entropy[o] = ent; lse[o] = l; sum_pt[o] = spt[i];
      if (known) {
        const float w = weights ? weights[o] : 1.f;
        acc_num += (double)w * (double)ent;
        acc_den += (double)w;
      }
    }
  }
  __shared__ double rn[4], rd[4];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    acc_num += __shfl_xor_sync(0xffffffffu, acc_num, o);
    acc_den += __shfl_xor_sync(0xffffffffu, acc_den, o);
  }
  if ((threadIdx.x & 31) == 0) { rn[threadIdx.x >> 5] = acc_num; rd[threadIdx.x >> 5] = acc_den; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 4; ++w) { acc_num += rn[w]; acc_den += rd[w]; }
    if (acc_den != 0.0 || acc_num != 0.0) { atomicAdd(sums, acc_num); atomicAdd(sums + 1, acc_den); }
  }
}

template <int V>
__global__ void __launch_bounds__(128)
sce_backward_kernel(const float* __restrict__ sim, const float* __restrict__ gt, const float* __restrict__ weights,
                    const float* __restrict__ entropy, const float* __restrict__ lse,
                    const float* __restrict__ sum_pt, const double* __restrict__ sums,
                    const float* __restrict__ upstream, float* __restrict__ grad_sim,
                    float* __restrict__ grad_weights, int D, size_t HW, size_t groups_per_sample, float step,
                    float inv_div, float norm, int has_weights) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (q >= groups_per_sample) return;
  const size_t pix = q * V;
  const double den = has_weights ? sums[1] + 1e-15 : sums[1];      // loss.py:74-78
  const float up = __ldg(upstream);
  const float loss = (float)(sums[0] / den);
  float g[V], c[V], l[V], inv_spt[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const bool in = pix + i < HW;
    const size_t o = (size_t)b * HW + pix + i;
    g[i] = in ? gt[o] : INFINITY;
    const bool known = in && !isinf(g[i]);
    const float w = (known && weights) ? weights[o] : 1.f;
    c[i] = known ? up * (float)((double)w / den) : 0.f;
    l[i] = in ? lse[o] : 0.f;
    inv_spt[i] = known ? 1.f / sum_pt[o] : 0.f;
    if (in && grad_weights) grad_weights[o] = known ? up * (float)(((double)entropy[o] - (double)loss) / den) : 0.f;
  }
  const float* s = sim + (size_t)b * D * HW + pix;
  float* gs = grad_sim + (size_t)b * D * HW + pix;
  for (int d = 0; d < D; ++d) {
    float v[V], r[V];
    if (V == 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(s + (size_t)d * HW));
      v[0] = t.x; v[1 % V] = t.y; v[2 % V] = t.z; v[3 % V] = t.w;
    } else {
      v[0] = __ldg(s + (size_t)d * HW);
    }
    const float disparity = step * (float)d;
#pragma unroll
    for (int i = 0; i < V; ++i)
      r[i] = c[i] == 0.f ? 0.f : c[i] * (expf(v[i] - l[i]) - laplace(g[i], disparity, inv_div, norm) * inv_spt[i]);
    if (V == 4) *reinterpret_cast<float4*>(gs + (size_t)d * HW) = make_float4(r[0], r[1 % V], r[2 % V], r[3 % V]);
    else gs[(size_t)d * HW] = r[0];
  }
}

}  // namespace
}  // namespace pds

extern "C" int pds_subpixel_cross_entropy_forward(const float* similarities, const float* ground_truth,
                                                  const float* weights, float* entropy, float* lse, float* sum_pt,
                                                  double* sums, int B, int D, int H, int W, float diversity,
                                                  int disparity_step, void* stream) {
  using namespace pds;
  PDS_CHECK_ARG(similarities && ground_truth && entropy && lse && sum_pt && sums,
                "pds_subpixel_cross_entropy_forward: null pointer");
  PDS_CHECK_ARG(B >= 0 && D >= 1 && H >= 1 && W >= 1 && diversity > 0.f && disparity_step >= 1,
                "pds_subpixel_cross_entropy_forward: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  PDS_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), st));
  if (B == 0) return PDS_OK;
  const size_t HW = (size_t)H * W;
  const bool vec = HW % 4 == 0 && ((uintptr_t)similarities & 15) == 0;
  const size_t groups = vec ? HW / 4 : HW;
  dim3 grid((unsigned)((groups + 127) / 128), (unsigned)B);
  PDS_KERNEL("subpixel_cross_entropy_forward", st);
  PDS_KERNEL_WORK(0, (double)B * HW * (4.0 * D + 16.0 + (weights ? 4.0 : 0.0)));
  const float inv_div = 1.f / diversity, norm = 1.f / (2.f * diversity);
  if (vec) sce_forward_kernel<4><<<grid, 128, 0, st>>>(similarities, ground_truth, weights, entropy, lse, sum_pt, sums, D, HW, groups, (float)disparity_step, inv_div, norm);
  else sce_forward_kernel<1><<<grid, 128, 0, st>>>(similarities, ground_truth, weights, entropy, lse, sum_pt, sums, D, HW, groups, (float)disparity_step, inv_div, norm);
  PDS_LAUNCH_CHECK("sce_forward_kernel");
  return PDS_OK;
}

extern "C" int pds_subpixel_cross_entropy_backward(const float* similarities, const float* ground_truth,
                                                   const float* weights, const float* entropy, const float* lse,
                                                   const float* sum_pt, const double* sums, const float* upstream,
                                                   float* grad_similarities, float* grad_weights, int B, int D,
                                                   int H, int W, float diversity, int disparity_step,
                                                   void* stream) {
  using namespace pds;
  PDS_CHECK_ARG(similarities && ground_truth && entropy && lse && sum_pt && sums && upstream && grad_similarities,
                "pds_subpixel_cross_entropy_backward: null pointer");
  PDS_CHECK_ARG(B >= 0 && D >= 1 && H >= 1 && W >= 1 && diversity > 0.f && disparity_step >= 1,
                "pds_subpixel_cross_entropy_backward: bad arguments");
  PDS_CHECK_ARG(!grad_weights || weights, "pds_subpixel_cross_entropy_backward: grad_weights without weights");
  if (B == 0) return PDS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t HW = (size_t)H * W;
  const bool vec = HW % 4 == 0 && ((uintptr_t)similarities & 15) == 0 && ((uintptr_t)grad_similarities & 15) == 0;
  const size_t groups = vec ? HW / 4 : HW;
  dim3 grid((unsigned)((groups + 127) / 128), (unsigned)B);
  PDS_KERNEL("subpixel_cross_entropy_backward", st);
  PDS_KERNEL_WORK(0, (double)B * HW * (8.0 * D + 20.0));
  const float inv_div = 1.f / diversity, norm = 1.f / (2.f * diversity);
  if (vec) sce_backward_kernel<4><<<grid, 128, 0, st>>>(similarities, ground_truth, weights, entropy, lse, sum_pt, sums, upstream, grad_similarities, grad_weights, D, HW, groups, (float)disparity_step, inv_div, norm, weights ? 1 : 0);
  else sce_backward_kernel<1><<<grid, 128, 0, st>>>(similarities, ground_truth, weights, entropy, lse, sum_pt, sums, upstream, grad_similarities, grad_weights, D, HW, groups, (float)disparity_step, inv_div, norm, weights ? 1 : 0);
  PDS_LAUNCH_CHECK("sce_backward_kernel");
  return PDS_OK;
}
