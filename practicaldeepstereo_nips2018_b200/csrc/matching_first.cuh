// x0 of disparity slice d regenerated from the per-sample terms of the factorised first
// convolution (derivation in matching_factor.cu / DESIGN.md 4.1):
//   x0_d = A + shift_d(B~) - [x = W-1, d >= 1] Q[W-d],   B~[-1] = Q[0].
// Shared by the passes that need the residual stream of block 1 without materialising it.
#pragma once
#include <cuda_runtime.h>

namespace pds {

// channels 8*c8 .. +7 at (y, x): a4 / b4 / q4 point at the (sample, channel-group pair) planes
// [2][H][W] of float4, `pix` = y * W + x.  Same arithmetic (and rounding) as tc_compose_first.
__device__ __forceinline__ void first_x0(const float4* a4, const float4* b4, const float4* q4, size_t HW,
                                         size_t pix, int x, int W, int d, float (&v)[8]) {
  const size_t row = pix - x;
  const float4 lo = __ldg(a4 + pix), hi = __ldg(a4 + HW + pix);
  v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
  auto add = [&](const float4* src, size_t at, float sign) {
    const float4 l = __ldg(src + at), h = __ldg(src + HW + at);
    v[0] = fmaf(sign, l.x, v[0]); v[1] = fmaf(sign, l.y, v[1]); v[2] = fmaf(sign, l.z, v[2]); v[3] = fmaf(sign, l.w, v[3]);
    v[4] = fmaf(sign, h.x, v[4]); v[5] = fmaf(sign, h.y, v[5]); v[6] = fmaf(sign, h.z, v[6]); v[7] = fmaf(sign, h.w, v[7]);
  };
  if (x >= d) add(b4, pix - d, 1.f);
  else if (x == d - 1) add(q4, row, 1.f);
  if (x == W - 1 && d >= 1 && d <= W) add(q4, row + (W - d), -1.f);
}

}  // namespace pds
