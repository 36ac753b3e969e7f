// Tail of the hourglass (reference regularization.py:90-92, 125-126):
// _upsample_to_fullsize = ConvTranspose3d(F/2 = 4 -> 1, kernel (3,4,4), stride
// (1,2,2), padding 1) without activation / normalisation, fused with the
// InstanceNorm3d of the block that produces its input.
//
// The layer has ONE output channel: it is not a tensor-core shape but a
// bandwidth kernel -- read the half-size volume (4 ch, channels-last) once,
// write the (B, 2D, 4H, 4W) cost volume once.  Algorithmic bytes per pair at
// C2: 212 MB in + 212 MB out.
//
//   out[z, 2y+cy, 2x+cx] = bias + sum over tz in 0..2, ty, tx in {0,1}, ci of
//       in[z+tz-1, y+cy-ty, x+cx-tx][ci] * W[ci][0][2-tz][1-cy+2ty][1-cx+2tx]
//
// One thread owns one (y, x) column of the input grid and marches along z with
// three running 2x2 output patches (out z-1, z, z+1); the 3x3 neighbourhood of
// the current input plane is read through L1 (neighbouring threads share it).
// The 192 weights live in the kernel parameter (constant bank): every FFMA takes
// its weight as a constant operand, no shared-memory or register traffic.
#include <stdlib.h>

#include "conv_layers.cuh"

namespace pds {
namespace {

struct TailParams {
  const float4* in;      // [B][D][H][W][4] post-LeakyReLU, pre-InstanceNorm
  float* out;            // [B][D][2H][2W]
  const double* stats;   // [B][4][2] sum, sum of squares of `in` (null: input already normalised)
  int B, D, H, W, zseg, nseg;
  float gamma[4], beta[4];
  float w[3][4][4][4];   // [kd][kh][kw][ci]
  float bias;
};

// Row DY (0..2 <-> input y - 1..y + 1) of the 3 x 3 input neighbourhoods of NP pixels of plane zi
// -> every accumulator that row feeds: out z = zi - tz + 1 uses kd = 2 - tz (slot 2 - tz: 0 is
// out zi-1, 2 is out zi+1); out row class cy takes kernel row 1 - cy + 2*ty from input row
// 1 + cy - ty.  One fused multiply-add chain per accumulator, the same order in every kernel of
// this file (the fused and the plain paths agree bit for bit); the pixel loop is innermost so the
// NP pixels share each weight fetch.
template <int DY, int NP>
__device__ __forceinline__ void tail_accumulate_row(float (&acc)[NP][3][4], const float4 (&v)[NP][3],
                                                    const float (&w)[3][4][4][4]) {
#pragma unroll
  for (int cy = 0; cy < 2; ++cy) {
    const int ty = 1 + cy - DY;
    if (ty < 0 || ty > 1) continue;
#pragma unroll
    for (int tz = 0; tz < 3; ++tz)
#pragma unroll
      for (int cx = 0; cx < 2; ++cx)
#pragma unroll
        for (int tx = 0; tx < 2; ++tx) {
          const float* ww = w[2 - tz][1 - cy + 2 * ty][1 - cx + 2 * tx];
#pragma unroll
          for (int px = 0; px < NP; ++px) {
            const float4& q = v[px][1 + cx - tx];
            float a = acc[px][2 - tz][cy * 2 + cx];
            a = fmaf(q.x, ww[0], a); a = fmaf(q.y, ww[1], a); a = fmaf(q.z, ww[2], a); a = fmaf(q.w, ww[3], a);
            acc[px][2 - tz][cy * 2 + cx] = a;
          }
        }
  }
}

// th.max semantics (estimator.cu): strictly greater keeps the lowest index; the first NaN wins.
__device__ __forceinline__ bool takes_over(float v, float best) {
  return (v > best) || ((v != v) && (best == best));
}

// Running state of SubpixelMap for one output pixel (estimator.py:59-91, same scheme as
// subpixel_map_kernel): maximum with its index, the R values before it (from a rolling
// history) and the R values after it (captured as the scan passes them).
template <int R>
struct MapState {
  float best, before[R], after[R], hist[R];
  int idx;
  __device__ __forceinline__ void init() {
    best = 0.f; idx = -1;
#pragma unroll
    for (int r = 0; r < R; ++r) { before[r] = 0.f; after[r] = 0.f; hist[r] = 0.f; }
  }
  // A scan may cover only a SEGMENT [z0, z1) of the disparity axis (segments are merged by
  // subpixel_merge_kernel): values just before the segment only warm the history up, values just
  // after it are only captured into the window of a maximum near the segment's end.
  __device__ __forceinline__ void warm(float x) {
#pragma unroll
    for (int r = R - 1; r > 0; --r) hist[r] = hist[r - 1];
    hist[0] = x;
  }
  __device__ __forceinline__ void push(float x, int d, bool first) {
    const int off = d - idx;
    if (first || takes_over(x, best)) {
      best = x; idx = d;
#pragma unroll
      for (int r = 0; r < R; ++r) before[r] = hist[r];
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) if (off == r + 1) after[r] = x;
    }
    warm(x);
  }
  __device__ __forceinline__ void tail(float x, int d) {
    const int off = d - idx;
#pragma unroll
    for (int r = 0; r < R; ++r) if (off == r + 1) after[r] = x;
  }
  // softmax over the window, max-subtracted, summation in shift order (estimator.py:66-90)
  __device__ __forceinline__ float disparity(int D, int step) const {
    float e[2 * R + 1], sum = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * R + 1; ++k) {
      const int j = idx + k - R;
      const float s = k < R ? before[R - 1 - k] : (k == R ? best : after[k - R - 1]);
      const bool valid = (j >= 0) && (j < D);
      e[k] = valid ? expf(s - best) : 0.f;
      sum += e[k];
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * R + 1; ++k) {
      const int j = idx + k - R;
      const bool valid = (j >= 0) && (j < D);
      acc += (e[k] / sum) * (valid ? (float)(step * j) : 0.f);
    }
    return acc;
  }
};

struct FusedParams {
  float* disparity;      // (B, 2H - crop_top, 2W - crop_left)
  int64_t* argmax;       // same shape or null
  int step, crop_top, crop_left;
  float* state;          // [segment][2 + 2R fields][B][2H][2W]: partial SubpixelMap states
};

// R == 0: writes the cost volume (B, D, 2H, 2W).  R >= 1: the volume is never written -- every
// thread feeds the four output pixels it owns into the SubpixelMap state (window radius R) and
// stores their disparities, cropped (SizeAdapter.unpad), at the end of its march along z.
template <int R>
__global__ void __launch_bounds__(256, R > 0 ? 2 : 0)   // fused: two CTAs per SM despite the estimator state
hourglass_tail_kernel(const __grid_constant__ TailParams p, const FusedParams f) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  const int b = blockIdx.z / p.nseg, seg = blockIdx.z - b * p.nseg;
  const int z0 = seg * p.zseg, z1 = min(p.D, z0 + p.zseg);
  if (x >= p.W || y >= p.H) return;
  if (R > 0 && (2 * y + 1 < f.crop_top || 2 * x + 1 < f.crop_left)) return;   // all four pixels cropped
  // InstanceNorm of the input as one multiply-add per channel
  float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.stats) {
    const double n = (double)p.D * p.H * p.W;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double s = p.stats[(b * 4 + c) * 2], q = p.stats[(b * 4 + c) * 2 + 1];
      const double mean = s / n;
      double var = q / n - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + 1e-5));
      sc[c] = rstd * p.gamma[c];
      sh[c] = p.beta[c] - (float)mean * sc[c];
    }
  }
  const size_t plane = (size_t)p.H * p.W;
  const float4* base = p.in + (size_t)b * p.D * plane;
  const int OW = 2 * p.W;
  const size_t oplane = (size_t)4 * plane;
  float* obase = R == 0 ? p.out + (size_t)b * p.D * oplane + (size_t)(2 * y) * OW + 2 * x : nullptr;
  bool okx[3], oky[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) { okx[d] = x + d - 1 >= 0 && x + d - 1 < p.W; oky[d] = y + d - 1 >= 0 && y + d - 1 < p.H; }

  float acc[1][3][4];   // [pixel][out z - zi + 1][cy*2+cx]
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[0][k][c] = 0.f;
  MapState<(R > 0 ? R : 1)> state[4];
  if (R > 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) state[c].init();
  }

  // fused: the scan of a segment extends R outputs to either side (history warm-up / window capture)
  const int zs = R > 0 ? max(0, z0 - R) : z0, ze = R > 0 ? min(p.D, z1 + R) : z1;
  for (int zi = zs - 1; zi <= ze; ++zi) {
    if (zi >= 0 && zi < p.D) {
      float4 v[3][1][3];
      const float4* pl = base + (size_t)zi * plane + (size_t)y * p.W + x;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          if (oky[dy] && okx[dx]) {
            float4 t = __ldg(pl + (dy - 1) * p.W + (dx - 1));
            t.x = fmaf(t.x, sc[0], sh[0]); t.y = fmaf(t.y, sc[1], sh[1]);
            t.z = fmaf(t.z, sc[2], sh[2]); t.w = fmaf(t.w, sc[3], sh[3]);
            v[dy][0][dx] = t;
          } else {
            v[dy][0][dx] = make_float4(0.f, 0.f, 0.f, 0.f);   // zero padding of the NORMALISED tensor
          }
        }
      tail_accumulate_row<0, 1>(acc, v[0], p.w);
      tail_accumulate_row<1, 1>(acc, v[1], p.w);
      tail_accumulate_row<2, 1>(acc, v[2], p.w);
    }
    const int zo = zi - 1;
    if (zo >= zs && zo < ze) {
      if (R == 0) {
        float* o = obase + (size_t)zo * oplane;
        *reinterpret_cast<float2*>(o) = make_float2(acc[0][0][0] + p.bias, acc[0][0][1] + p.bias);
        *reinterpret_cast<float2*>(o + OW) = make_float2(acc[0][0][2] + p.bias, acc[0][0][3] + p.bias);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float val = acc[0][0][c] + p.bias;
          if (zo < z0) state[c].warm(val);
          else if (zo < z1) state[c].push(val, zo, zo == z0);
          else state[c].tail(val, zo);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[0][0][c] = acc[0][1][c]; acc[0][1][c] = acc[0][2][c]; acc[0][2][c] = 0.f; }
  }
  if (R > 0) {
    if (f.state) {
      // partial state of this segment for the four pixels: fields best, idx, before[R], after[R]
      const int OH = 2 * p.H;
      const size_t fplane = (size_t)p.B * OH * OW;
      float* sp = f.state + (size_t)seg * (2 + 2 * R) * fplane + ((size_t)b * OH + 2 * y) * OW + 2 * x;
      constexpr int RR = R > 0 ? R : 1;
#pragma unroll
      for (int k = 0; k < 2 + 2 * RR; ++k) {
        float q[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          q[c] = k == 0 ? state[c].best : (k == 1 ? __int_as_float(state[c].idx)
                        : (k < 2 + RR ? state[c].before[k - 2] : state[c].after[k - 2 - RR]));
        *reinterpret_cast<float2*>(sp + (size_t)k * fplane) = make_float2(q[0], q[1]);
        *reinterpret_cast<float2*>(sp + (size_t)k * fplane + OW) = make_float2(q[2], q[3]);
      }
      return;
    }
    const int Hc = 2 * p.H - f.crop_top, Wc = OW - f.crop_left;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int oy = 2 * y + (c >> 1) - f.crop_top, ox = 2 * x + (c & 1) - f.crop_left;
      if (oy < 0 || ox < 0) continue;
      const size_t o = ((size_t)b * Hc + oy) * Wc + ox;
      f.disparity[o] = state[c].disparity(p.D, f.step);
      if (f.argmax) f.argmax[o] = state[c].idx;
    }
  }
}

// Shared-memory tiled form of the plain (cost-volume writing) kernel.  The per-thread kernel above
// is instruction-issue bound: ~470 instructions per thread and plane for 192 useful FFMAs (border
// predicates, 9x redundant normalisation, and -- sm_100 stages constant operands through uniform
// registers -- one LDCU.128 per four weights).  Here a CTA of 32 x 8 threads covers 64 x 8 input
// columns: each input plane is staged ONCE (66 x 10 voxels with the halo, normalised on the way in,
// zeros outside the volume), every thread owns the two columns tx and tx + 32 (conflict-free
// LDS.128, coalesced stores) and the two pixels share every weight fetch.  The next plane's global
// loads are in flight while the current one is consumed; one __syncthreads per plane.
constexpr int kTileW = 66, kTileH = 10, kTileN = kTileW * kTileH, kTilePer = (kTileN + 255) / 256;

__global__ void __launch_bounds__(256, 2)
hourglass_tail_tiled_kernel(const __grid_constant__ TailParams p) {
  __shared__ float4 tile[2][kTileH][kTileW];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  const int x0 = blockIdx.x * 64, y0 = blockIdx.y * 8, y = y0 + ty;
  const int b = blockIdx.z / p.nseg, seg = blockIdx.z - b * p.nseg;
  const int z0 = seg * p.zseg, z1 = min(p.D, z0 + p.zseg);
  float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.stats) {
    const double n = (double)p.D * p.H * p.W;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double s = p.stats[(b * 4 + c) * 2], q = p.stats[(b * 4 + c) * 2 + 1];
      const double mean = s / n;
      double var = q / n - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + 1e-5));
      sc[c] = rstd * p.gamma[c];
      sh[c] = p.beta[c] - (float)mean * sc[c];
    }
  }
  const size_t plane = (size_t)p.H * p.W;
  const float4* base = p.in + (size_t)b * p.D * plane;
  // the (up to) kTilePer tile elements this thread stages per plane
  int soff[kTilePer];        // offset inside a plane; -1: outside the image (zero), -2: no element
  int sidx[kTilePer];        // float4 index inside one tile buffer
#pragma unroll
  for (int k = 0; k < kTilePer; ++k) {
    const int e = tid + 256 * k;
    const int ly = e / kTileW, lx = e - ly * kTileW, gy = y0 + ly - 1, gx = x0 + lx - 1;
    const bool ok = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
    soff[k] = e < kTileN ? (ok ? gy * p.W + gx : -1) : -2;
    sidx[k] = e < kTileN ? e : 0;
  }
  // fetch only issues the loads (raw values); the normalisation happens in stash, after the
  // current plane has been consumed, so nothing waits on the loads in between
  auto fetch = [&](int zi, float4 (&r)[kTilePer], bool& zok) {
    zok = zi >= 0 && zi < p.D;
#pragma unroll
    for (int k = 0; k < kTilePer; ++k) {
      r[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (zok && soff[k] >= 0) r[k] = __ldg(base + (size_t)zi * plane + soff[k]);
    }
  };
  auto stash = [&](int buf, const float4 (&r)[kTilePer], bool zok) {
    float4* t = &tile[buf][0][0];
#pragma unroll
    for (int k = 0; k < kTilePer; ++k)
      if (soff[k] > -2) {
        const bool live = zok && soff[k] >= 0;              // else: zero padding of the NORMALISED tensor
        t[sidx[k]] = live ? make_float4(fmaf(r[k].x, sc[0], sh[0]), fmaf(r[k].y, sc[1], sh[1]),
                                        fmaf(r[k].z, sc[2], sh[2]), fmaf(r[k].w, sc[3], sh[3]))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
  };
  const int OW = 2 * p.W;
  const size_t oplane = (size_t)4 * plane;
  float* obase = p.out + (size_t)b * p.D * oplane + (size_t)(2 * y) * OW + 2 * (x0 + tx);
  const bool in0 = y < p.H && x0 + tx < p.W, in1 = y < p.H && x0 + 32 + tx < p.W;
  float acc[2][3][4];
#pragma unroll
  for (int px = 0; px < 2; ++px)
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[px][k][c] = 0.f;

  float4 r[kTilePer];
  bool rz;
  fetch(z0 - 1, r, rz);
  stash(0, r, rz);
  __syncthreads();
  int cur = 0;
  for (int zi = z0 - 1; zi <= z1; ++zi) {
    if (zi < z1) fetch(zi + 1, r, rz);                     // in flight while plane zi is consumed
    if (zi >= 0 && zi < p.D) {
      float4 v[2][3];
#define PDS_TAIL_ROW(DY)                                                                      \
      _Pragma("unroll") for (int px = 0; px < 2; ++px)                                         \
      _Pragma("unroll") for (int dx = 0; dx < 3; ++dx) v[px][dx] = tile[cur][ty + DY][tx + 32 * px + dx]; \
      tail_accumulate_row<DY, 2>(acc, v, p.w);
      PDS_TAIL_ROW(0)
      PDS_TAIL_ROW(1)
      PDS_TAIL_ROW(2)
#undef PDS_TAIL_ROW
    }
    const int zo = zi - 1;
    if (zo >= z0 && zo < z1) {
      float* o = obase + (size_t)zo * oplane;
      if (in0) {
        *reinterpret_cast<float2*>(o) = make_float2(acc[0][0][0] + p.bias, acc[0][0][1] + p.bias);
        *reinterpret_cast<float2*>(o + OW) = make_float2(acc[0][0][2] + p.bias, acc[0][0][3] + p.bias);
      }
      if (in1) {
        *reinterpret_cast<float2*>(o + 64) = make_float2(acc[1][0][0] + p.bias, acc[1][0][1] + p.bias);
        *reinterpret_cast<float2*>(o + 64 + OW) = make_float2(acc[1][0][2] + p.bias, acc[1][0][3] + p.bias);
      }
    }
#pragma unroll
    for (int px = 0; px < 2; ++px)
#pragma unroll
      for (int c = 0; c < 4; ++c) { acc[px][0][c] = acc[px][1][c]; acc[px][1][c] = acc[px][2][c]; acc[px][2][c] = 0.f; }
    if (zi < z1) stash(cur ^ 1, r, rz);
    __syncthreads();
    cur ^= 1;
  }
}

// Merges the per-segment SubpixelMap states (lowest index wins ties, the first NaN wins) and
// writes the cropped disparity.  One thread per output pixel.
template <int R>
__global__ void __launch_bounds__(256)
subpixel_merge_kernel(const float* __restrict__ state, float* __restrict__ disparity, int64_t* __restrict__ argmax,
                      int B, int OH, int OW, int D, int nseg, int step, int crop_top, int crop_left) {
  const int Hc = OH - crop_top, Wc = OW - crop_left;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * Hc * Wc) return;
  const int ox = (int)(i % Wc), oy = (int)((i / Wc) % Hc), b = (int)(i / ((size_t)Wc * Hc));
  const size_t fplane = (size_t)B * OH * OW;
  const size_t at = ((size_t)b * OH + oy + crop_top) * OW + ox + crop_left;
  int win = 0;
  float best = state[at];
  for (int s = 1; s < nseg; ++s) {
    const float v = state[(size_t)s * (2 + 2 * R) * fplane + at];
    if (takes_over(v, best)) { best = v; win = s; }
  }
  MapState<R> m;
  const float* sp = state + (size_t)win * (2 + 2 * R) * fplane + at;
  m.best = best;
  m.idx = __float_as_int(sp[fplane]);
#pragma unroll
  for (int r = 0; r < R; ++r) { m.before[r] = sp[(size_t)(2 + r) * fplane]; m.after[r] = sp[(size_t)(2 + R + r) * fplane]; }
  disparity[i] = m.disparity(D, step);
  if (argmax) argmax[i] = m.idx;
}

}  // namespace

// w_host: the layer's weight in PyTorch layout (Cin = 4, Cout = 1, 3, 4, 4), on the HOST.
// disparity != null: fused with SubpixelMap (window radius R = half_support_window / step in
// 1..4) and the SizeAdapter crop; `out` is not written.
size_t hourglass_tail_state_bytes(int B, int D, int H, int W) {
  const int nseg = (D + 11) / 12;      // upper bound over the segment lengths in use
  return align_up((size_t)nseg * (2 + 2 * 4) * B * (2 * H) * (2 * W) * sizeof(float), 256);
}

int hourglass_tail_forward(const float* in, float* out, const double* stats, const float* gamma_host,
                           const float* beta_host, const float* w_host, float bias, int B, int D,
                           int H, int W, cudaStream_t st, float* disparity, int64_t* argmax, int R,
                           int step, int crop_top, int crop_left, float* state) {
  if (B == 0 || D == 0 || H == 0 || W == 0) return PDS_OK;
  TailParams p;
  p.in = reinterpret_cast<const float4*>(in); p.out = out; p.stats = stats;
  p.B = B; p.D = D; p.H = H; p.W = W;
  const bool fused = disparity != nullptr;
  // fused with a state buffer: the disparity axis stays segmented (parallelism), every segment
  // leaves a partial SubpixelMap state and subpixel_merge_kernel finishes; without one a thread
  // scans the whole axis
  const int zseg_env = getenv("PDS_B200_TAIL_ZSEG") ? atoi(getenv("PDS_B200_TAIL_ZSEG")) : 48;
  const int by = getenv("PDS_B200_TAIL_BY") ? atoi(getenv("PDS_B200_TAIL_BY")) : 8;
  p.zseg = (fused && !state) ? D : (D > zseg_env ? zseg_env : D);
  p.nseg = (D + p.zseg - 1) / p.zseg;
  for (int c = 0; c < 4; ++c) { p.gamma[c] = gamma_host ? gamma_host[c] : 1.f; p.beta[c] = beta_host ? beta_host[c] : 0.f; }
  for (int ci = 0; ci < 4; ++ci)
    for (int kd = 0; kd < 3; ++kd)
      for (int kh = 0; kh < 4; ++kh)
        for (int kw = 0; kw < 4; ++kw) p.w[kd][kh][kw][ci] = w_host[((ci * 3 + kd) * 4 + kh) * 4 + kw];
  p.bias = bias;
  FusedParams f;
  f.disparity = disparity; f.argmax = argmax; f.step = step; f.crop_top = crop_top; f.crop_left = crop_left;
  f.state = (fused && p.nseg > 1) ? state : nullptr;
  dim3 grid((unsigned)((W + 31) / 32), (unsigned)((H + by - 1) / by), (unsigned)(B * p.nseg));
  if (grid.z > 65535) { set_error("hourglass_tail: batch too large"); return PDS_ERR_UNSUPPORTED; }
  if (fused && (R < 1 || R > 4 || crop_top < 0 || crop_left < 0 || crop_top > 2 * H || crop_left > 2 * W)) {
    set_error("hourglass_tail: fused estimator needs a window radius in 1..4 and a crop inside the image");
    return PDS_ERR_UNSUPPORTED;
  }
  PDS_KERNEL(fused ? "hourglass_tail+subpixel_map" : "hourglass_tail(tconv 4->1 + IN)", st);
  PDS_KERNEL_WORK(2.0 * 192 * B * D * H * W,
                  (double)B * D * H * W * 16 + (fused ? 4.0 * B * (2 * H - crop_top) * (2 * W - crop_left)
                                                       : (double)B * D * H * W * 16));
  static const bool tiled = !(getenv("PDS_B200_TAIL_TILED") && atoi(getenv("PDS_B200_TAIL_TILED")) == 0);
  switch (fused ? R : 0) {
    case 0:
      if (tiled) hourglass_tail_tiled_kernel<<<dim3((unsigned)((W + 63) / 64), (unsigned)((H + 7) / 8), grid.z), dim3(32, 8), 0, st>>>(p);
      else hourglass_tail_kernel<0><<<grid, dim3(32, by), 0, st>>>(p, f);
      break;
    case 1: hourglass_tail_kernel<1><<<grid, dim3(32, by), 0, st>>>(p, f); break;
    case 2: hourglass_tail_kernel<2><<<grid, dim3(32, by), 0, st>>>(p, f); break;
    case 3: hourglass_tail_kernel<3><<<grid, dim3(32, by), 0, st>>>(p, f); break;
    default: hourglass_tail_kernel<4><<<grid, dim3(32, by), 0, st>>>(p, f); break;
  }
  PDS_LAUNCH_CHECK("hourglass_tail_kernel");
  if (f.state) {
    const int OH = 2 * H, OW = 2 * W;
    const size_t n = (size_t)B * (OH - crop_top) * (OW - crop_left);
    PDS_KERNEL("subpixel_merge", st);
    PDS_KERNEL_WORK(0, (double)n * (4.0 * p.nseg + 4.0 * (2 + 2 * R)));
    const unsigned g = (unsigned)((n + 255) / 256);
    switch (R) {
      case 1: subpixel_merge_kernel<1><<<g, 256, 0, st>>>(state, disparity, argmax, B, OH, OW, D, p.nseg, step, crop_top, crop_left); break;
      case 2: subpixel_merge_kernel<2><<<g, 256, 0, st>>>(state, disparity, argmax, B, OH, OW, D, p.nseg, step, crop_top, crop_left); break;
      case 3: subpixel_merge_kernel<3><<<g, 256, 0, st>>>(state, disparity, argmax, B, OH, OW, D, p.nseg, step, crop_top, crop_left); break;
      default: subpixel_merge_kernel<4><<<g, 256, 0, st>>>(state, disparity, argmax, B, OH, OW, D, p.nseg, step, crop_top, crop_left); break;
    }
    PDS_LAUNCH_CHECK("subpixel_merge_kernel");
  }
  return PDS_OK;
}

}  // namespace pds
