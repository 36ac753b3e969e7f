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

// SubpixelMap window of one output pixel (estimator.py:59-91): the maximum with its index, the R
// values before it and the R values after it.
template <int R>
struct MapState {
  float best, before[R], after[R];
  int idx;
  __device__ __forceinline__ void init() {
    best = 0.f; idx = -1;
#pragma unroll
    for (int r = 0; r < R; ++r) { before[r] = 0.f; after[r] = 0.f; }
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

// Per-thread form of the plain (cost-volume writing) kernel (PDS_B200_TAIL_TILED=0; the tiled form
// below is the default): one thread owns one (y, x) column of the input grid.
__global__ void __launch_bounds__(256)
hourglass_tail_kernel(const __grid_constant__ TailParams p) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  const int b = blockIdx.z / p.nseg, seg = blockIdx.z - b * p.nseg;
  const int z0 = seg * p.zseg, z1 = min(p.D, z0 + p.zseg);
  if (x >= p.W || y >= p.H) return;
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
  float* obase = p.out + (size_t)b * p.D * oplane + (size_t)(2 * y) * OW + 2 * x;
  bool okx[3], oky[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) { okx[d] = x + d - 1 >= 0 && x + d - 1 < p.W; oky[d] = y + d - 1 >= 0 && y + d - 1 < p.H; }

  float acc[1][3][4];   // [pixel][out z - zi + 1][cy*2+cx]
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[0][k][c] = 0.f;
  for (int zi = z0 - 1; zi <= z1; ++zi) {
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
    if (zo >= z0 && zo < z1) {
      float* o = obase + (size_t)zo * oplane;
      *reinterpret_cast<float2*>(o) = make_float2(acc[0][0][0] + p.bias, acc[0][0][1] + p.bias);
      *reinterpret_cast<float2*>(o + OW) = make_float2(acc[0][0][2] + p.bias, acc[0][0][3] + p.bias);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[0][0][c] = acc[0][1][c]; acc[0][1][c] = acc[0][2][c]; acc[0][2][c] = 0.f; }
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

// ARGMAX (f1, round 2): the cost volume is NOT written.  Every thread keeps the running maximum and its
// index (th.max semantics) of the eight output pixels it owns -- 16 registers, three instructions per
// value instead of a store -- and leaves (best, index) per pixel and z segment in `state`
// ([segment][2][B][2H][2W]); tail_window_kernel recomputes the few values around the maximum.
template <bool ARGMAX>
__global__ void __launch_bounds__(256, 2)
hourglass_tail_tiled_kernel(const __grid_constant__ TailParams p, float* __restrict__ state) {
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

  // ARGMAX: running maximum / first index of the eight pixels of this thread.  With finite statistics
  // (i.e. a finite input tensor) no value can be NaN short of an overflow, and th.max reduces to
  // "strictly greater takes over": one compare, one select, one FMNMX per value.  Non-finite
  // statistics select the exact form (the first NaN wins and stays).
  float best[2][4];
  int bidx[2][4];
  bool exact = false;
#pragma unroll
  for (int c = 0; c < 4; ++c) exact = exact || !(fabsf(sc[c]) <= 3.0e38f) || !(fabsf(sh[c]) <= 3.0e38f);
#pragma unroll
  for (int px = 0; px < 2; ++px)
#pragma unroll
    for (int c = 0; c < 4; ++c) { best[px][c] = __int_as_float(0xff800000); bidx[px][c] = z0; }

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
    if (ARGMAX) {
      if (zo >= z0 && zo < z1) {
        if (!exact) {
#pragma unroll
          for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float val = acc[px][0][c] + p.bias;
              bidx[px][c] = val > best[px][c] ? zo : bidx[px][c];
              best[px][c] = fmaxf(best[px][c], val);
            }
        } else {
#pragma unroll
          for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float val = acc[px][0][c] + p.bias;
              const bool t = zo == z0 || takes_over(val, best[px][c]);
              best[px][c] = t ? val : best[px][c];
              bidx[px][c] = t ? zo : bidx[px][c];
            }
        }
      }
    } else if (zo >= z0 && zo < z1) {
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
  if (ARGMAX) {
    const int OH = 2 * p.H;
    const size_t fplane = (size_t)p.B * OH * OW;
    float* sp = state + (size_t)seg * 2 * fplane + ((size_t)b * OH + 2 * y) * OW + 2 * (x0 + tx);
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      if (!(px == 0 ? in0 : in1)) continue;
      float* q = sp + 64 * px;
      *reinterpret_cast<float2*>(q) = make_float2(best[px][0], best[px][1]);
      *reinterpret_cast<float2*>(q + OW) = make_float2(best[px][2], best[px][3]);
      *reinterpret_cast<float2*>(q + fplane) = make_float2(__int_as_float(bidx[px][0]), __int_as_float(bidx[px][1]));
      *reinterpret_cast<float2*>(q + fplane + OW) = make_float2(__int_as_float(bidx[px][2]), __int_as_float(bidx[px][3]));
    }
  }
}

// Second half of the fused estimator: one thread per (cropped) output pixel merges the z segments'
// (best, index) -- segments in ascending order, th.max semantics -- and RECOMPUTES the R cost values on
// either side of the maximum: the seven input planes around it, four voxels each, in exactly the
// operation order of tail_accumulate_row (planes ascending, input rows ascending, tx ascending, the
// four channels), so every value is bit-identical to what the plain kernel stores.  Then the windowed
// softmax of SubpixelMap (estimator.py:66-90) and the SizeAdapter crop.
template <int R, int CY, int CX>
__device__ __forceinline__ void tail_window_pixel(const TailParams& p, const float* __restrict__ state, int b, int oy, int ox,
                                                  const float (&sc)[4], const float (&sh)[4], float* __restrict__ disparity,
                                                  int64_t* __restrict__ argmax, size_t out_index, int step) {
  const int OH = 2 * p.H, OW = 2 * p.W;
  const size_t fplane = (size_t)p.B * OH * OW, at = ((size_t)b * OH + oy) * OW + ox;
  MapState<R> m;
  m.init();
  m.best = state[at];
  m.idx = __float_as_int(state[fplane + at]);
  for (int sg = 1; sg < p.nseg; ++sg) {
    const float v = state[(size_t)sg * 2 * fplane + at];
    if (takes_over(v, m.best)) { m.best = v; m.idx = __float_as_int(state[(size_t)sg * 2 * fplane + fplane + at]); }
  }
  const int y = oy >> 1, x = ox >> 1;
  const size_t plane = (size_t)p.H * p.W;
  const float4* base = p.in + (size_t)b * p.D * plane;
  float a[2 * R + 1];            // outputs z = idx - R + k
#pragma unroll
  for (int k = 0; k < 2 * R + 1; ++k) a[k] = 0.f;
  // validity / offsets of the 2 x 2 input voxels this output pixel reads (rows y + CY - 1, y + CY; columns x + CX, x + CX - 1)
  bool ok[2][2];
  int off[2][2];
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int tx = 0; tx < 2; ++tx) {
      const int gy = y + CY + dy - 1, gx = x + CX - tx;
      ok[dy][tx] = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
      off[dy][tx] = gy * p.W + gx;
    }
#pragma unroll
  for (int j = -R - 1; j <= R + 1; ++j) {                 // input plane idx + j; every index below is a compile-time constant
    const int zi = m.idx + j;
    if (zi < 0 || zi >= p.D) continue;                    // the plain kernel skips these planes as well
    float4 q[2][2];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int tx = 0; tx < 2; ++tx) {
        q[dy][tx] = make_float4(0.f, 0.f, 0.f, 0.f);      // zero padding of the NORMALISED tensor
        if (ok[dy][tx]) {
          const float4 t = __ldg(base + (size_t)zi * plane + off[dy][tx]);
          q[dy][tx] = make_float4(fmaf(t.x, sc[0], sh[0]), fmaf(t.y, sc[1], sh[1]), fmaf(t.z, sc[2], sh[2]), fmaf(t.w, sc[3], sh[3]));
        }
      }
#pragma unroll
    for (int k = 0; k < 2 * R + 1; ++k) {
      if (k == R) continue;                               // the maximum itself comes from the first kernel
      const int tz = j - (k - R) + 1;                     // out z = zi - tz + 1 (compile-time after unrolling)
      if (tz < 0 || tz > 2) continue;
      float v = a[k];
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)                      // DY = CY + dy: ty = 1 - dy, kernel row 1 - CY + 2 * ty
#pragma unroll
        for (int tx = 0; tx < 2; ++tx) {
          const float* ww = p.w[2 - tz][1 - CY + 2 * (1 - dy)][1 - CX + 2 * tx];
          v = fmaf(q[dy][tx].x, ww[0], v); v = fmaf(q[dy][tx].y, ww[1], v);
          v = fmaf(q[dy][tx].z, ww[2], v); v = fmaf(q[dy][tx].w, ww[3], v);
        }
      a[k] = v;
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) { m.before[r] = a[R - 1 - r] + p.bias; m.after[r] = a[R + 1 + r] + p.bias; }
  disparity[out_index] = m.disparity(p.D, step);
  if (argmax) argmax[out_index] = m.idx;
}

// grid (input-column blocks, output rows, B): a thread owns the two output pixels (oy, 2x) and (oy, 2x + 1); the
// row parity is uniform per CTA, the column parity a compile-time constant, so every weight is a
// constant-bank operand (with run-time parities the 192 weights went through local memory: 70 us).
template <int R>
__global__ void __launch_bounds__(128)
tail_window_kernel(const __grid_constant__ TailParams p, const float* __restrict__ state, float* __restrict__ disparity,
                   int64_t* __restrict__ argmax, int step, int crop_top, int crop_left) {
  const int OW = 2 * p.W, Hc = 2 * p.H - crop_top, Wc = OW - crop_left;
  const int b = blockIdx.z, oy = crop_top + blockIdx.y;
  const int x = (crop_left >> 1) + blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= p.W) return;
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
  const size_t orow = ((size_t)b * Hc + blockIdx.y) * Wc;
  const int ox0 = 2 * x, ox1 = 2 * x + 1;
  if (oy & 1) {
    if (ox0 >= crop_left) tail_window_pixel<R, 1, 0>(p, state, b, oy, ox0, sc, sh, disparity, argmax, orow + (ox0 - crop_left), step);
    if (ox1 >= crop_left) tail_window_pixel<R, 1, 1>(p, state, b, oy, ox1, sc, sh, disparity, argmax, orow + (ox1 - crop_left), step);
  } else {
    if (ox0 >= crop_left) tail_window_pixel<R, 0, 0>(p, state, b, oy, ox0, sc, sh, disparity, argmax, orow + (ox0 - crop_left), step);
    if (ox1 >= crop_left) tail_window_pixel<R, 0, 1>(p, state, b, oy, ox1, sc, sh, disparity, argmax, orow + (ox1 - crop_left), step);
  }
}

}  // namespace

// w_host: the layer's weight in PyTorch layout (Cin = 4, Cout = 1, 3, 4, 4), on the HOST.
// disparity != null: fused with SubpixelMap (window radius R = half_support_window / step in
// 1..4) and the SizeAdapter crop; `out` is not written.
size_t hourglass_tail_state_bytes(int B, int D, int H, int W) {
  const int nseg = (D + 11) / 12;      // upper bound over the segment lengths in use (PDS_B200_TAIL_ZSEG >= 12)
  return align_up((size_t)nseg * 2 * B * (2 * H) * (2 * W) * sizeof(float), 256);
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
  int zseg_env = getenv("PDS_B200_TAIL_ZSEG") ? atoi(getenv("PDS_B200_TAIL_ZSEG")) : 48;
  if (zseg_env < 12) zseg_env = 12;     // hourglass_tail_state_bytes sizes the state for segments of >= 12
  const int by = getenv("PDS_B200_TAIL_BY") ? atoi(getenv("PDS_B200_TAIL_BY")) : 8;
  p.zseg = D > zseg_env ? zseg_env : D;
  p.nseg = (D + p.zseg - 1) / p.zseg;
  for (int c = 0; c < 4; ++c) { p.gamma[c] = gamma_host ? gamma_host[c] : 1.f; p.beta[c] = beta_host ? beta_host[c] : 0.f; }
  for (int ci = 0; ci < 4; ++ci)
    for (int kd = 0; kd < 3; ++kd)
      for (int kh = 0; kh < 4; ++kh)
        for (int kw = 0; kw < 4; ++kw) p.w[kd][kh][kw][ci] = w_host[((ci * 3 + kd) * 4 + kh) * 4 + kw];
  p.bias = bias;
  dim3 grid((unsigned)((W + 31) / 32), (unsigned)((H + by - 1) / by), (unsigned)(B * p.nseg));
  if (grid.z > 65535) { set_error("hourglass_tail: batch too large"); return PDS_ERR_UNSUPPORTED; }
  if (fused && (R < 1 || R > 4 || crop_top < 0 || crop_left < 0 || crop_top > 2 * H || crop_left > 2 * W || !state)) {
    set_error("hourglass_tail: fused estimator needs a window radius in 1..4, a crop inside the image and a state buffer");
    return PDS_ERR_UNSUPPORTED;
  }
  const dim3 tgrid((unsigned)((W + 63) / 64), (unsigned)((H + 7) / 8), grid.z);
  if (fused) {
    // f1: arg-max inside the transposed convolution, window recomputed (no cost volume in HBM)
    {
      PDS_KERNEL("hourglass_tail[arg-max, no volume]", st);
      PDS_KERNEL_WORK(2.0 * 192 * B * D * H * W, (double)B * D * H * W * 16 + 8.0 * p.nseg * B * 4 * H * W);
      hourglass_tail_tiled_kernel<true><<<tgrid, dim3(32, 8), 0, st>>>(p, state);
      PDS_LAUNCH_CHECK("hourglass_tail_tiled_kernel");
    }
    const size_t n = (size_t)B * (2 * H - crop_top) * (2 * W - crop_left);
    if (n == 0) return PDS_OK;
    PDS_KERNEL("tail_window(subpixel_map)", st);
    PDS_KERNEL_WORK(2.0 * 48 * 2 * R * n, (double)n * (8.0 * p.nseg + 4.0 + 16.0 * (2 * R + 3)));
    const int cols = W - (crop_left >> 1);
    const dim3 g((unsigned)((cols + 127) / 128), (unsigned)(2 * H - crop_top), (unsigned)B);
    if (g.y > 65535 || g.z > 65535) { set_error("hourglass_tail: image too tall / batch too large for the window kernel"); return PDS_ERR_UNSUPPORTED; }
    switch (R) {
      case 1: tail_window_kernel<1><<<g, 128, 0, st>>>(p, state, disparity, argmax, step, crop_top, crop_left); break;
      case 2: tail_window_kernel<2><<<g, 128, 0, st>>>(p, state, disparity, argmax, step, crop_top, crop_left); break;
      case 3: tail_window_kernel<3><<<g, 128, 0, st>>>(p, state, disparity, argmax, step, crop_top, crop_left); break;
      default: tail_window_kernel<4><<<g, 128, 0, st>>>(p, state, disparity, argmax, step, crop_top, crop_left); break;
    }
    PDS_LAUNCH_CHECK("tail_window_kernel");
    return PDS_OK;
  }
  PDS_KERNEL("hourglass_tail(tconv 4->1 + IN)", st);
  PDS_KERNEL_WORK(2.0 * 192 * B * D * H * W, (double)B * D * H * W * 16 + (double)B * D * H * W * 16);
  static const bool tiled = !(getenv("PDS_B200_TAIL_TILED") && atoi(getenv("PDS_B200_TAIL_TILED")) == 0);
  if (tiled) hourglass_tail_tiled_kernel<false><<<tgrid, dim3(32, 8), 0, st>>>(p, nullptr);
  else hourglass_tail_kernel<<<grid, dim3(32, by), 0, st>>>(p);
  PDS_LAUNCH_CHECK("hourglass_tail_kernel");
  return PDS_OK;
}

}  // namespace pds
