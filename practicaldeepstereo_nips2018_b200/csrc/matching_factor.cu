// Factorisation of the matching operation's first TWO convolutions (reference matching.py:53-62,
// 97-112: conv0 on cat[left, shift_d(right)], then the first convolution of residual block 1).
// Both are linear and the disparity only shifts the right descriptor, so per sample (not per
// disparity slice) we compute
//     A  = conv0_L(L) + b0          Bf = conv0_R(R)           Q = kx=2 taps of conv0_R in place
//     PA = conv1(A) + b1            PB = conv1(Bf)
// and a few single-column corrections; every slice then follows by shifting:
//     x0_d = A + shift_d(B~) - [x = W-1, d >= 1] Q[W-d]
//     conv1(x0_d) + b1 = PA + shift_d(PG~) - [x = W-1] E_d - [x = W-2] F_d          (1 <= d < W)
// B~ is Bf extended by one column on the left (B~[-1] = Q[0]: the column next to the zero fill
// sees R's column 0 through the right-neighbour taps); PG~ = conv1 of B~ on x' in [-2, W-1]:
//     PG~[-2] = U2(Q[:,0])   PG~[-1] = U1(Q[:,0]) + U2(Bf[:,0])   PG~[0] = PB[0] + U0(Q[:,0])   PG~[x'] = PB[x']
// with U_k(c)[y] = sum_{dy,ci} W1[dy][kx=k][ci] c[y+dy-1][ci] (conv1 restricted to one kernel column).
// At the last image columns the shifted tensors would read what the reference pads with zeros:
//     E_d = U2(Bf[:, W-d]) + U1(Q[:, W-d])      F_d = U2(Q[:, W-d])
// (derivation in DESIGN.md 4.1).  d = 0: x0 = A + Bf, conv1 = PA + PB; d >= W: x0 = A, conv1 = PA.
//
// Saves, at C2, one 64->64 convolution over all 48 slices and the pass that materialised x0; x0 is
// re-generated on the fly where the residual addition needs it.  All tensors here are fp32 planes
// [b][C/4][H][W][4]; `cols` is [b][J][H][C] with J = 3 + 2 (D - 1).
#include "conv_tc.cuh"
#include "matching_first.cuh"
#include "tc_ptx.cuh"

namespace pds {
namespace {

using namespace ptx;

// fp32 planes [n][C/4][HW][4] -> split AP planes [n][S][C/8][HW][8]
template <bool FP16, int S>
__global__ void __launch_bounds__(256)
planes_to_ap_kernel(const float* __restrict__ y, uint16_t* __restrict__ out_ap, int C, size_t HW) {
  const int c8 = blockIdx.y, n = blockIdx.z;
  const float4* y4 = reinterpret_cast<const float4*>(y) + ((size_t)n * (C / 4) + 2 * c8) * HW;
  float4* o4 = reinterpret_cast<float4*>(out_ap);
  for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += (size_t)gridDim.x * blockDim.x) {
    const float4 lo = __ldg(y4 + pix), hi = __ldg(y4 + HW + pix);
    const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint16_t t[8][3];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_terms<FP16>(v[e], t[e]);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      float4 pk;
      pk.x = __uint_as_float(t[0][s] | ((uint32_t)t[1][s] << 16)); pk.y = __uint_as_float(t[2][s] | ((uint32_t)t[3][s] << 16));
      pk.z = __uint_as_float(t[4][s] | ((uint32_t)t[5][s] << 16)); pk.w = __uint_as_float(t[6][s] | ((uint32_t)t[7][s] << 16));
      o4[((size_t)(n * S + s) * (C / 8) + c8) * HW + pix] = pk;
    }
  }
}

// (Cout, Cin, 3, 3) -> [kx][dy][ci][co]
__global__ void transpose_w_kernel(const float* __restrict__ w, float* __restrict__ wt, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * C * 9) return;
  const int co = i % C, ci = (i / C) % C, dy = (i / (C * C)) % 3, k = i / (3 * C * C);
  wt[i] = w[((size_t)(co * C + ci) * 3 + dy) * 3 + k];
}

// Column corrections as a small tiled GEMM: one CTA per (64-row chunk, column job j, sample b)
// computes a 64 (y) x 64 (co) tile; K = 3 (dy) x C input channels per term.  The input column
// (with one zero row above / below) and one 64 x 64 weight slab at a time live in shared memory;
// a thread owns 4 rows x 4 channels (one 128-bit and four scalar shared loads per 16 FMAs).
constexpr int kColRows = 64;
__global__ void __launch_bounds__(256)
column_ops_kernel(const float* __restrict__ Bf, const float* __restrict__ Q, const float* __restrict__ wt,
                  float* __restrict__ cols, int C, int H, int W, int D) {
  extern __shared__ float xsm[];      // X [2 terms][kColRows + 2][C + 1] | W slab [C][C]
  const int j = blockIdx.y, b = blockIdx.z, J = gridDim.y;
  const int y0 = blockIdx.x * kColRows;
  const float* src[2] = {nullptr, nullptr};
  int col[2] = {0, 0}, kx[2] = {0, 0};
  if (j == 0) { src[0] = Q; col[0] = 0; kx[0] = 0; }
  else if (j == 1) { src[0] = Q; col[0] = 0; kx[0] = 1; src[1] = Bf; col[1] = 0; kx[1] = 2; }
  else if (j == 2) { src[0] = Q; col[0] = 0; kx[0] = 2; }
  else {
    const int d = (j - 3) / 2 + 1, c = W - d;
    if (c < 0) return;                                   // d > W: never read
    if ((j - 3) % 2 == 0) { src[0] = Bf; col[0] = c; kx[0] = 2; src[1] = Q; col[1] = c; kx[1] = 1; }
    else { src[0] = Q; col[0] = c; kx[0] = 2; }
  }
  const size_t HW = (size_t)H * W;
  const int rows = kColRows + 2, XP = C + 1;
  float* wsm = xsm + 2 * rows * XP;
  for (int t = 0; t < 2; ++t) {
    if (!src[t]) continue;
    const float* xs = src[t] + (size_t)b * C * HW;
    // consecutive threads walk the rows of one channel quad: 16-byte loads, stride W * 16 B
    for (int i = threadIdx.x; i < rows * (C / 4); i += blockDim.x) {
      const int r = i % rows, q = i / rows, yy = y0 - 1 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (yy >= 0 && yy < H) v = __ldg(reinterpret_cast<const float4*>(xs) + (size_t)q * HW + (size_t)yy * W + col[t]);
      float* dst = xsm + (t * rows + r) * XP + 4 * q;
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
  }
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;     // 4 channels x 4 rows per thread
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  for (int t = 0; t < 2; ++t) {
    if (!src[t]) continue;
    for (int dy = 0; dy < 3; ++dy) {
      __syncthreads();                 // previous slab consumed (and, first time, X staged)
      const float4* wsrc = reinterpret_cast<const float4*>(wt + ((size_t)(kx[t] * 3 + dy) * C) * C);
      for (int i = threadIdx.x; i < C * C / 4; i += blockDim.x) reinterpret_cast<float4*>(wsm)[i] = __ldg(wsrc + i);
      __syncthreads();
      const float* xb = xsm + (t * rows + 4 * ty + dy) * XP;
#pragma unroll 4
      for (int ci = 0; ci < C; ++ci) {
        const float4 w = *reinterpret_cast<const float4*>(wsm + ci * C + 4 * tx);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float a = xb[r * XP + ci];
          acc[r][0] = fmaf(w.x, a, acc[r][0]); acc[r][1] = fmaf(w.y, a, acc[r][1]);
          acc[r][2] = fmaf(w.z, a, acc[r][2]); acc[r][3] = fmaf(w.w, a, acc[r][3]);
        }
      }
    }
  }
  float* out = cols + ((size_t)b * J + j) * H * C;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int y = y0 + 4 * ty + r;
    if (y < H) *reinterpret_cast<float4*>(out + (size_t)y * C + 4 * tx) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  }
}

// t1 = LeakyReLU(conv1(x0_d) + b1) for every slice, + its InstanceNorm sums.
// grid (row blocks, C/8, B).  A CTA stages R rows of eight channels of the two per-sample terms
// in shared memory ONCE and produces those rows of ALL D disparity slices from them: a thread
// owns one image column (its left-term values stay in registers), walks d = 0 .. D-1 reading the
// right term at x - d from shared memory, and the kernel is a pure 4 B/element store stream (the
// first version re-read both terms from L2 for every slice and ran at half the store bandwidth).
// InstanceNorm sums: per thread over the R rows, butterfly over the warp (16 shuffles for the 16
// sums), double atomics in shared memory, one global double atomic per (slice, channel) and CTA.
//
// MODE 0: fp32 planes t + sums (one pass; a separate normalisation pass follows).
// MODE 1: sums only, nothing is stored.   MODE 2: the same values are recomputed, normalised with
// the sums of MODE 1 and written straight as split operand planes -- the two-pass form never
// writes or re-reads the fp32 activation (425 MB each way at 960x540, D = 192): the values cost a
// few shared-memory reads and adds, far less than the DRAM round trip they replace.
template <int R, int MODE, bool FP16, int S>
__global__ void __launch_bounds__(256, 2)
compose_second_kernel(const float* __restrict__ PA, const float* __restrict__ PB, const float* __restrict__ cols,
                      const float* __restrict__ bias, float* __restrict__ t, double* stats,
                      const float* __restrict__ gamma, const float* __restrict__ beta,
                      uint16_t* __restrict__ out_ap, int C, int H, int W, int D) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* sA = reinterpret_cast<float4*>(smem_raw);          // [2][R][W]: channels 0-3 | 4-7
  float4* sB = sA + 2 * R * W;                               // [2][R][W]
  double* sstat = reinterpret_cast<double*>(sB + 2 * R * W); // [D][16] sums (MODE 2: [D][16] floats scale | shift)
  float4* sC = reinterpret_cast<float4*>(sstat + D * 16);    // [J][R][2]: border-correction columns
  const int c8 = blockIdx.y, b = blockIdx.z, y0 = blockIdx.x * R, nr = min(R, H - y0), J = 3 + 2 * (D - 1);
  const int tid = threadIdx.x, lane = tid & 31;
  const size_t HW = (size_t)H * W;
  {
    const float4* a4 = reinterpret_cast<const float4*>(PA) + ((size_t)b * (C / 4) + 2 * c8) * HW + (size_t)y0 * W;
    const float4* b4 = reinterpret_cast<const float4*>(PB) + ((size_t)b * (C / 4) + 2 * c8) * HW + (size_t)y0 * W;
    for (int i = tid; i < nr * W; i += 256) {
      sA[i] = __ldg(a4 + i); sA[R * W + i] = __ldg(a4 + HW + i);
      sB[i] = __ldg(b4 + i); sB[R * W + i] = __ldg(b4 + HW + i);
    }
    if (MODE == 2) {
      float* sn = reinterpret_cast<float*>(sstat);
      for (int i = tid; i < D * 8; i += 256) {
        const int d = i >> 3, e = i & 7, c = c8 * 8 + e;
        const double s = stats[((size_t)(b * D + d) * C + c) * 2], q = stats[((size_t)(b * D + d) * C + c) * 2 + 1];
        const double mean = s / (double)HW;
        double var = q / (double)HW - mean * mean;
        if (var < 0.0) var = 0.0;
        const float scale = (float)(1.0 / sqrt(var + 1e-5)) * gamma[c];
        sn[d * 16 + e] = scale;
        sn[d * 16 + 8 + e] = beta[c] - (float)mean * scale;
      }
    } else {
      for (int i = tid; i < D * 16; i += 256) sstat[i] = 0.0;
    }
    const float* cb = cols + (size_t)b * J * H * C + 8 * c8;
    for (int i = tid; i < J * R * 2; i += 256) {
      const int h = i & 1, r = (i >> 1) % R, j = (i >> 1) / R;
      if (r < nr) sC[i] = __ldg(reinterpret_cast<const float4*>(cb + ((size_t)j * H + y0 + r) * C) + h);
    }
  }
  __syncthreads();
  float b1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) b1[e] = __ldg(bias + 8 * c8 + e);
  float4* tb = reinterpret_cast<float4*>(t) + ((size_t)b * D * (C / 4) + 2 * c8) * HW + (size_t)y0 * W;
  float4* o4 = reinterpret_cast<float4*>(out_ap);
  const float4* snorm = reinterpret_cast<const float4*>(sstat);
  for (int x0 = 0; x0 < W; x0 += 256) {
    const int x = x0 + tid;
    const bool valid = x < W;
    float4 alo[R], ahi[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      alo[r] = make_float4(b1[0], b1[1], b1[2], b1[3]); ahi[r] = make_float4(b1[4], b1[5], b1[6], b1[7]);
      if (valid && r < nr) {
        const float4 l = sA[r * W + x], h = sA[(R + r) * W + x];
        alo[r].x += l.x; alo[r].y += l.y; alo[r].z += l.z; alo[r].w += l.w;
        ahi[r].x += h.x; ahi[r].y += h.y; ahi[r].z += h.z; ahi[r].w += h.w;
      }
    }
    for (int d = 0; d < D; ++d) {
      const int xs = x - d;
      const bool shifted = d < W, has_b = valid && shifted && xs >= 0;
      const int jp = !(valid && shifted) ? -1 : (xs == 0 && d >= 1) ? 0 : xs == -1 ? 1 : xs == -2 ? 2 : -1;
      const int jm = !(valid && shifted && d >= 1) ? -1 : x == W - 1 ? 3 + 2 * (d - 1) : x == W - 2 ? 4 + 2 * (d - 1) : -1;
      float4* t4 = tb + (size_t)d * (C / 4) * HW + x;
      float sc[8], sh[8];
      if (MODE == 2) {
        const float4 sc0 = snorm[d * 4], sc1 = snorm[d * 4 + 1], sh0 = snorm[d * 4 + 2], sh1 = snorm[d * 4 + 3];
        sc[0] = sc0.x; sc[1] = sc0.y; sc[2] = sc0.z; sc[3] = sc0.w; sc[4] = sc1.x; sc[5] = sc1.y; sc[6] = sc1.z; sc[7] = sc1.w;
        sh[0] = sh0.x; sh[1] = sh0.y; sh[2] = sh0.z; sh[3] = sh0.w; sh[4] = sh1.x; sh[5] = sh1.y; sh[6] = sh1.z; sh[7] = sh1.w;
      }
      float sum[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) sum[e] = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (!(valid && r < nr)) continue;
        float v[8] = {alo[r].x, alo[r].y, alo[r].z, alo[r].w, ahi[r].x, ahi[r].y, ahi[r].z, ahi[r].w};
        if (has_b) {
          const float4 l = sB[r * W + xs], h = sB[(R + r) * W + xs];
          v[0] += l.x; v[1] += l.y; v[2] += l.z; v[3] += l.w;
          v[4] += h.x; v[5] += h.y; v[6] += h.z; v[7] += h.w;
        }
        if (jp >= 0 || jm >= 0) {                // border columns only
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int j = k == 0 ? jp : jm;
            if (j < 0) continue;
            const float sign = k == 0 ? 1.f : -1.f;
            const float4 l = sC[(j * R + r) * 2], h = sC[(j * R + r) * 2 + 1];
            v[0] = fmaf(sign, l.x, v[0]); v[1] = fmaf(sign, l.y, v[1]); v[2] = fmaf(sign, l.z, v[2]); v[3] = fmaf(sign, l.w, v[3]);
            v[4] = fmaf(sign, h.x, v[4]); v[5] = fmaf(sign, h.y, v[5]); v[6] = fmaf(sign, h.z, v[6]); v[7] = fmaf(sign, h.w, v[7]);
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          v[e] = fmaxf(v[e], 0.1f * v[e]);       // LeakyReLU(0.1)
          if (MODE != 2) { sum[e] += v[e]; sum[8 + e] = fmaf(v[e], v[e], sum[8 + e]); }
        }
        if (MODE == 0) {
          stg_stream(t4 + r * W, make_float4(v[0], v[1], v[2], v[3]));
          stg_stream(t4 + HW + r * W, make_float4(v[4], v[5], v[6], v[7]));
        }
        if (MODE == 2) {
          uint16_t tt[8][3];
#pragma unroll
          for (int e = 0; e < 8; ++e) split_terms<FP16>(fmaf(v[e], sc[e], sh[e]), tt[e]);
#pragma unroll
          for (int s = 0; s < S; ++s) {
            float4 pk;
            pk.x = __uint_as_float(tt[0][s] | ((uint32_t)tt[1][s] << 16)); pk.y = __uint_as_float(tt[2][s] | ((uint32_t)tt[3][s] << 16));
            pk.z = __uint_as_float(tt[4][s] | ((uint32_t)tt[5][s] << 16)); pk.w = __uint_as_float(tt[6][s] | ((uint32_t)tt[7][s] << 16));
            stg_stream(o4 + ((size_t)((b * D + d) * S + s) * (C / 8) + c8) * HW + (size_t)(y0 + r) * W + x, pk);
          }
        }
      }
      if (MODE != 2) {
        const float total = ptx::warp_transpose_reduce<16>(sum, lane);     // lane k (and k + 16): sum k
        if (lane < 16) atomicAdd(&sstat[d * 16 + lane], (double)total);
      }
    }
  }
  if (MODE == 2) return;
  __syncthreads();
  for (int i = tid; i < D * 16; i += 256) {
    const int d = i >> 4, k = i & 15, e = k & 7, which = k >> 3;
    atomicAdd(stats + ((size_t)(b * D + d) * C + 8 * c8 + e) * 2 + which, sstat[i]);
  }
}

// x1 = IN(t2) + x0_d with x0 regenerated from A / Bf / Q -> split AP planes.
// grid (row blocks, C/8, B): like compose_second_kernel, a CTA stages R rows of eight channels of
// the three per-sample terms in shared memory once and walks all D slices of those rows; per
// slice it streams t2 (fp32) in and the operand planes out -- the only DRAM traffic left.  The
// InstanceNorm scale / shift of every slice are computed once per CTA.
template <bool FP16, int S, int R>
__global__ void __launch_bounds__(256, 2)
norm_residual_first_kernel(const float* __restrict__ y, const double* __restrict__ stats, int n_rep, size_t rep_stride,
                           const float* __restrict__ gamma, const float* __restrict__ beta,
                           const float* __restrict__ A, const float* __restrict__ Bf, const float* __restrict__ Q,
                           uint16_t* __restrict__ out_ap, int C, int H, int W, int D) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* sA = reinterpret_cast<float4*>(smem_raw);          // [2][R][W]: channels 0-3 | 4-7
  float4* sB = sA + 2 * R * W;
  float4* sQ = sB + 2 * R * W;
  float4* snorm = sQ + 2 * R * W;                            // [D][4]: scale 0-3, scale 4-7, shift 0-3, shift 4-7
  const int c8 = blockIdx.y, b = blockIdx.z, y0 = blockIdx.x * R, nr = min(R, H - y0);
  const int tid = threadIdx.x;
  const size_t HW = (size_t)H * W;
  {
    const size_t base = ((size_t)b * (C / 4) + 2 * c8) * HW + (size_t)y0 * W;
    const float4* a4 = reinterpret_cast<const float4*>(A) + base;
    const float4* b4 = reinterpret_cast<const float4*>(Bf) + base;
    const float4* q4 = reinterpret_cast<const float4*>(Q) + base;
    for (int i = tid; i < nr * W; i += 256) {
      sA[i] = __ldg(a4 + i); sA[R * W + i] = __ldg(a4 + HW + i);
      sB[i] = __ldg(b4 + i); sB[R * W + i] = __ldg(b4 + HW + i);
      sQ[i] = __ldg(q4 + i); sQ[R * W + i] = __ldg(q4 + HW + i);
    }
    float* sn = reinterpret_cast<float*>(snorm);
    for (int i = tid; i < D * 8; i += 256) {
      const int d = i >> 3, e = i & 7, c = c8 * 8 + e;
      double s = 0.0, q = 0.0;     // n_rep private copies of the sums
      for (int r = 0; r < n_rep; ++r) {
        s += stats[r * rep_stride + ((size_t)(b * D + d) * C + c) * 2];
        q += stats[r * rep_stride + ((size_t)(b * D + d) * C + c) * 2 + 1];
      }
      const double mean = s / (double)HW;
      double var = q / (double)HW - mean * mean;
      if (var < 0.0) var = 0.0;
      const float scale = (float)(1.0 / sqrt(var + 1e-5)) * gamma[c];
      sn[d * 16 + e] = scale;
      sn[d * 16 + 8 + e] = beta[c] - (float)mean * scale;
    }
  }
  __syncthreads();
  float4* o4 = reinterpret_cast<float4*>(out_ap);
  for (int x0 = 0; x0 < W; x0 += 256) {
    const int x = x0 + tid;
    if (x >= W) continue;
    float4 alo[R], ahi[R];
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (r < nr) { alo[r] = sA[r * W + x]; ahi[r] = sA[(R + r) * W + x]; }
    for (int d = 0; d < D; ++d) {
      const int n = b * D + d;
      const float4* y4 = reinterpret_cast<const float4*>(y) + ((size_t)n * (C / 4) + 2 * c8) * HW + (size_t)y0 * W + x;
      float4 lo[R], hi[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (r < nr) { lo[r] = ldg_stream(y4 + r * W); hi[r] = ldg_stream(y4 + HW + r * W); }
      const float4 sc0 = snorm[d * 4], sc1 = snorm[d * 4 + 1], sh0 = snorm[d * 4 + 2], sh1 = snorm[d * 4 + 3];
      const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
      const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
      // x0 of slice d (same arithmetic as tc_compose_first): A + shifted Bf, Q at the two borders
      const int plus = x >= d ? x - d : (x == d - 1 ? 0 : -1);
      const float4* psrc = x >= d ? sB : sQ;
      const int minus = (x == W - 1 && d >= 1 && d <= W) ? W - d : -1;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r >= nr) continue;
        float x0v[8] = {alo[r].x, alo[r].y, alo[r].z, alo[r].w, ahi[r].x, ahi[r].y, ahi[r].z, ahi[r].w};
        if (plus >= 0) {
          const float4 l = psrc[r * W + plus], h = psrc[(R + r) * W + plus];
          x0v[0] += l.x; x0v[1] += l.y; x0v[2] += l.z; x0v[3] += l.w;
          x0v[4] += h.x; x0v[5] += h.y; x0v[6] += h.z; x0v[7] += h.w;
        }
        if (minus >= 0) {
          const float4 l = sQ[r * W + minus], h = sQ[(R + r) * W + minus];
          x0v[0] -= l.x; x0v[1] -= l.y; x0v[2] -= l.z; x0v[3] -= l.w;
          x0v[4] -= h.x; x0v[5] -= h.y; x0v[6] -= h.z; x0v[7] -= h.w;
        }
        const float v[8] = {lo[r].x, lo[r].y, lo[r].z, lo[r].w, hi[r].x, hi[r].y, hi[r].z, hi[r].w};
        uint16_t t[8][3];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_terms<FP16>(fmaf(v[e], sc[e], sh[e]) + x0v[e], t[e]);
#pragma unroll
        for (int s = 0; s < S; ++s) {
          float4 pk;
          pk.x = __uint_as_float(t[0][s] | ((uint32_t)t[1][s] << 16)); pk.y = __uint_as_float(t[2][s] | ((uint32_t)t[3][s] << 16));
          pk.z = __uint_as_float(t[4][s] | ((uint32_t)t[5][s] << 16)); pk.w = __uint_as_float(t[6][s] | ((uint32_t)t[7][s] << 16));
          stg_stream(o4 + ((size_t)(n * S + s) * (C / 8) + c8) * HW + (size_t)(y0 + r) * W + x, pk);
        }
      }
    }
  }
}

}  // namespace

int tc_transpose_weights(const float* w_oihw, float* wt, int C, cudaStream_t st) {
  PDS_KERNEL("tc_transpose_weights", st);
  transpose_w_kernel<<<(unsigned)((C * C * 9 + 255) / 256), 256, 0, st>>>(w_oihw, wt, C);
  PDS_LAUNCH_CHECK("transpose_w_kernel");
  return PDS_OK;
}

int tc_planes_to_ap(const float* y, uint16_t* out_ap, int n, int C, int H, int W, int S, int fp16, cudaStream_t st) {
  const size_t HW = (size_t)H * W;
  if (n == 0 || HW == 0) return PDS_OK;
  unsigned gx = (unsigned)((HW + 255) / 256);
  if (gx > 64) gx = 64;
  dim3 grid(gx, (unsigned)(C / 8), (unsigned)n);
  PDS_KERNEL("tc_planes_to_ap", st);
  PDS_KERNEL_WORK(0, (double)n * C * HW * (4.0 + 2.0 * S));
#define PDS_P2A_CASE(FF, SS) \
  if ((fp16 != 0) == FF && S == SS) planes_to_ap_kernel<FF, SS><<<grid, 256, 0, st>>>(y, out_ap, C, HW);
  PDS_P2A_CASE(true, 1) PDS_P2A_CASE(true, 2) PDS_P2A_CASE(true, 3)
  PDS_P2A_CASE(false, 1) PDS_P2A_CASE(false, 2) PDS_P2A_CASE(false, 3)
#undef PDS_P2A_CASE
  PDS_LAUNCH_CHECK("planes_to_ap_kernel");
  return PDS_OK;
}

size_t tc_column_jobs(int D) { return (size_t)(3 + 2 * (D > 1 ? D - 1 : 0)); }

int tc_column_ops(const float* Bf, const float* Q, const float* wt, float* cols, int B, int C, int H, int W, int D,
                  cudaStream_t st) {
  if (B == 0) return PDS_OK;
  if (C > 256 || C % 4 || 256 % (C / 4)) { set_error("tc_column_ops: unsupported channel count %d", C); return PDS_ERR_UNSUPPORTED; }
  if (C != 64) { set_error("tc_column_ops: 64 channels only"); return PDS_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)((H + kColRows - 1) / kColRows), (unsigned)tc_column_jobs(D), (unsigned)B);
  PDS_KERNEL("tc_column_ops", st);
  PDS_KERNEL_WORK(2.0 * 3 * C * C * H * 1.5 * grid.y * B, 4.0 * grid.y * B * H * C);
  const size_t smem = (size_t)(2 * (kColRows + 2) * (C + 1) + C * C) * sizeof(float);
  PDS_CUDA(allow_dynamic_smem(column_ops_kernel, (int)smem));
  column_ops_kernel<<<grid, 256, smem, st>>>(Bf, Q, wt, cols, C, H, W, D);
  PDS_LAUNCH_CHECK("column_ops_kernel");
  return PDS_OK;
}

namespace {
template <int MODE, bool FP16, int S>
int launch_compose_second(const float* PA, const float* PB, const float* cols, const float* bias, float* t,
                          double* stats, const float* gamma, const float* beta, uint16_t* out_ap, int B, int C,
                          int H, int W, int D, cudaStream_t st) {
  // rows per CTA: as many as leave two CTAs per SM (wide images fall back to fewer rows)
  auto smem_for = [&](int R) {
    return (size_t)4 * R * W * sizeof(float4) + (size_t)D * 16 * sizeof(double) +
           (size_t)(3 + 2 * (D - 1)) * R * 2 * sizeof(float4);
  };
  const int R = smem_for(4) <= 100 * 1024 ? 4 : (smem_for(2) <= 100 * 1024 ? 2 : 1);
  const size_t smem = smem_for(R);
  if (smem > 200 * 1024) { set_error("tc_compose_second: image too wide (%d)", W); return PDS_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)((H + R - 1) / R), (unsigned)(C / 8), (unsigned)B);
#define PDS_COMPOSE_CASE(RR) \
  if (R == RR) {  \
    PDS_CUDA(allow_dynamic_smem(compose_second_kernel<RR, MODE, FP16, S>, (int)smem));  \
    compose_second_kernel<RR, MODE, FP16, S><<<grid, 256, smem, st>>>(PA, PB, cols, bias, t, stats, gamma, beta, out_ap, C, H, W, D);  \
  }
  PDS_COMPOSE_CASE(4) PDS_COMPOSE_CASE(2) PDS_COMPOSE_CASE(1)
#undef PDS_COMPOSE_CASE
  PDS_LAUNCH_CHECK("compose_second_kernel");
  return PDS_OK;
}
}  // namespace

int tc_compose_second(const float* PA, const float* PB, const float* cols, const float* bias, float* t,
                      double* stats, int B, int C, int H, int W, int D, cudaStream_t st) {
  if (B == 0 || (size_t)H * W == 0) return PDS_OK;
  PDS_KERNEL("tc_compose_second", st);
  PDS_KERNEL_WORK(0, (double)B * C * H * W * (8.0 + 4.0 * D));
  return launch_compose_second<0, true, 1>(PA, PB, cols, bias, t, stats, nullptr, nullptr, nullptr, B, C, H, W, D, st);
}

int tc_compose_second_stats(const float* PA, const float* PB, const float* cols, const float* bias, double* stats,
                            int B, int C, int H, int W, int D, cudaStream_t st) {
  if (B == 0 || (size_t)H * W == 0) return PDS_OK;
  PDS_KERNEL("tc_compose_second[sums]", st);
  PDS_KERNEL_WORK(0, (double)B * C * H * W * 8.0);
  return launch_compose_second<1, true, 1>(PA, PB, cols, bias, nullptr, stats, nullptr, nullptr, nullptr, B, C, H, W, D, st);
}

int tc_compose_second_norm(const float* PA, const float* PB, const float* cols, const float* bias,
                           const double* stats, const float* gamma, const float* beta, uint16_t* out_ap, int B,
                           int C, int H, int W, int D, int S, int fp16, cudaStream_t st) {
  if (B == 0 || (size_t)H * W == 0) return PDS_OK;
  PDS_KERNEL("tc_compose_second[norm -> planes]", st);
  PDS_KERNEL_WORK(0, (double)B * C * H * W * (8.0 + 2.0 * S * D));
  double* sp = const_cast<double*>(stats);
#define PDS_CSN_CASE(FF, SS) \
  if ((fp16 != 0) == FF && S == SS) \
    return launch_compose_second<2, FF, SS>(PA, PB, cols, bias, nullptr, sp, gamma, beta, out_ap, B, C, H, W, D, st);
  PDS_CSN_CASE(true, 1) PDS_CSN_CASE(true, 2) PDS_CSN_CASE(true, 3)
  PDS_CSN_CASE(false, 1) PDS_CSN_CASE(false, 2) PDS_CSN_CASE(false, 3)
#undef PDS_CSN_CASE
  set_error("tc_compose_second_norm: unsupported split %d", S);
  return PDS_ERR_UNSUPPORTED;
}

int tc_norm_residual_first(const float* y, const double* stats, const float* gamma, const float* beta,
                           const float* A, const float* Bf, const float* Q, uint16_t* out_ap, int B, int C, int H,
                           int W, int D, int S, int fp16, cudaStream_t st, int n_rep, size_t rep_stride) {
  const size_t HW = (size_t)H * W;
  if (B == 0 || HW == 0) return PDS_OK;
  auto smem_for = [&](int R) { return (size_t)6 * R * W * sizeof(float4) + (size_t)D * 4 * sizeof(float4); };
  const int R = smem_for(4) <= 100 * 1024 ? 4 : (smem_for(2) <= 100 * 1024 ? 2 : 1);
  const size_t smem = smem_for(R);
  if (smem > 200 * 1024) { set_error("tc_norm_residual_first: image too wide (%d)", W); return PDS_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)((H + R - 1) / R), (unsigned)(C / 8), (unsigned)B);
  PDS_KERNEL("tc_norm_residual_first", st);
  PDS_KERNEL_WORK(0, (double)B * D * C * HW * (4.0 + 2.0 * S));
#define PDS_NRF_CASE(FF, SS, RR) \
  if ((fp16 != 0) == FF && S == SS && R == RR) {  \
    PDS_CUDA(allow_dynamic_smem(norm_residual_first_kernel<FF, SS, RR>, (int)smem));  \
    norm_residual_first_kernel<FF, SS, RR><<<grid, 256, smem, st>>>(y, stats, n_rep, rep_stride, gamma, beta, A, Bf, Q, out_ap, C, H, W, D);  \
  }
#define PDS_NRF_ROWS(FF, SS) PDS_NRF_CASE(FF, SS, 4) PDS_NRF_CASE(FF, SS, 2) PDS_NRF_CASE(FF, SS, 1)
  PDS_NRF_ROWS(true, 1) PDS_NRF_ROWS(true, 2) PDS_NRF_ROWS(true, 3)
  PDS_NRF_ROWS(false, 1) PDS_NRF_ROWS(false, 2) PDS_NRF_ROWS(false, 3)
#undef PDS_NRF_ROWS
#undef PDS_NRF_CASE
  PDS_LAUNCH_CHECK("norm_residual_first_kernel");
  return PDS_OK;
}

}  // namespace pds
