// tcgen05 implicit-GEMM 3x3 convolution (stride 1, zero padding 1) for the
// matching operation (reference matching.py:69-112; Conv2d semantics of
// network_blocks.py:19-24, 47-58), sm_100a only.
//
// GEMM view per disparity slice: D[pixel, cout] = sum over (tap, cin) of
// A[pixel + tap, cin] * W[tap, cin, cout].
//   * One MMA tile = 128 pixels = 8 (x) by 16 (y); a CTA tile is NT such tiles
//     side by side sharing ONE haloed input tile in shared memory:
//     (8*NT + 2) x 18 pixels per 16-channel K chunk, loaded once by TMA (OOB zero
//     fill == the convolution's zero padding; for Matching's shifted right
//     descriptor the box is simply placed at x - d, matching.py:56-60).
//   * Operands are K-major, no-swizzle UMMA tiles.  Activations live in HBM as
//     [plane = 8-channel group][y][x][8] 16-bit so that a TMA box lands in shared
//     memory as [plane][y][x][16 B]: 8 consecutive pixels x 16 B is exactly one
//     UMMA core matrix, the next image row is the next core-matrix group
//     (SBO = halo pitch), the second 8-channel group is LBO away -- and a tap
//     (dy, dx) is nothing but a start-address offset of (dy*pitch + dx)*16 B.
//     No im2col, no per-tap reload.
//   * fp32 accuracy on 16-bit tensor cores: x = t0 + t1 (+ t2).  The B operand of
//     a (chunk, tap) holds the S weight terms CONCATENATED ALONG N, so one MMA
//     with activation term s and N = 64*(S-s) produces the products
//     a_s*w_0 .. a_s*w_{S-1-s} into S-s adjacent 64-column accumulators:
//     accumulator k collects every product of order of magnitude k.  Measured
//     (tools/mma_microbench.cu): an M128 N64 MMA is operand-fetch bound (48
//     cycles), N >= 128 runs at the N/2-cycle math floor -- concatenation is what
//     makes the split affordable, and the per-order accumulators keep the
//     tensor core's truncating accumulation away from the small terms.
//   * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread, fully
//     unrolled issue loop: descriptors advance by compile-time constants),
//     warps 2-5 = epilogue (TMEM -> registers -> sum of the accumulators, bias,
//     LeakyReLU, InstanceNorm partial sums -> HBM).  Accumulators are
//     double-buffered in TMEM; CTAs are persistent (grid = #SMs).  When the whole
//     weight tensor fits (64 -> 64 at S = 2: 147 KB) it is loaded ONCE per CTA and
//     stays resident in shared memory.
//   * FUSE = true (the 64 -> 64 layers over all disparity slices): the InstanceNorm that follows
//     the convolution (network_blocks.py:47-58) runs INSIDE the same launch.  Tiles are handed out
//     in slice order by an atomic counter, every finished tile bumps its slice's counter, and
//     four extra warps per CTA pick up slices whose last tile has retired -- they read the fp32
//     activation while it is still in L2, apply IN (+ the residual of the block), and write the
//     next convolution's operand planes.  The separate normalisation passes (HBM-bound, ~0.15 ms
//     each at 960x540 D=192) disappear behind the tensor-core time.  Nothing waits on a CTA that
//     is not resident (tiles and work items are claimed, not assigned), so launches on several
//     streams cannot dead-lock each other.
#include <stdlib.h>

#include <string>

#include "conv_tc.cuh"
#include "matching_first.cuh"
#include "tc_ptx.cuh"

namespace pds {
namespace {

using namespace ptx;

constexpr int kPH = 18;          // halo rows of a CTA tile (16 + 2)
constexpr int kThreads = 192;    // TMA producer warp, MMA issuer warp, four epilogue warps
// FUSE kernels: 16 warps (512 threads x 128 registers = the SM's register file)
//   warp 0 TMA producer, warp 1 MMA issuer, warp 2 progress warp (publishes finished slices)
//   warps 4-11 eight epilogue warps, two per TMEM lane quarter
//   warps 3, 12-15 normalisation warps: finished slices are normalised behind the convolution
// (setmaxnreg was tried to move registers from the control warps to the normalisation warps: ptxas
// honours it, but with the normalisation code inlined -- a prerequisite -- it spills more than it gains)
constexpr int kFusedEpilogueWarp0 = 4, kFusedEpilogueWarps = 8;
constexpr int kProgressWarp = 2;
constexpr int kNormWarp0 = 12, kNormWarps = 5;      // warp 3 is the fifth
constexpr int kFusedThreads = 32 * 16;
constexpr int kNormSliceClasses = 2;   // a normalisation warp serves every second slice
constexpr int kStatReplicas = 8;       // one private copy of a slice's sums per epilogue warp
constexpr int kNormPixels = 128;       // pixels of one 8-channel group per normalisation unit
constexpr int kMaxStages = 6;

struct alignas(64) TcKernelParams {
  CUtensorMap map0;              // primary input (AP planes)
  const CUtensorMap* maps_d;     // [d]: right descriptors at disparity d (first convolution) or null
  const uint16_t* w;             // [chunk][tap][2][S][N][8]
  const float* bias;             // [N]
  float* out_f32;
  uint16_t* out_ap;
  float* out_sig;
  double* stats;
  int n_slices, n0, n_div, H, W, tiles_x, tiles_y;
  int nchunks, nchunks1;             // chunks [0, nchunks1) come from map0
  int planes1, planes2;              // 8-channel planes per (slice, term) of the two inputs
  int in_global;                     // primary input indexed by sample (first convolution)
  int Cout, epilogue, fp16;
  int stages;
  uint32_t stage_bytes, wres_bytes;
  float inv_wscale;
  // ---- FUSE kernels: dynamic tile scheduler + trailing InstanceNorm of the launch's own output ----
  // sched[0]: next tile to claim; sched[1 + s]: tiles of local slice s whose output and sums are
  // stored.  Zeroed before the launch.
  int* sched;
  double* stats_rep;                 // [kStatReplicas][n_slices][N][2] private copies of the sums (zeroed)
  const float* gamma;                // InstanceNorm affine of this block
  const float* beta;
  int norm_mode;                     // TC_NORM_*
  const uint16_t* res_ap;            // TC_NORM_RESIDUAL: residual planes (may alias norm_out)
  uint16_t* norm_out;                // operand planes [n][S][N/8][H][W][8] written by the norm warps
  const float* fA;                   // TC_NORM_RESIDUAL_FIRST: per-sample terms the residual x0 is rebuilt from
  const float* fB;
  const float* fQ;
  long long* trace;                  // debug (build with -DPDS_TC_TRACE, run with PDS_B200_TC_TRACE=1): issue-region stamps of CTA 0, [n][2] after a count
  int continuous;                    // static schedule: one elected lane issues without leaving the stream (PDS_B200_TC_CONTINUOUS)
};

constexpr uint32_t pow2_cols(uint32_t c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }
__host__ __device__ constexpr bool tc_flat_rows(int PW) { return 8 * PW <= 256; }

// shared memory after the weight / stage buffers: barriers, TMEM slot, bias; FUSE adds the tile
// ids in flight and the 1 KB exchange buffer of the cross-warp InstanceNorm sums
constexpr int kNumBars = 2 * kMaxStages + 5 + 2 + 4;
constexpr size_t kTailFixed = 8 * kNumBars + 32 + 64 * sizeof(float);      // barriers | tmem slot | bias
constexpr size_t kTailFused = 48;                                          // tile ids in flight | slices to publish
constexpr size_t tail_bytes(bool fuse) { return kTailFixed + (fuse ? kTailFused : 0) + 128; }

// Bounded spin on a global progress counter (a protocol bug traps after ~4 s instead of hanging).
#ifndef PDS_FUSE_POLL_NS
#define PDS_FUSE_POLL_NS 200
#endif
// The spin uses RELAXED loads (an acquire load invalidates the SM's whole L1 on every poll) and one
// acquire fence once the counter has arrived.
__device__ __forceinline__ void wait_counter(const int* ctr, int target) {
  if (ld_relaxed_gpu(ctr) < target) {
    const unsigned long long t0 = global_ns();
    uint32_t spins = 0;
    while (ld_relaxed_gpu(ctr) < target) {
      __nanosleep(PDS_FUSE_POLL_NS);
      if ((++spins & 0xff) == 0 && global_ns() - t0 > 4000000000ull) __trap();
    }
  }
  fence_acq_rel_gpu();
}

// ---- trailing normalisation (FUSE kernels) ---------------------------------------------------
// A unit = kNormPixels pixels (two per lane) of one 8-channel group of one slice:
//   out = IN(t) [+ residual planes | + x0 rebuilt from A / Bf / Q]  ->  split operand planes,
// the arithmetic of tc_norm_split_kernel / norm_residual_first_kernel (bit-identical results up
// to the order of the double-precision sum atomics).  fp16 terms only (the default precision).
// A warp walks a contiguous range of units of one slice with the loads of unit u + 1 in flight
// while unit u is converted (the warps live on memory-level parallelism: ~1 us to L2 and back).
// The code is kept small and scalar: it shares the 128 registers per thread of a 512-thread CTA.

// InstanceNorm scale / shift of eight channels from the slice's sums (the kStatReplicas private
// copies the epilogue warps add into are summed here, in a fixed order).
__device__ __noinline__ void norm_coefficients(const TcKernelParams& p, int nl, int c8, int lane, float (&a)[8],
                                               float (&b)[8]) {
  constexpr int C = 64;
  const size_t HW = (size_t)p.H * p.W;
  const int c = c8 * 8 + (lane & 7);
  const double* st = p.stats_rep + ((size_t)nl * C + c) * 2;
  const size_t rep = (size_t)p.n_slices * C * 2;
  double s = 0.0, q = 0.0;
#pragma unroll
  for (int r = 0; r < kStatReplicas; ++r) { s += __ldcg(st + r * rep); q += __ldcg(st + r * rep + 1); }
  const double mean = s / (double)HW;
  double var = q / (double)HW - mean * mean;
  if (var < 0.0) var = 0.0;
  const float scale = (float)(1.0 / sqrt(var + 1e-5)) * __ldg(p.gamma + c);
  const float shift = __ldg(p.beta + c) - (float)mean * scale;
#pragma unroll
  for (int e = 0; e < 8; ++e) { a[e] = __shfl_sync(0xffffffffu, scale, e); b[e] = __shfl_sync(0xffffffffu, shift, e); }
}

// v (8 channels of one pixel, normalised [+ residual]) -> S operand planes
template <int S>
__device__ __forceinline__ void norm_emit(uint4* o, size_t term, const float (&v)[8]) {
  uint16_t t[8][3];
#pragma unroll
  for (int e = 0; e < 8; ++e) split_terms<true>(v[e], t[e]);
#pragma unroll
  for (int s = 0; s < S; ++s) {
    uint4 pk;
    pk.x = t[0][s] | ((uint32_t)t[1][s] << 16); pk.y = t[2][s] | ((uint32_t)t[3][s] << 16);
    pk.z = t[4][s] | ((uint32_t)t[5][s] << 16); pk.w = t[6][s] | ((uint32_t)t[7][s] << 16);
    __stcg(o + (size_t)s * term, pk);
  }
}

template <int S, bool RES>
struct NormPixel {
  float4 lo, hi;
  uint4 rr[RES ? S : 1];
};

// TC_NORM_PLAIN / TC_NORM_RESIDUAL: units [u0, u1) of slice nl; unit = c8 * blocks + pixel block of
// kNormPixels pixels; a lane owns SUB x U pixels of it, U at a time, the next U in flight
template <int S, bool RES, int U>
__device__ __noinline__ void norm_range_planes(const TcKernelParams& p, int nl, int u0, int u1, int blocks, int lane) {
  constexpr int C = 64, SUB = kNormPixels / (32 * U);
  const size_t HW = (size_t)p.H * p.W, term = (size_t)(C / 8) * HW;
  const float4* ybase = reinterpret_cast<const float4*>(p.out_f32) + (size_t)nl * (C / 4) * HW;
  const uint4* rbase = reinterpret_cast<const uint4*>(p.res_ap) + (size_t)nl * S * term;
  uint4* obase = reinterpret_cast<uint4*>(p.norm_out) + (size_t)nl * S * term;
  float a[8], b[8];
  int cur = -1;
  NormPixel<S, RES> nx[U];      // the sub-unit in flight
  auto issue = [&](int sub) {   // sub-unit index = unit * SUB + part
    const int unit = sub / SUB, part = sub - unit * SUB;
    const int c8 = unit / blocks, pb = unit - c8 * blocks;
    const size_t pix = (size_t)pb * kNormPixels + part * (32 * U) + lane;
    const float4* y4 = ybase + (size_t)(2 * c8) * HW;
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (pix + 32 * k < HW) {   // written by other SMs during this launch: L2 loads, never .nc / L1
        nx[k].lo = __ldcg(y4 + pix + 32 * k); nx[k].hi = __ldcg(y4 + HW + pix + 32 * k);
        if (RES) {
#pragma unroll
          for (int s = 0; s < S; ++s) nx[k].rr[s] = __ldcg(rbase + (size_t)s * term + (size_t)c8 * HW + pix + 32 * k);
        }
      }
    }
  };
  auto convert = [&](const NormPixel<S, RES>& f, uint4* o) {
    float v[8] = {f.lo.x, f.lo.y, f.lo.z, f.lo.w, f.hi.x, f.hi.y, f.hi.z, f.hi.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], a[e], b[e]);
    if (RES) {
      float res[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int s = S - 1; s >= 0; --s) {   // smallest term first
        const uint32_t w[4] = {f.rr[s].x, f.rr[s].y, f.rr[s].z, f.rr[s].w};
#pragma unroll
        for (int e = 0; e < 8; ++e) res[e] += term_value<true>((uint16_t)(w[e >> 1] >> (16 * (e & 1))));
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += res[e];
    }
    norm_emit<S>(o, term, v);
  };
  const int s0 = u0 * SUB, s1 = u1 * SUB;
  issue(s0);
#pragma unroll 1
  for (int sub = s0; sub < s1; ++sub) {
    NormPixel<S, RES> c[U];
#pragma unroll
    for (int k = 0; k < U; ++k) c[k] = nx[k];
    if (sub + 1 < s1) issue(sub + 1);
    const int unit = sub / SUB, part = sub - unit * SUB;
    const int c8 = unit / blocks, pb = unit - c8 * blocks;
    if (c8 != cur) { norm_coefficients(p, nl, c8, lane, a, b); cur = c8; }
    const size_t pix = (size_t)pb * kNormPixels + part * (32 * U) + lane;
    uint4* o = obase + (size_t)c8 * HW + pix;
#pragma unroll
    for (int k = 0; k < U; ++k)
      if (pix + 32 * k < HW) convert(c[k], o + 32 * k);
  }
}

// TC_NORM_RESIDUAL_FIRST: x0_d = A + shift_d(B~) - [x = W-1, d >= 1] Q[W-d] (matching_first.cuh) is rebuilt
// per pixel; A and B~ are fetched with the activation, the Q column of the last image column on demand
template <int S>
__device__ __noinline__ void norm_range_first(const TcKernelParams& p, int nl, int u0, int u1, int blocks, int lane) {
  constexpr int C = 64;
  const size_t HW = (size_t)p.H * p.W, term = (size_t)(C / 8) * HW;
  const int ng = p.n0 + nl, bs = ng / p.n_div, d = ng - bs * p.n_div, W = p.W;
  const float4* ybase = reinterpret_cast<const float4*>(p.out_f32) + (size_t)nl * (C / 4) * HW;
  uint4* obase = reinterpret_cast<uint4*>(p.norm_out) + (size_t)nl * S * term;
  const size_t fb = (size_t)bs * (C / 4) * HW;
  float a[8], b[8];
  int cur = -1;
#pragma unroll 1
  for (int sub = u0 * (kNormPixels / 64); sub < u1 * (kNormPixels / 64); ++sub) {   // 64 pixels (two per lane) at a time
    const int u = sub / (kNormPixels / 64), part = sub - u * (kNormPixels / 64);
    const int c8 = u / blocks, pb = u - c8 * blocks;
    const size_t plane = (size_t)(2 * c8) * HW;
    const float4* y4 = ybase + plane;
    const float4* fa4 = reinterpret_cast<const float4*>(p.fA) + fb + plane;
    const float4* fb4 = reinterpret_cast<const float4*>(p.fB) + fb + plane;
    const float4* fq4 = reinterpret_cast<const float4*>(p.fQ) + fb + plane;
    float4 lo[2], hi[2], xa[2][4];
    int xc[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const size_t pix = (size_t)pb * kNormPixels + part * 64 + lane + 32 * k;
      if (pix < HW) {
        lo[k] = __ldcg(y4 + pix); hi[k] = __ldcg(y4 + HW + pix);
        const int x = (int)(pix % (size_t)W);
        xc[k] = x;
        xa[k][0] = __ldg(fa4 + pix); xa[k][1] = __ldg(fa4 + HW + pix);
        if (x >= d || x == d - 1) {
          const float4* src = x >= d ? fb4 + (pix - d) : fq4 + (pix - x);
          xa[k][2] = __ldg(src); xa[k][3] = __ldg(src + HW);
        }
      }
    }
    if (c8 != cur) { norm_coefficients(p, nl, c8, lane, a, b); cur = c8; }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const size_t pix = (size_t)pb * kNormPixels + part * 64 + lane + 32 * k;
      if (pix >= HW) continue;
      const int x = xc[k];
      // same operation order as first_x0: A, then + B~, then - Q at the last column
      float x0v[8] = {xa[k][0].x, xa[k][0].y, xa[k][0].z, xa[k][0].w, xa[k][1].x, xa[k][1].y, xa[k][1].z, xa[k][1].w};
      if (x >= d || x == d - 1) {
        const float t2[8] = {xa[k][2].x, xa[k][2].y, xa[k][2].z, xa[k][2].w, xa[k][3].x, xa[k][3].y, xa[k][3].z, xa[k][3].w};
#pragma unroll
        for (int e = 0; e < 8; ++e) x0v[e] = fmaf(1.f, t2[e], x0v[e]);
      }
      if (x == W - 1 && d >= 1 && d <= W) {
        const float4 l = __ldg(fq4 + (pix - x) + (W - d)), h = __ldg(fq4 + HW + (pix - x) + (W - d));
        const float t3[8] = {l.x, l.y, l.z, l.w, h.x, h.y, h.z, h.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) x0v[e] = fmaf(-1.f, t3[e], x0v[e]);
      }
      float v[8] = {lo[k].x, lo[k].y, lo[k].z, lo[k].w, hi[k].x, hi[k].y, hi[k].z, hi[k].w};
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], a[e], b[e]) + x0v[e];
      norm_emit<S>(obase + (size_t)c8 * HW + pix, term, v);
    }
  }
}

template <int S>
__device__ __forceinline__ void norm_range(const TcKernelParams& p, int nl, int u0, int u1, int blocks, int lane) {
  if (u0 >= u1) return;
  if (p.norm_mode == TC_NORM_PLAIN) norm_range_planes<S, false, 4>(p, nl, u0, u1, blocks, lane);
  else if (p.norm_mode == TC_NORM_RESIDUAL) norm_range_planes<S, true, 2>(p, nl, u0, u1, blocks, lane);
  else norm_range_first<S>(p, nl, u0, u1, blocks, lane);
}

// S terms, NT MMA tiles per CTA tile, N rows per weight term, WRES: weights resident, FUSE: dynamic
// tile scheduler + trailing normalisation warps (N = 64, TC_EPI_ACT only).
template <int S, int NT, int N, bool WRES, bool FUSE>
__global__ void __launch_bounds__(FUSE ? kFusedThreads : kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ TcKernelParams p) {
  constexpr int PW = 8 * NT + 2;                          // halo pitch in pixels
  constexpr uint32_t A_TERM_BYTES = 2 * kPH * PW * 16;    // one term of one 16-channel chunk
  constexpr uint32_t W_CHUNK_BYTES = 9 * 2 * S * N * 16;  // all taps and terms of one chunk
  constexpr uint32_t ACC_COLS = S * N;                    // accumulators of one MMA tile
  constexpr uint32_t BUF_COLS = NT * ACC_COLS;
  constexpr int NBUF = 2 * BUF_COLS <= 512 ? 2 : 1;
  constexpr uint32_t TMEM_COLS = pow2_cols(NBUF * BUF_COLS);
  static_assert(BUF_COLS <= 512, "accumulators do not fit TMEM");
  static_assert(!FUSE || (N == 64 && NT % 2 == 0), "the fused normalisation handles 64 output channels, two epilogue warps per lane quarter");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  const uint32_t wres_base = smem_u32(smem);
  const uint32_t stage_base = wres_base + p.wres_bytes;
  const uint32_t bar_base = stage_base + (uint32_t)p.stages * p.stage_bytes;
  // barriers: full[stages], empty[stages], tfull[2], tempty[2], wfull, tid[2]; then tmem pointer; then bias
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + 2 + a); };
  const uint32_t wfull_bar = bar_base + 8u * (2 * kMaxStages + 4);
  auto tid_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + 5 + a); };
  auto sig_full_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + 7 + a); };
  auto sig_empty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + 9 + a); };
  unsigned char* tail = smem + p.wres_bytes + (size_t)p.stages * p.stage_bytes + 8 * kNumBars;
  uint32_t* tmem_slot = (uint32_t*)(tail + 8);
  float* sbias = (float*)(tail + 32);
  volatile int* stage_tile = (volatile int*)(tail + 32 + 64 * sizeof(float));   // [kMaxStages] tile id carried by a stage
  volatile int* acc_tile = stage_tile + kMaxStages;                            // [2] tile id of an accumulator buffer
  volatile int* sig_slice = acc_tile + 2;                                      // [2] slice whose tile has just been stored

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), FUSE ? 32 * kFusedEpilogueWarps : 128); mbar_init(tid_bar(a), 1);
      mbar_init(sig_full_bar(a), kFusedEpilogueWarps); mbar_init(sig_empty_bar(a), 1);
    }
    mbar_init(wfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    pdl_trigger();
    if (WRES) {   // constant data: may be fetched while the previous kernel is still running
      mbar_expect_tx(wfull_bar, (uint32_t)p.nchunks * W_CHUNK_BYTES);
      for (int c = 0; c < p.nchunks; ++c)
        bulk_load(wres_base + c * W_CHUNK_BYTES, p.w + (size_t)c * (W_CHUNK_BYTES / 2), W_CHUNK_BYTES, wfull_bar);
    }
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) sbias[i] = p.bias[i];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything below reads / writes tensors of the stream's earlier kernels

  // static schedule (FUSE = false): every CTA walks ONE contiguous range of tiles -- consecutive
  // tiles share halo rows in L2 and stay within one or two slices, so the InstanceNorm sums are
  // flushed once per slice change.  FUSE: tiles are claimed from sched[0] in slice order.
  const int tiles_per_slice = p.tiles_x * p.tiles_y;
  const int total_tiles = tiles_per_slice * p.n_slices;
  const int tiles_per_cta = (total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile_begin = FUSE ? 0 : min((int)blockIdx.x * tiles_per_cta, total_tiles);
  const int tile_end = FUSE ? total_tiles : min(tile_begin + tiles_per_cta, total_tiles);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      int tile = FUSE ? atomicAdd(p.sched, 1) : tile_begin;
      while (tile < tile_end) {
        // FUSE: the next tile is claimed one tile ahead, so the atomic's latency hides behind this tile's loads
        const int next = FUSE ? atomicAdd(p.sched, 1) : tile + 1;
        const int nl = tile / tiles_per_slice, r = tile - nl * tiles_per_slice;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        const int x0 = tx * 8 * NT, y0 = ty * 16;
        const int ng = p.n0 + nl;
        const int b = ng / p.n_div, d = ng - b * p.n_div;
        const int slice1 = p.in_global ? b : nl;
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (FUSE && c == 0) stage_tile[stage] = tile;      // published by the arrive below
          mbar_expect_tx(full_bar(stage), S * A_TERM_BYTES + (WRES ? 0u : W_CHUNK_BYTES));
          const uint32_t sa = stage_base + stage * p.stage_bytes;
          const bool second = c >= p.nchunks1;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            // box origin (pixel xo, row y0 - 1, first plane); d >= W: the shifted image is all
            // zeros -> park the box fully out of bounds
            const int xo = !second ? x0 - 1 : (d >= p.W ? -(PW + 16) : x0 - 1 - d);
            const int pl = !second ? (slice1 * S + s) * p.planes1 + 2 * c
                                   : (b * S + s) * p.planes2 + 2 * (c - p.nchunks1);
            const CUtensorMap* mp = !second ? &p.map0 : p.maps_d + d;
            if (tc_flat_rows(PW)) tma_load_4d(sa + s * A_TERM_BYTES, mp, 8 * xo, y0 - 1, pl, 0, full_bar(stage));
            else tma_load_4d(sa + s * A_TERM_BYTES, mp, 0, xo, y0 - 1, pl, full_bar(stage));
          }
          if (!WRES)
            bulk_load(sa + S * A_TERM_BYTES, p.w + (size_t)c * (W_CHUNK_BYTES / 2), W_CHUNK_BYTES,
                      full_bar(stage));
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
        tile = next;
      }
      if (FUSE) {   // end marker: a stage that carries tile id -1 and no data
        mbar_wait(empty_bar(stage), phase ^ 1);
        stage_tile[stage] = -1;
        mbar_arrive(full_bar(stage));
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the (warp-uniform) loops, one elected lane issues =====
    {
      const uint32_t fmt = (1u << 4) | (p.fp16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(128 >> 4) << 24);
      uint32_t idesc[S];
#pragma unroll
      for (int s = 0; s < S; ++s) idesc[s] = fmt | ((uint32_t)((N * (S - s)) >> 3) << 17);
      if (WRES) { mbar_wait(wfull_bar, 0); tc_fence_after(); }
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
#ifdef PDS_TC_TRACE
      const long long t_start = clock64();
      int n_regions = 0;
#endif
      if (!FUSE && p.continuous) {
        // Static schedule, continuous issue: ONE elected lane runs the whole role.  Between two
        // chunks the per-chunk form leaves the tensor pipe idle for ~500 cycles (commit, warp
        // reconvergence, barrier wait, election, ~50 uniform-datapath instructions of descriptor
        // set-up: stamps of PDS_B200_TC_TRACE, 1.8 K of the 10.8 K cycles of a tile).  Here the lane
        // never leaves the issue stream: the next chunk's operands are awaited between the bursts of
        // the two activation terms -- while the pipe still holds the first burst -- and the
        // descriptors advance by additions.
        if (elect_one()) {
          bool have = false;
          for (int tile = tile_begin; tile < tile_end; ++tile) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_base = tmem_base + acc * BUF_COLS;
            for (int c = 0; c < p.nchunks; ++c) {
              if (!have) mbar_wait(full_bar(stage), phase);      // only the very first chunk of the launch
              have = false;
              tc_fence_after();
              const uint32_t sa = stage_base + stage * p.stage_bytes;
              const uint32_t sw = WRES ? wres_base + c * W_CHUNK_BYTES : sa + S * A_TERM_BYTES;
              const uint32_t nstage = stage + 1 == (uint32_t)p.stages ? 0u : stage + 1, nphase = nstage == 0 ? phase ^ 1 : phase;
              const bool more = !(tile + 1 == tile_end && c + 1 == p.nchunks);
#ifdef PDS_TC_TRACE
              if (p.trace && blockIdx.x == 0 && n_regions < 60) p.trace[1 + 2 * n_regions] = clock64() - t_start;
#endif
              const uint64_t wdesc = umma_desc_kmajor(sw, S * N * 16, 128);
#pragma unroll
              for (int s = 0; s < S; ++s) {
                const uint64_t adesc = umma_desc_kmajor(sa + s * A_TERM_BYTES, kPH * PW * 16, PW * 16);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                  const uint64_t bd = wdesc + (uint64_t)(tap * 2 * S * N);          // 16-byte units
#pragma unroll
                  for (int i = 0; i < NT; ++i) {
                    const uint64_t ad = adesc + (uint64_t)((tap / 3) * PW + 8 * i + tap % 3);
                    tc_mma(d_base + i * ACC_COLS + s * N, ad, bd, idesc[s],
                           (s == 0 && tap == 0) ? (c != 0 ? 1u : 0u) : 1u);
                  }
                }
                if (s == 0 && more) { mbar_wait(full_bar(nstage), nphase); have = true; }
              }
              tc_commit(empty_bar(stage));       // frees the smem stage once these MMAs retire
              if (c == p.nchunks - 1) tc_commit(tfull_bar(acc));   // accumulators ready for the epilogue
#ifdef PDS_TC_TRACE
              if (p.trace && blockIdx.x == 0 && n_regions < 60) { p.trace[2 + 2 * n_regions] = clock64() - t_start; p.trace[0] = ++n_regions; }
#endif
              stage = nstage; phase = nphase;
            }
            if (NBUF == 2) { acc ^= 1; if (acc == 0) acc_phase ^= 1; } else { acc_phase ^= 1; }
          }
        }
        __syncwarp();
      } else
      for (int tile = tile_begin; FUSE || tile < tile_end; ++tile) {
        if (FUSE) {
          mbar_wait(full_bar(stage), phase);           // first chunk of the next tile, or the end marker
          const int id = stage_tile[stage];
          mbar_wait(tempty_bar(acc), acc_phase ^ 1);   // the epilogue has read this buffer's previous tile id
          if (lane == 0) { acc_tile[acc] = id; mbar_arrive(tid_bar(acc)); }
          __syncwarp();
          if (id < 0) break;
        } else {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        }
        tc_fence_after();
        const uint32_t d_base = tmem_base + acc * BUF_COLS;
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = stage_base + stage * p.stage_bytes;
          const uint32_t sw = WRES ? wres_base + c * W_CHUNK_BYTES : sa + S * A_TERM_BYTES;
#ifdef PDS_TC_TRACE
          if (p.trace && blockIdx.x == 0 && lane == 0 && n_regions < 60) p.trace[1 + 2 * n_regions] = clock64() - t_start;
#endif
          if (elect_one()) {
            const uint64_t wdesc = umma_desc_kmajor(sw, S * N * 16, 128);
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const uint64_t adesc = umma_desc_kmajor(sa + s * A_TERM_BYTES, kPH * PW * 16, PW * 16);
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const uint64_t bd = wdesc + (uint64_t)(tap * 2 * S * N);          // 16-byte units
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                  const uint64_t ad = adesc + (uint64_t)((tap / 3) * PW + 8 * i + tap % 3);
                  tc_mma(d_base + i * ACC_COLS + s * N, ad, bd, idesc[s],
                         (s == 0 && tap == 0) ? (c != 0 ? 1u : 0u) : 1u);
                }
              }
            }
            tc_commit(empty_bar(stage));       // frees the smem stage once these MMAs retire
            if (c == p.nchunks - 1) tc_commit(tfull_bar(acc));   // accumulators ready for the epilogue
          }
          __syncwarp();
#ifdef PDS_TC_TRACE
          if (p.trace && blockIdx.x == 0 && lane == 0 && n_regions < 60) { p.trace[2 + 2 * n_regions] = clock64() - t_start; p.trace[0] = ++n_regions; }
#endif
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
        if (NBUF == 2) { acc ^= 1; if (acc == 0) acc_phase ^= 1; } else { acc_phase ^= 1; }
      }
    }
  } else if (!FUSE) {
    // ===== epilogue warps (2..5), static schedule: TMEM lanes 32*(warp%4) .. +31 =====
    const int q = warp & 3;
    const int row = 32 * q + lane;
    const int px = row & 7, py = row >> 3;
    const size_t HW = (size_t)p.H * p.W;
    uint32_t acc = 0, acc_phase = 0;
    // InstanceNorm sums of the current slice: lane l holds channels l and 32 + l
    double sa0 = 0.0, sa1 = 0.0, sb0 = 0.0, sb1 = 0.0;
    int stat_slice = -1;
    auto flush_stats = [&]() {
      if (stat_slice >= 0) {
        double* dst = p.stats + ((size_t)stat_slice * N + lane) * 2;
        atomicAdd(dst, sa0); atomicAdd(dst + 1, sa1);
        atomicAdd(dst + 64, sb0); atomicAdd(dst + 65, sb1);
      }
      sa0 = sa1 = sb0 = sb1 = 0.0;
    };
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int nl = tile / tiles_per_slice, r = tile - nl * tiles_per_slice;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int y = ty * 16 + py;
      const int ng = p.n0 + nl;
      if (p.epilogue == TC_EPI_ACT && ng != stat_slice) { flush_stats(); stat_slice = ng; }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(32 * q) << 16) + acc * BUF_COLS;
#pragma unroll 1
      for (int i = 0; i < NT; ++i) {
        const int x = tx * 8 * NT + 8 * i + px;
        const bool valid = (x < p.W) && (y < p.H);
        const size_t pix = (size_t)y * p.W + x;
        if constexpr (N == 16) {   // last convolution: bias -> signatures (B, Cout, D, H, W)
          float v[16];
          tmem_ld<16>(t_base + i * ACC_COLS + (S - 1) * N, v);
#pragma unroll
          for (int s = S - 2; s >= 0; --s) {
            float u[16];
            tmem_ld<16>(t_base + i * ACC_COLS + s * N, u);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += u[j];
          }
          if (valid) {
            const int b = ng / p.n_div, dd = ng - b * p.n_div;
#pragma unroll
            for (int ch = 0; ch < 16; ++ch)
              if (ch < p.Cout)
                p.out_sig[(((size_t)b * p.Cout + ch) * p.n_div + dd) * HW + pix] =
                    fmaf(v[ch], p.inv_wscale, sbias[ch]);
          }
        } else {
#pragma unroll 1
        for (int col0 = 0; col0 < N; col0 += 32) {
          float v[32];
          tmem_ld<32>(t_base + i * ACC_COLS + (S - 1) * N + col0, v);   // smallest terms first
#pragma unroll
          for (int s = S - 2; s >= 0; --s) {
            float u[32];
            tmem_ld<32>(t_base + i * ACC_COLS + s * N + col0, u);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += u[j];
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float t = fmaf(v[j], p.inv_wscale, sbias[col0 + j]);
            if (p.epilogue == TC_EPI_ACT) t = t > 0.f ? t : 0.1f * t;
            v[j] = t;
          }
          if (p.epilogue == TC_EPI_PLAIN) {
            if (valid) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint16_t t[8][3];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  if (p.fp16) split_terms<true>(v[8 * g + e], t[e]); else split_terms<false>(v[8 * g + e], t[e]);
                }
#pragma unroll
                for (int s = 0; s < S; ++s) {
                  union { uint16_t h[8]; uint4 u; } pk;
#pragma unroll
                  for (int e = 0; e < 8; ++e) pk.h[e] = t[e][s];
                  reinterpret_cast<uint4*>(p.out_ap)[((size_t)(nl * S + s) * (N / 8) + col0 / 8 + g) * HW + pix] = pk.u;
                }
              }
            }
          } else {
            if (valid) {
              float4* o = reinterpret_cast<float4*>(p.out_f32) + ((size_t)nl * (N / 4) + col0 / 4) * HW + pix;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                o[(size_t)k * HW] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            }
            if (p.epilogue != TC_EPI_ACT) continue;
            // InstanceNorm partial sums over this warp's 32 pixels, one channel per lane
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (!valid) v[j] = 0.f;
              sq[j] = v[j] * v[j];
            }
            const float s1 = warp_transpose_reduce(v, lane);
            const float s2 = warp_transpose_reduce(sq, lane);
            if (col0 == 0) { sa0 += (double)s1; sa1 += (double)s2; }
            else { sb0 += (double)s1; sb1 += (double)s2; }
          }
        }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (NBUF == 2) { acc ^= 1; if (acc == 0) acc_phase ^= 1; } else { acc_phase ^= 1; }
    }
    if (p.epilogue == TC_EPI_ACT) flush_stats();
  } else if (warp >= kFusedEpilogueWarp0 && warp < kNormWarp0) {
    // ===== FUSE epilogue warps (4..11): two per TMEM lane quarter (lanes 32*(warp%4) .. +31), each takes
    // half of the CTA tile's MMA tiles; 16 output channels at a time (the 512-thread CTA leaves 128
    // registers per thread).  Every tile: bias, LeakyReLU, fp32 store, the tile's InstanceNorm sums as
    // fire-and-forget double atomics into this warp's PRIVATE copy of the slice's sums (no two warps of
    // a CTA share an address), then the slice id goes to the progress warp. =====
    const int q = warp & 3, half = (warp - kFusedEpilogueWarp0) >> 2;
    const int row = 32 * q + lane;
    const int px = row & 7, py = row >> 3;
    const size_t HW = (size_t)p.H * p.W;
    uint32_t acc = 0, acc_phase = 0, sig = 0, sig_phase = 0;
    double* my_stats = p.stats_rep + (size_t)(warp - kFusedEpilogueWarp0) * p.n_slices * N * 2;
    for (;;) {
      mbar_wait(tid_bar(acc), acc_phase);
      const int tile = acc_tile[acc];
      if (tile < 0) break;
      const int nl = tile / tiles_per_slice, r = tile - nl * tiles_per_slice;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int y = ty * 16 + py;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(32 * q) << 16) + acc * BUF_COLS;
      // after the transposed reduction lane l holds sum(x) of channel 16 j + l (l < 16) or sum(x^2) of
      // channel 16 j + l - 16 (l >= 16)
      float ssum[N / 16];
#pragma unroll
      for (int j = 0; j < N / 16; ++j) ssum[j] = 0.f;
#pragma unroll 1
      for (int i = half * (NT / 2); i < (half + 1) * (NT / 2); ++i) {
        const int x = tx * 8 * NT + 8 * i + px;
        const bool valid = (x < p.W) && (y < p.H);
        const size_t pix = (size_t)y * p.W + x;
#pragma unroll
        for (int jc = 0; jc < N / 16; ++jc) {
          const int col0 = 16 * jc;
          float v[32];                                                   // [0, 16): values, [16, 32): their squares
          {
            float lo16[16];
            tmem_ld<16>(t_base + i * ACC_COLS + (S - 1) * N + col0, lo16);   // smallest terms first
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = lo16[j];
          }
#pragma unroll
          for (int s = S - 2; s >= 0; --s) {
            float u[16];
            tmem_ld<16>(t_base + i * ACC_COLS + s * N + col0, u);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += u[j];
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float t = fmaf(v[j], p.inv_wscale, sbias[col0 + j]);
            v[j] = t > 0.f ? t : 0.1f * t;
          }
          if (valid) {
            float4* o = reinterpret_cast<float4*>(p.out_f32) + ((size_t)nl * (N / 4) + col0 / 4) * HW + pix;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              o[(size_t)k * HW] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (!valid) v[j] = 0.f;
            v[16 + j] = v[j] * v[j];
          }
          const float part = warp_transpose_reduce(v, lane);     // 32 pixels, as the static kernel sums them
          ssum[jc] = (NT / 2 == 1) ? part : ssum[jc] + part;
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (NBUF == 2) { acc ^= 1; if (acc == 0) acc_phase ^= 1; } else { acc_phase ^= 1; }
#ifndef PDS_FUSE_DEBUG_NO_STATS
      {
        double* dst = my_stats + ((size_t)nl * N + (lane & 15)) * 2 + (lane >> 4);
#pragma unroll
        for (int j = 0; j < N / 16; ++j) atomicAdd(dst + 32 * j, (double)ssum[j]);
        // the progress warp fences at gpu scope and bumps the slice's counter: no epilogue warp ever
        // waits for its own stores to be acknowledged (nothing to publish when the normalisation runs
        // as separate passes after this launch)
        if (p.norm_mode != TC_NORM_NONE) {
          if (lane == 0) mbar_wait(sig_empty_bar(sig), sig_phase ^ 1);
          if (warp == kFusedEpilogueWarp0 && lane == 0) sig_slice[sig] = nl;
          __syncwarp();
          if (lane == 0) mbar_arrive(sig_full_bar(sig));     // release: orders this warp's stores and atomics before it
          sig ^= 1; if (sig == 0) sig_phase ^= 1;
        }
      }
#endif
    }
#ifndef PDS_FUSE_DEBUG_NO_STATS
    // end marker for the progress warp
    if (p.norm_mode != TC_NORM_NONE) {
      if (lane == 0) mbar_wait(sig_empty_bar(sig), sig_phase ^ 1);
      if (warp == kFusedEpilogueWarp0 && lane == 0) sig_slice[sig] = -1;
      __syncwarp();
      if (lane == 0) mbar_arrive(sig_full_bar(sig));
    }
#endif
  } else if (warp == kProgressWarp) {
    // ===== progress warp: publishes finished tiles.  The epilogue warps arrive (release) on sig_full
    // after a tile's stores and sum atomics; one thread here acquires, fences at gpu scope (cumulative
    // over what it has observed: the grid-sync idiom) and moves the slice's counter. =====
#ifndef PDS_FUSE_DEBUG_NO_STATS
    if (lane == 0 && p.norm_mode != TC_NORM_NONE) {
      uint32_t sig = 0, sig_phase = 0;
      for (;;) {
        mbar_wait(sig_full_bar(sig), sig_phase);
        const int nl = sig_slice[sig];
        mbar_arrive(sig_empty_bar(sig));
        if (nl < 0) break;
        __threadfence();
        red_release_gpu_add(p.sched + 1 + nl, 1);
        sig ^= 1; if (sig == 0) sig_phase ^= 1;
      }
    }
#endif
  } else if (warp >= kNormWarp0 || warp == 3) {
    // ===== normalisation warps: slice classes in order, a fixed share of every slice's units.  Nothing
    // here waits on a CTA that may not be resident: a slice's counter depends only on tiles, and tiles
    // are claimed by whoever runs. =====
#ifndef PDS_FUSE_DEBUG_NO_NORM_WARPS
    const int g = (int)blockIdx.x * kNormWarps + (warp == 3 ? 0 : warp - kNormWarp0 + 1), G = (int)gridDim.x * kNormWarps;
    const int cls = g % kNormSliceClasses, idx = g / kNormSliceClasses;
    const int cnt = (G - cls + kNormSliceClasses - 1) / kNormSliceClasses;       // warps of this class
    const int blocks = (int)(((size_t)p.H * p.W + kNormPixels - 1) / kNormPixels);
    const long long units = (long long)blocks * (N / 8);
    const int u0 = (int)(units * idx / cnt), u1 = (int)(units * (idx + 1) / cnt);
    for (int nl = cls; nl < (p.norm_mode != TC_NORM_NONE ? p.n_slices : 0); nl += kNormSliceClasses) {
      wait_counter(p.sched + 1 + nl, tiles_per_slice);
      norm_range<S>(p, nl, u0, u1, blocks, lane);
    }
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- auxiliary kernels ------------------------------------------------------------------

// (Cout, src_cin, 3, 3) fp32 -> [chunk][tap][2][S][N][8] terms of w * wscale (zero rows for
// cout >= Cout).  The layer's input channels are the source channels [ci_off, ci_off + Cin).
// qmode: only the kx = 2 column of the kernel survives, moved to the centre column (the
// "right-neighbour taps applied in place" operator of the factorised first convolution).
template <bool FP16>
__global__ void tc_prepare_weights_kernel(const float* __restrict__ w, uint16_t* __restrict__ out,
                                          int Cout, int Cin, int N, int S, float wscale, int src_cin,
                                          int ci_off, int qmode) {
  const size_t total = (size_t)(Cin / 16) * 9 * 2 * N * 8;   // per term
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int e = i % 8;
  const int co = (i / 8) % N;
  const int j = (i / (8 * (size_t)N)) % 2;
  const int tap = (i / (16 * (size_t)N)) % 9;
  const int c = i / (144 * (size_t)N);
  const int ci = c * 16 + j * 8 + e;
  int tap_src = tap;
  bool live = co < Cout;
  if (qmode) {
    live = live && (tap % 3 == 1);
    tap_src = tap + 1;
  }
  const float x = live ? w[((size_t)co * src_cin + ci_off + ci) * 9 + tap_src] * wscale : 0.f;
  uint16_t t[3];
  split_terms<FP16>(x, t);
  for (int s = 0; s < S; ++s)
    out[((((size_t)c * 9 + tap) * 2 + j) * S + s) * N * 8 + (size_t)co * 8 + e] = t[s];
}

__global__ void tc_pad_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int Cout, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = (b && i < Cout) ? b[i] : 0.f;
}

// First convolution of the matching operation, factorised (it is linear in the concatenation):
//   conv0(cat[L, shift_d R]) = A + shift_d(Bf) + edge terms,   A = conv(L; W_left) + bias,
//   Bf = conv(R; W_right),  Q = the kx = 2 taps of W_right applied in place.
// shift_d(Bf)[x] = Bf[x - d]; the zero-filled columns x < d see only R's column 0 through the
// right-neighbour taps (x = d - 1 -> Q[0]); at the image's last column those taps read the zero
// padding in the reference but R[W - d] in Bf -> subtract Q[W - d] (d >= 1).
// A / Bf / Q: fp32 planes [b][C/4][H][W][4]; out: split AP planes [b*D + d][S][C/8][H][W][8].
template <bool FP16, int S>
__global__ void __launch_bounds__(256)
tc_compose_first_kernel(const float* __restrict__ A, const float* __restrict__ Bf,
                        const float* __restrict__ Q, uint16_t* __restrict__ out, int C, int H, int W,
                        int D) {
  const int c8 = blockIdx.y, n = blockIdx.z;
  const int b = n / D, d = n - b * D;
  const size_t HW = (size_t)H * W;
  const size_t base = ((size_t)b * (C / 4) + 2 * c8) * HW;
  if (threadIdx.x == 0) pdl_trigger();
  pdl_wait();
  const float4* a4 = reinterpret_cast<const float4*>(A) + base;
  const float4* b4 = reinterpret_cast<const float4*>(Bf) + base;
  const float4* q4 = reinterpret_cast<const float4*>(Q) + base;
  float4* o4 = reinterpret_cast<float4*>(out);
  for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(pix % W);
    const size_t row = pix - x;
    float4 lo = __ldg(a4 + pix), hi = __ldg(a4 + HW + pix);
    float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    auto add = [&](const float4* src, size_t at, float sign) {
      const float4 l = __ldg(src + at), h = __ldg(src + HW + at);
      v[0] = fmaf(sign, l.x, v[0]); v[1] = fmaf(sign, l.y, v[1]); v[2] = fmaf(sign, l.z, v[2]); v[3] = fmaf(sign, l.w, v[3]);
      v[4] = fmaf(sign, h.x, v[4]); v[5] = fmaf(sign, h.y, v[5]); v[6] = fmaf(sign, h.z, v[6]); v[7] = fmaf(sign, h.w, v[7]);
    };
    if (x >= d) add(b4, pix - d, 1.f);
    else if (x == d - 1) add(q4, row, 1.f);
    if (x == W - 1 && d >= 1 && d <= W) add(q4, row + (W - d), -1.f);
    uint16_t t[8][3];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_terms<FP16>(v[e], t[e]);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      float4 pk;
      pk.x = __uint_as_float(t[0][s] | ((uint32_t)t[1][s] << 16)); pk.y = __uint_as_float(t[2][s] | ((uint32_t)t[3][s] << 16));
      pk.z = __uint_as_float(t[4][s] | ((uint32_t)t[5][s] << 16)); pk.w = __uint_as_float(t[6][s] | ((uint32_t)t[7][s] << 16));
      stg_stream(o4 + ((size_t)(n * S + s) * (C / 8) + c8) * HW + pix, pk);
    }
  }
}

// (B, C, H, W) fp32 -> AP [B][S][C/8][H][W][8]
template <bool FP16>
__global__ void tc_pack_nchw_kernel(const float* __restrict__ in, uint16_t* __restrict__ ap,
                                    int C, size_t HW, int S, size_t total, int x4_width) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t pix = i % HW;
  const int c8 = (i / HW) % (C / 8);
  const size_t b = i / (HW * (C / 8));
  uint16_t t[8][3];
#pragma unroll
  for (int e = 0; e < 8; ++e) split_terms<FP16>(in[(b * C + c8 * 8 + e) * HW + pix], t[e]);
  for (int s = 0; s < S; ++s) {
    union { uint16_t h[8]; uint4 u; } pk;
#pragma unroll
    for (int e = 0; e < 8; ++e) pk.h[e] = t[e][s];
    if (x4_width) {    // four x-phase sub-volumes (conv_tcg.cuh, TCG_CONV3_S1X4): [b][S][x & 3][C/8][rows][W/4][8]
      const size_t row = pix / (size_t)x4_width, x = pix - row * x4_width;
      reinterpret_cast<uint4*>(ap)[(((b * S + s) * 4 + (x & 3)) * (C / 8) + c8) * (HW / 4) + row * (x4_width / 4) + (x >> 2)] = pk.u;
    } else {
      reinterpret_cast<uint4*>(ap)[((b * S + s) * (C / 8) + c8) * HW + pix] = pk.u;
    }
  }
}

// InstanceNorm from accumulated sums on fp32 planes; one thread = 8 channels of a pixel, two
// pixels per iteration with every load issued before the first use (the pass is a pure HBM
// stream: 4 B read (+ 2S B residual) and 2S B written per element).
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
  uint4 r;
  // not .nc: the residual planes are updated in place by this very kernel
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void stg_stream_u4(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w) : "memory");
}

template <bool FP16, int S, bool RES>
__global__ void __launch_bounds__(256, 3)
tc_norm_split_kernel(const float* __restrict__ y, const double* __restrict__ stats, int n_rep, size_t rep_stride,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     const uint16_t* res_ap, uint16_t* out_ap, int C, size_t HW) {
  constexpr int U = 2;   // pixels per thread per iteration
  const int c8 = blockIdx.y, n = blockIdx.z;
  __shared__ float sc[8], sh[8];
  if (threadIdx.x == 0) pdl_trigger();
  pdl_wait();
  if (threadIdx.x < 8) {
    const int c = c8 * 8 + threadIdx.x;
    double s = 0.0, q = 0.0;     // n_rep private copies of the sums (one per epilogue warp of the dynamic kernel)
    for (int r = 0; r < n_rep; ++r) {
      s += stats[r * rep_stride + ((size_t)n * C + c) * 2];
      q += stats[r * rep_stride + ((size_t)n * C + c) * 2 + 1];
    }
    const double mean = s / (double)HW;
    double var = q / (double)HW - mean * mean;
    if (var < 0.0) var = 0.0;
    const float scale = (float)(1.0 / sqrt(var + 1e-5)) * gamma[c];
    sc[threadIdx.x] = scale;
    sh[threadIdx.x] = beta[c] - (float)mean * scale;
  }
  __syncthreads();
  float a[8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { a[e] = sc[e]; b[e] = sh[e]; }
  const float4* y4 = reinterpret_cast<const float4*>(y) + ((size_t)n * (C / 4) + 2 * c8) * HW;
  const uint4* r4 = reinterpret_cast<const uint4*>(res_ap);
  uint4* o4 = reinterpret_cast<uint4*>(out_ap);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t p0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p0 < HW; p0 += U * stride) {
    float4 lo[U], hi[U];
    uint4 rr[U][S];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t pix = p0 + u * stride;
      if (pix < HW) {
        lo[u] = ldg_stream(y4 + pix);
        hi[u] = ldg_stream(y4 + HW + pix);
        if (RES) {
#pragma unroll
          for (int s = 0; s < S; ++s) rr[u][s] = ldg_stream_u4(r4 + ((size_t)(n * S + s) * (C / 8) + c8) * HW + pix);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t pix = p0 + u * stride;
      if (pix >= HW) continue;
      float v[8] = {lo[u].x, lo[u].y, lo[u].z, lo[u].w, hi[u].x, hi[u].y, hi[u].z, hi[u].w};
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], a[e], b[e]);
      if (RES) {
        float res[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = S - 1; s >= 0; --s) {   // smallest term first
          const uint32_t w[4] = {rr[u][s].x, rr[u][s].y, rr[u][s].z, rr[u][s].w};
#pragma unroll
          for (int e = 0; e < 8; ++e) res[e] += term_value<FP16>((uint16_t)(w[e >> 1] >> (16 * (e & 1))));
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += res[e];
      }
      uint16_t t[8][3];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_terms<FP16>(v[e], t[e]);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        uint4 pk;
        pk.x = t[0][s] | ((uint32_t)t[1][s] << 16); pk.y = t[2][s] | ((uint32_t)t[3][s] << 16);
        pk.z = t[4][s] | ((uint32_t)t[5][s] << 16); pk.w = t[6][s] | ((uint32_t)t[7][s] << 16);
        stg_stream_u4(o4 + ((size_t)(n * S + s) * (C / 8) + c8) * HW + pix, pk);
      }
    }
  }
}

// ---- host side -----------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult res;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &res) == cudaSuccess &&
        res == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// AP tensor viewed as (8 ch, W_eff, H, planes) with row pitch W.  (The element
// type only sets the element size: fp16 and bf16 planes use the same map.)
int encode_ap_map(CUtensorMap* map, const void* base, int W_eff, int W, int H, size_t planes,
                  int PW) {
  // The pixel and its 8 channels form ONE tensor-map dimension when the box row fits (8 * PW <=
  // 256 elements): a box row is then a single 16 * PW byte request instead of PW requests of 16
  // bytes (same bytes in shared memory; out-of-bounds pixels are still zero-filled, element-wise).
  const bool flat = tc_flat_rows(PW);
  const cuuint64_t dims4[4] = {8, (cuuint64_t)W_eff, (cuuint64_t)H, (cuuint64_t)planes};
  const cuuint64_t strides4[3] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
  const cuuint32_t box4[4] = {8, (cuuint32_t)PW, (cuuint32_t)kPH, 2};
  const cuuint64_t dimsf[4] = {(cuuint64_t)W_eff * 8, (cuuint64_t)H, (cuuint64_t)planes, 1};
  const cuuint64_t stridesf[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)planes * H * W * 16};
  const cuuint32_t boxf[4] = {(cuuint32_t)PW * 8, (cuuint32_t)kPH, 2, 1};
  const cuuint64_t* dims = flat ? dimsf : dims4;
  const cuuint64_t* strides = flat ? stridesf : strides4;
  const cuuint32_t* box = flat ? boxf : box4;
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (W_eff=%d W=%d H=%d planes=%zu)", (int)r,
              W_eff, W, H, planes);
    return PDS_ERR_CUDA;
  }
  return PDS_OK;
}

template <int S, int NT, int N, bool WRES, bool FUSE>
int launch_tc(TcKernelParams& p, cudaStream_t st) {
  constexpr int PW = 8 * NT + 2;
  constexpr uint32_t A_TERM_BYTES = 2 * kPH * PW * 16;
  constexpr uint32_t W_CHUNK_BYTES = 9 * 2 * S * N * 16;
  p.wres_bytes = WRES ? (uint32_t)p.nchunks * W_CHUNK_BYTES : 0;
  p.stage_bytes = (uint32_t)align_up((size_t)S * A_TERM_BYTES + (WRES ? 0 : W_CHUNK_BYTES), 128);
  const size_t budget = 227 * 1024 - tail_bytes(FUSE) - p.wres_bytes;
  int stages = (int)(budget / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  static const int stage_cap = getenv("PDS_B200_TC_STAGES") ? atoi(getenv("PDS_B200_TC_STAGES")) : 0;
  if (stage_cap >= 2 && stages > stage_cap) stages = stage_cap;   // leaves shared memory for co-resident CTAs
  if (stages < 2) {
    set_error("conv3x3_tc: stage of %u bytes (+ %u resident) does not fit twice in shared memory",
              p.stage_bytes, p.wres_bytes);
    return PDS_ERR_UNSUPPORTED;
  }
  p.stages = stages;
  const size_t smem = p.wres_bytes + (size_t)stages * p.stage_bytes + tail_bytes(FUSE);
  PDS_CUDA(allow_dynamic_smem(conv3x3_tc_kernel<S, NT, N, WRES, FUSE>, 227 * 1024));
  const int total = p.tiles_x * p.tiles_y * p.n_slices;
  const int grid = total < num_sms() ? total : num_sms();
  static const std::string base = "conv3x3_tc<S=" + std::to_string(S) + ",NT=" + std::to_string(NT) +
                                  ",N=" + std::to_string(N) + (WRES ? ",Wres" : ",Wstream") + (FUSE ? ",dyn>" : ">");
  // PDS_B200_PROFILE_DETAIL=1: one profiler class per epilogue variant and slice count
  static const bool detail = getenv("PDS_B200_PROFILE_DETAIL") && atoi(getenv("PDS_B200_PROFILE_DETAIL"));
  static std::string names[8][2][2];
  const bool per_sample = p.n_div == 1 && !p.in_global;     // descriptor-level launch (not one slice per disparity)
  const bool with_norm = FUSE && p.norm_mode != TC_NORM_NONE;
  std::string& nm = names[p.epilogue & 7][per_sample ? 0 : 1][with_norm ? 1 : 0];
  if (nm.empty())
    nm = base + (detail ? "[epi " + std::to_string(p.epilogue) + (per_sample ? ", per sample]" : ", all slices]")
                        : (per_sample ? "[per sample]" : "")) + (with_norm ? "[+IN in the launch]" : "");
  PDS_KERNEL(nm.c_str(), st);
  {
    // reference FLOPs of the layer (real Cout, all Cin); bytes: AP terms in (both inputs), output as written
    // (FUSE: the operand planes the normalisation warps write -- and the residual they read -- instead
    // of the fp32 activation, which is consumed from L2)
    const double px = (double)p.H * p.W * p.n_slices;
    double out_b = p.epilogue == TC_EPI_SIG ? 4.0 * p.Cout : (p.epilogue == TC_EPI_PLAIN ? 2.0 * S * N : 4.0 * N);
    if (with_norm) out_b = 2.0 * S * N * (p.norm_mode == TC_NORM_RESIDUAL ? 2.0 : 1.0);
    PDS_KERNEL_WORK(2.0 * 9 * 16 * p.nchunks * p.Cout * px, px * out_b + (p.in_global ? 0.0 : px * 2.0 * S * 16 * p.nchunks));
  }
  PDS_CUDA(launch_pdl(conv3x3_tc_kernel<S, NT, N, WRES, FUSE>, dim3(grid), dim3(FUSE ? kFusedThreads : kThreads), smem,
                      st, p));
  return PDS_OK;
}

// MMA tiles per CTA tile for S terms (double-buffered accumulators: 2*NT*S*64 <= 512 columns)
int nt_for(int S) { return S == 1 ? 4 : (S == 2 ? 2 : 1); }

}  // namespace

bool tc_available() { return encode_fn() != nullptr; }

size_t tc_conv_max_maps(int n_div) { return (size_t)(n_div > 0 ? n_div : 0) + 1; }

int tc_prepare_weights(const TcLayer& l, const float* w_oihw, const float* bias, cudaStream_t st,
                       int src_cin, int ci_off, int qmode) {
  if (src_cin <= 0) src_cin = l.Cin;
  const size_t per_term = l.w_elems() / l.S;
  {
    PDS_KERNEL("tc_prepare_weights", st);
    const unsigned g = (unsigned)((per_term + 255) / 256);
    if (l.fp16)
      tc_prepare_weights_kernel<true><<<g, 256, 0, st>>>(w_oihw, l.w, l.Cout, l.Cin, l.N, l.S, l.wscale, src_cin, ci_off, qmode);
    else
      tc_prepare_weights_kernel<false><<<g, 256, 0, st>>>(w_oihw, l.w, l.Cout, l.Cin, l.N, l.S, l.wscale, src_cin, ci_off, qmode);
    PDS_LAUNCH_CHECK("tc_prepare_weights_kernel");
  }
  PDS_KERNEL("tc_pad_bias", st);
  tc_pad_bias_kernel<<<1, 256, 0, st>>>(bias, l.bias, l.Cout, l.N);
  PDS_LAUNCH_CHECK("tc_pad_bias_kernel");
  return PDS_OK;
}

int tc_compose_first(const float* A, const float* Bf, const float* Q, uint16_t* out_ap, int B, int C, int H,
                     int W, int D, int S, int fp16, cudaStream_t st) {
  const size_t HW = (size_t)H * W;
  if (B == 0 || HW == 0 || D == 0) return PDS_OK;
  unsigned gx = (unsigned)((HW + 255) / 256);
  if (gx > 128) gx = 128;
  dim3 grid(gx, (unsigned)(C / 8), (unsigned)(B * D));
  PDS_KERNEL("tc_compose_first", st);
  PDS_KERNEL_WORK(0, (double)B * C * HW * (12.0 + 2.0 * S * D));
#define PDS_COMPOSE_CASE(FF, SS) \
  if ((fp16 != 0) == FF && S == SS) PDS_CUDA(launch_pdl(tc_compose_first_kernel<FF, SS>, grid, dim3(256), 0, st, A, Bf, Q, out_ap, C, H, W, D));
  PDS_COMPOSE_CASE(true, 1) PDS_COMPOSE_CASE(true, 2) PDS_COMPOSE_CASE(true, 3)
  PDS_COMPOSE_CASE(false, 1) PDS_COMPOSE_CASE(false, 2) PDS_COMPOSE_CASE(false, 3)
#undef PDS_COMPOSE_CASE
  PDS_LAUNCH_CHECK("tc_compose_first_kernel");
  return PDS_OK;
}

int tc_pack_nchw(const float* in, uint16_t* ap, int B, int C, int H, int W, int S, int fp16, cudaStream_t st,
                 int x4_width) {
  const size_t HW = (size_t)H * W, total = (size_t)B * (C / 8) * HW;
  if (total == 0) return PDS_OK;
  PDS_KERNEL("tc_pack_nchw", st);
  PDS_KERNEL_WORK(0, (double)B * C * HW * (4 + 2 * S));
  const unsigned g = (unsigned)((total + 255) / 256);
  if (fp16) tc_pack_nchw_kernel<true><<<g, 256, 0, st>>>(in, ap, C, HW, S, total, x4_width);
  else tc_pack_nchw_kernel<false><<<g, 256, 0, st>>>(in, ap, C, HW, S, total, x4_width);
  PDS_LAUNCH_CHECK("tc_pack_nchw_kernel");
  return PDS_OK;
}

int tc_norm_split(const float* y, const double* stats, const float* gamma, const float* beta,
                  const uint16_t* res_ap, uint16_t* out_ap, int n_slices, int C, int H, int W, int S,
                  int fp16, cudaStream_t st, int n_rep, size_t rep_stride) {
  const size_t HW = (size_t)H * W;
  if (n_slices == 0 || HW == 0) return PDS_OK;
  // two pixels per thread per iteration; several iterations per CTA amortise its prologue (the
  // double-precision statistics and a barrier) once there are enough CTAs to fill the GPU
  static const int iters_env = getenv("PDS_B200_NORM_ITERS") ? atoi(getenv("PDS_B200_NORM_ITERS")) : 4;
  unsigned gx = (unsigned)((HW + 511) / 512);
  const size_t ctas = (size_t)gx * (C / 8) * n_slices;
  const unsigned iters = ctas >= (size_t)16 * num_sms() * iters_env ? (unsigned)iters_env : 1u;
  gx = (gx + iters - 1) / iters;
  if (gx > 128) gx = 128;
  dim3 grid(gx, (unsigned)(C / 8), (unsigned)n_slices);
  PDS_KERNEL(res_ap ? "tc_norm_residual_split" : "tc_norm_split", st);
  PDS_KERNEL_WORK(0, (double)n_slices * C * HW * (4 + 2 * S + (res_ap ? 2 * S : 0)));
#define PDS_NORM_CASE(FF, SS)                                                                              \
  if ((fp16 != 0) == FF && S == SS) {                                                                      \
    if (res_ap) PDS_CUDA(launch_pdl(tc_norm_split_kernel<FF, SS, true>, grid, dim3(256), 0, st, y, stats, n_rep, rep_stride, gamma, beta, res_ap, out_ap, C, HW)); \
    else PDS_CUDA(launch_pdl(tc_norm_split_kernel<FF, SS, false>, grid, dim3(256), 0, st, y, stats, n_rep, rep_stride, gamma, beta, res_ap, out_ap, C, HW)); \
  }
  PDS_NORM_CASE(true, 1) PDS_NORM_CASE(true, 2) PDS_NORM_CASE(true, 3)
  PDS_NORM_CASE(false, 1) PDS_NORM_CASE(false, 2) PDS_NORM_CASE(false, 3)
#undef PDS_NORM_CASE
  PDS_LAUNCH_CHECK("tc_norm_split_kernel");
  return PDS_OK;
}

// PDS_B200_FUSE_NORM=1 (read when a handle is created): InstanceNorm passes inside the convolution
// launches (built, bit-identical, slower than the separate passes today: opt-in)
bool tc_fused_norm_enabled() {
  const char* e = getenv("PDS_B200_FUSE_NORM");
  return e && atoi(e) == 1;
}
// PDS_B200_DYNAMIC_CONV=1: the dynamically scheduled kernel (tiles claimed in slice order, eight epilogue
// warps) for the 64 -> 64 layers.  Measured at 960x540 D=192: 281 us per launch against 283-287 us for the
// statically scheduled kernel -- no gain without the fusion it was built for, so the static kernel stays
// the default; PDS_B200_FUSE_NORM=1 implies the dynamic kernel.
bool tc_dynamic_conv_enabled() {
  const char* e = getenv("PDS_B200_DYNAMIC_CONV");
  return (e && atoi(e) == 1) || tc_fused_norm_enabled();
}

// scratch of one fused launch: kStatReplicas private copies of the sums, then the scheduler words
static size_t sched_stats_bytes(int n_slices) { return (size_t)kStatReplicas * (n_slices > 0 ? n_slices : 0) * 64 * 2 * sizeof(double); }
size_t tc_sched_bytes(int n_slices) {
  return align_up(sched_stats_bytes(n_slices) + (1 + (size_t)(n_slices > 0 ? n_slices : 0)) * sizeof(int), 256);
}

long long* g_tc_trace = nullptr;    // PDS_B200_TC_TRACE: stamps of the latest all-slices 64 -> 64 launch (pds_tc_trace_dump)

int tc_conv3x3(const TcConvArgs& a, cudaStream_t st) {
  const TcLayer& l = *a.layer;
  if (!encode_fn()) {
    set_error("conv3x3_tc: cuTensorMapEncodeTiled is not available from this driver");
    return PDS_ERR_UNSUPPORTED;
  }
  if (a.n_slices == 0) return PDS_OK;
  TcKernelParams p = {};
  if (getenv("PDS_B200_TC_TRACE") && a.n_slices > 8) {
    if (!g_tc_trace) { cudaMalloc(&g_tc_trace, 128 * sizeof(long long)); }
    cudaMemsetAsync(g_tc_trace, 0, 128 * sizeof(long long), st);
    p.trace = g_tc_trace;
  }
  {
    static const bool cont = !(getenv("PDS_B200_TC_CONTINUOUS") && atoi(getenv("PDS_B200_TC_CONTINUOUS")) == 0);
    p.continuous = cont ? 1 : 0;
  }
  p.maps_d = a.in2 ? a.maps_dev : nullptr;
  p.w = l.w; p.bias = l.bias;
  p.out_f32 = a.out_f32; p.out_ap = a.out_ap; p.out_sig = a.out_sig; p.stats = a.stats;
  p.n_slices = a.n_slices; p.n0 = a.n0; p.n_div = a.n_div > 0 ? a.n_div : 1; p.H = a.H; p.W = a.W;
  p.nchunks = l.Cin / 16;
  p.nchunks1 = a.in_C / 16;
  p.planes1 = a.in_C / 8;
  p.planes2 = a.in2 ? a.in2_C / 8 : 0;
  p.in_global = a.in2 ? 1 : 0;
  p.Cout = l.Cout; p.epilogue = a.epilogue; p.fp16 = l.fp16;
  p.inv_wscale = 1.0f / l.wscale;
  if ((a.in2 ? a.in_C + a.in2_C : a.in_C) != l.Cin || l.Cin % 16 || (l.N != 16 && l.N != 64) ||
      l.S < 1 || l.S > 3 || (a.epilogue == TC_EPI_SIG) != (l.N == 16)) {
    set_error("conv3x3_tc: unsupported configuration (Cin=%d, N=%d, S=%d)", l.Cin, l.N, l.S);
    return PDS_ERR_UNSUPPORTED;
  }
  if (a.norm_mode != TC_NORM_NONE && (a.epilogue != TC_EPI_ACT || !a.norm_out || !a.stats || !a.out_f32 ||
                                      (a.norm_mode == TC_NORM_RESIDUAL && !a.res_ap) ||
                                      (a.norm_mode == TC_NORM_RESIDUAL_FIRST && !(a.fA && a.fB && a.fQ && a.n0 == 0)))) {
    set_error("conv3x3_tc: incomplete arguments for the trailing normalisation");
    return PDS_ERR_INVALID_ARGUMENT;
  }
  const int NT = nt_for(l.S);
  p.tiles_x = (a.W + 8 * NT - 1) / (8 * NT);
  p.tiles_y = (a.H + 15) / 16;
  const int PW = 8 * NT + 2;
  int rc = encode_ap_map(&p.map0, a.in, a.W, a.W, a.H, (size_t)a.in_slices * l.S * (a.in_C / 8), PW);
  if (rc != PDS_OK) return rc;
  // weights resident when the whole layer fits beside >= 3 activation stages
  const size_t w_bytes = l.w_elems() * 2;
  const size_t a_stage = (size_t)l.S * 2 * kPH * PW * 16;
  const bool wres = w_bytes + 3 * a_stage + tail_bytes(false) <= 227 * 1024;
  // The dynamic kernel (tiles claimed in slice order, eight epilogue warps): 64 -> 64 layers with two
  // fp16 terms and resident weights over enough slices.  a.fuse: the InstanceNorm runs INSIDE the launch
  // (trailing normalisation warps); otherwise the launch only leaves the sums (one private copy per
  // epilogue warp) and the normalisation follows as a separate pass -- the default: with at most five
  // warps x 128 registers beside the convolution the in-kernel passes cannot keep enough loads in flight
  // and the fused launch is slower than the two kernels.
  const bool dynamic = a.sched && a.epilogue == TC_EPI_ACT && l.N == 64 && l.Cout == 64 && l.S == 2 && l.fp16 &&
                       !a.in2 && w_bytes + 3 * a_stage + tail_bytes(true) <= 227 * 1024 && a.n_slices >= 4;
  if (dynamic) {
    const bool fuse = a.fuse && a.norm_mode != TC_NORM_NONE;
    p.stats_rep = (double*)a.sched; p.sched = (int*)((char*)a.sched + sched_stats_bytes(a.n_slices));
    p.gamma = l.gamma; p.beta = l.beta; p.norm_mode = fuse ? a.norm_mode : TC_NORM_NONE;
    p.res_ap = a.res_ap; p.norm_out = a.norm_out; p.fA = a.fA; p.fB = a.fB; p.fQ = a.fQ;
    rc = launch_tc<2, 2, 64, true, true>(p, st);
    if (rc != PDS_OK || fuse) return rc;
    const size_t rep = (size_t)a.n_slices * 64 * 2;
    if (a.norm_mode == TC_NORM_PLAIN)
      return tc_norm_split(a.out_f32, p.stats_rep, l.gamma, l.beta, nullptr, a.norm_out, a.n_slices, l.N, a.H, a.W, l.S,
                           l.fp16, st, kStatReplicas, rep);
    if (a.norm_mode == TC_NORM_RESIDUAL)
      return tc_norm_split(a.out_f32, p.stats_rep, l.gamma, l.beta, a.res_ap, a.norm_out, a.n_slices, l.N, a.H, a.W, l.S,
                           l.fp16, st, kStatReplicas, rep);
    if (a.norm_mode == TC_NORM_RESIDUAL_FIRST)
      return tc_norm_residual_first(a.out_f32, p.stats_rep, l.gamma, l.beta, a.fA, a.fB, a.fQ, a.norm_out,
                                    a.n_slices / p.n_div, l.N, a.H, a.W, p.n_div, l.S, l.fp16, st, kStatReplicas, rep);
    return PDS_OK;
  }
  rc = PDS_ERR_UNSUPPORTED;
#define PDS_TC_CASE(SS, NN)                                                       \
  if (l.S == SS && l.N == NN)                                                     \
    rc = wres ? launch_tc<SS, (SS == 1 ? 4 : (SS == 2 ? 2 : 1)), NN, true, false>(p, st) \
              : launch_tc<SS, (SS == 1 ? 4 : (SS == 2 ? 2 : 1)), NN, false, false>(p, st);
  PDS_TC_CASE(1, 64) PDS_TC_CASE(1, 16) PDS_TC_CASE(2, 64) PDS_TC_CASE(2, 16)
  PDS_TC_CASE(3, 64) PDS_TC_CASE(3, 16)
#undef PDS_TC_CASE
  if (rc != PDS_OK) {
    if (rc == PDS_ERR_UNSUPPORTED) set_error("conv3x3_tc: no kernel for S=%d N=%d", l.S, l.N);
    return rc;
  }
  // the same normalisation as separate passes (three-term precisions, few slices, PDS_B200_FUSE_NORM=0)
  const double* stats = a.stats + (size_t)a.n0 * l.N * 2;
  if (a.norm_mode == TC_NORM_PLAIN)
    return tc_norm_split(a.out_f32, stats, l.gamma, l.beta, nullptr, a.norm_out, a.n_slices, l.N, a.H, a.W, l.S, l.fp16, st);
  if (a.norm_mode == TC_NORM_RESIDUAL)
    return tc_norm_split(a.out_f32, stats, l.gamma, l.beta, a.res_ap, a.norm_out, a.n_slices, l.N, a.H, a.W, l.S, l.fp16, st);
  if (a.norm_mode == TC_NORM_RESIDUAL_FIRST)
    return tc_norm_residual_first(a.out_f32, a.stats, l.gamma, l.beta, a.fA, a.fB, a.fQ, a.norm_out,
                                  a.n_slices / p.n_div, l.N, a.H, a.W, p.n_div, l.S, l.fp16, st);
  return PDS_OK;
}

// Right-descriptor maps of the first convolution: map d views the planes with width
// W - d so that columns at or beyond the shifted image's right edge read as zero.
int tc_encode_shift_maps(CUtensorMap* maps_host, const uint16_t* in2, int in_slices, int S, int C,
                         int H, int W, int D) {
  const int PW = 8 * nt_for(S) + 2;
  for (int d = 0; d < D; ++d) {
    const int weff = W - d > 0 ? W - d : 1;
    int rc = encode_ap_map(&maps_host[d], in2, weff, W, H, (size_t)in_slices * S * (C / 8), PW);
    if (rc != PDS_OK) return rc;
  }
  return PDS_OK;
}

}  // namespace pds

// debug: prints the issue-region stamps (start, length, gap) of CTA 0 of the latest traced launch
extern "C" void pds_tc_trace_dump() {
  if (!pds::g_tc_trace) { printf("no trace (PDS_B200_TC_TRACE unset)\n"); return; }
  long long h[128];
  cudaDeviceSynchronize();
  cudaMemcpy(h, pds::g_tc_trace, sizeof(h), cudaMemcpyDeviceToHost);
  const int n = (int)h[0];
  printf("conv3x3_tc issue regions of CTA 0 (start, length, gap):");
  for (int r = 0; r < n; ++r) printf(" [%lld %lld %lld]", h[1 + 2 * r], h[2 + 2 * r] - h[1 + 2 * r], r + 1 < n ? h[3 + 2 * r] - h[2 + 2 * r] : 0ll);
  printf("\n");
}
