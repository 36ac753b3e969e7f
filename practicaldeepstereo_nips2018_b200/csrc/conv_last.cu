// Last convolution of the matching operation (reference matching.py:91-93, 110-112: Conv2d 3x3
// 64 -> 8, bias, no activation) with the TAPS ON THE M AXIS of the tensor-core tile, sm_100a only.
//
// With eight output channels the implicit-GEMM form of conv_tc.cu (M = 128 pixels, N = 16) is bound
// by the A-operand fetch: 47 cycles per MMA however small N is (tools/mma_microbench.cu), 144 MMAs
// per 256 pixels, 183 us at 960x540 D=192.  Here the product is turned round:
//
//     Y[(tap, co), q] = sum_ci W[co][ci][tap] * X[q][ci]          for every pixel q of a haloed tile
//     out[p, co]      = bias[co] + sum_tap Y[(tap, co), p + offset(tap)]
//
// i.e. ONE GEMM per 16-channel chunk with M = 9 taps x 8 channels = 72 rows (padded to 128) of
// weights as the A operand and N = 128 haloed pixels (16 x 8) as the B operand: 12 MMAs of 64 cycles
// per 84 output pixels (14 x 6) instead of 47 MMAs of 47 cycles.  The activation tile is the same TMA
// box conv_tc.cu loads ([plane][y][x][16 B] = K-major core matrices); the shift-and-add over the
// taps happens in the epilogue through shared memory.
//   * fp32-grade split operands as everywhere (x = x0 + x1, w = w0 + w1 in fp16, weights x 2^8): the
//     products x0 w0 go to one TMEM accumulator, x0 w1 + x1 w0 to a second one, added in fp32.
//   * Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue in TWO GROUPS of four
//     (one warp per TMEM lane quarter) that take alternate tiles -- group g always drains accumulator
//     buffer g -- so that one group's shift-and-add overlaps the other group's TMEM drain and the next
//     tile's MMAs (the epilogue is a chain of short dependent steps: with a single group it took twice
//     the MMA time).  Phase 1: TMEM -> sum of the two orders -> Y in the group's shared-memory buffer;
//     phase 2: the 84 x 8 outputs, nine shared-memory reads each, straight into the (B, 8, D, H, W)
//     signatures.  Accumulators are double-buffered (2 x 256 TMEM columns); weights (32 KB) stay
//     resident; CTAs are persistent.
#include <stdlib.h>

#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace pds {
namespace {

using namespace ptx;

constexpr int kTW = 16, kTH = 8;                 // haloed tile (pixels): N = 128
constexpr int kOW = kTW - 2, kOH = kTH - 2;      // output tile: 14 x 6
constexpr int kNpix = kTW * kTH;                 // 128
constexpr int kRows = 128;                       // M: 72 weight rows (tap * 8 + co), zero rows above
constexpr int kStages = 14;                    // 8 KB each: the loads in flight must cover ~1.5 us of TMA latency (6 stages: 209 us, latency-bound)
constexpr int kEpiWarps = 8;
constexpr int kThreadsLast = 32 * (2 + kEpiWarps);
constexpr int kYPitch = kNpix + 4;               // floats per Y row (bank-conflict-free 128-bit stores)
constexpr uint32_t kTermBytes = 2 * kNpix * 16;  // one term of one 16-channel chunk of the tile: 4 KB
constexpr uint32_t kStageBytes = 2 * kTermBytes; // both terms
constexpr uint32_t kWTermBytes = 2 * kRows * 16; // one term of one chunk of the weights: 4 KB
constexpr int kChunks = 4;                       // 64 input channels
constexpr uint32_t kWBytes = kChunks * 2 * kWTermBytes;          // 32 KB
constexpr uint32_t kYBytes = 72 * kYPitch * 4;                   // 38 016 B
constexpr uint32_t kSmemLast = kWBytes + kStages * kStageBytes + 2 * kYBytes + 8 * (2 * kStages + 5) + 64 + 128;

// explicit shared-memory accesses (a pointer derived from the aligned dynamic buffer is generic: LD.E / ST.E)
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct alignas(64) LastParams {
  CUtensorMap map;          // activation planes [slice * 2 terms * 8 planes][H][W][8] fp16
  const uint16_t* w;        // [chunk][term][K half][row 128][8]
  const float* bias;        // [8]
  float* out;               // (B, 8, D, H, W)
  int n_slices, n_div, H, W, tiles_x, tiles_y;
  float inv_wscale;
};

// (8, 64, 3, 3) fp32 -> [chunk][term][K half][row = tap * 8 + co (128, zero above 72)][8] fp16 terms of w * wscale
__global__ void last_prepare_weights_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, float wscale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;     // over chunk, K half, row, e
  if (i >= kChunks * 2 * kRows * 8) return;
  const int e = i % 8, row = (i / 8) % kRows, j = (i / (8 * kRows)) % 2, c = i / (16 * kRows);
  const int tap = row / 8, co = row % 8, ci = c * 16 + j * 8 + e;
  const float x = row < 72 ? w[((size_t)co * 64 + ci) * 9 + tap] * wscale : 0.f;
  uint16_t t[3];
  split_terms<true>(x, t);
  for (int s = 0; s < 2; ++s)
    out[((((size_t)c * 2 + s) * 2 + j) * kRows + row) * 8 + e] = t[s];
}

__global__ void __launch_bounds__(kThreadsLast, 1)
conv_last_kernel(const __grid_constant__ LastParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  const uint32_t w_base = smem_u32(smem);
  const uint32_t stage_base = w_base + kWBytes;
  float* ysm_base = (float*)(smem + kWBytes + kStages * kStageBytes);
  const uint32_t bar_base = stage_base + kStages * kStageBytes + 2 * kYBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t wfull_bar = bar_base + 8u * (2 * kStages + 4);
  unsigned char* tail = smem + kWBytes + kStages * kStageBytes + 2 * kYBytes + 8 * (2 * kStages + 5);
  uint32_t* tmem_slot = (uint32_t*)(tail + 8);
  float* sbias = (float*)(tail + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 32 * kEpiWarps / 2); }
    mbar_init(wfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(wfull_bar, kWBytes);
    bulk_load(w_base, p.w, kWBytes, wfull_bar);
  }
  if (threadIdx.x < 8) sbias[threadIdx.x] = p.bias[threadIdx.x];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // every CTA walks one contiguous range of tiles
  const int tiles_per_slice = p.tiles_x * p.tiles_y;
  const int total_tiles = tiles_per_slice * p.n_slices;
  const int tiles_per_cta = (total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile_begin = min((int)blockIdx.x * tiles_per_cta, total_tiles);
  const int tile_end = min(tile_begin + tiles_per_cta, total_tiles);

  if (warp == 0) {
    // ===== TMA producer: one stage = both terms of one 16-channel chunk of the haloed tile =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int n = tile / tiles_per_slice, r = tile - n * tiles_per_slice;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        const int x0 = tx * kOW - 1, y0 = ty * kOH - 1;            // halo origin; out of bounds reads as zero padding
        for (int c = 0; c < kChunks; ++c) {
          mbar_wait(empty_bar(stage), phase ^ 1);
#ifdef PDS_LAST_DEBUG_HALF_TMA      // timing experiment: half the TMA requests (results are wrong)
          mbar_expect_tx(full_bar(stage), kStageBytes / 2);
          const uint32_t sa = stage_base + stage * kStageBytes;
          for (int s = 0; s < 1; ++s)
#else
          mbar_expect_tx(full_bar(stage), kStageBytes);
          const uint32_t sa = stage_base + stage * kStageBytes;
#pragma unroll
          for (int s = 0; s < 2; ++s)
#endif
            tma_load_4d(sa + s * kTermBytes, &p.map, 8 * x0, y0, (n * 2 + s) * 8 + 2 * c, 0, full_bar(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // instruction descriptor: kind::f16 (fp16 operands), fp32 accumulate, M = 128, N = 128, both K-major
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kNpix >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    mbar_wait(wfull_bar, 0);
    tc_fence_after();
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;      // acc = parity of the tile inside this CTA's range
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + acc * 256, d1 = d0 + 128;      // order 0 | order 1
      for (int c = 0; c < kChunks; ++c) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = stage_base + stage * kStageBytes;
          // A: weights [K half][row][16 B]: LBO = 128 rows x 16 B, SBO = 8 rows x 16 B
          const uint64_t w0 = umma_desc_kmajor(w_base + (c * 2 + 0) * kWTermBytes, kRows * 16, 128);
          const uint64_t w1 = umma_desc_kmajor(w_base + (c * 2 + 1) * kWTermBytes, kRows * 16, 128);
          // B: pixels [plane][y][x][16 B], rows dense: LBO = one plane, SBO = 8 pixels x 16 B
          const uint64_t x0 = umma_desc_kmajor(sa, kNpix * 16, 128);
          const uint64_t x1 = umma_desc_kmajor(sa + kTermBytes, kNpix * 16, 128);
          tc_mma(d0, w0, x0, idesc, c != 0 ? 1u : 0u);
          tc_mma(d1, w1, x0, idesc, c != 0 ? 1u : 0u);
          tc_mma(d1, w0, x1, idesc, 1u);
          tc_commit(empty_bar(stage));
          if (c == kChunks - 1) tc_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ===== epilogue warps 2..9: group g = (warp - 2) / 4 takes the CTA's tiles of parity g and always drains
    // accumulator buffer g; quarter q = TMEM lanes 32 q .. +31 (weight rows) =====
    const int q = warp & 3, g = (warp - 2) >> 2, e = ((warp - 2) & 3) * 32 + lane;
    const int row = 32 * q + lane;
    const size_t HW = (size_t)p.H * p.W;
    const uint32_t ysm = smem_u32(ysm_base) + g * kYBytes;     // this group's Y buffer (shared-memory address)
    const int p2_co = e / kOW, p2_ox = e - p2_co * kOW;          // phase 2: this thread's output column
    const float p2_bias = sbias[p2_co & 7];
    uint32_t acc_phase = 0;
    // tile coordinates advance incrementally (integer divisions per tile were a visible share of the epilogue)
    int n, ty, tx, b, d;
    {
      const int t0 = tile_begin + g;
      n = t0 / tiles_per_slice;
      const int r = t0 - n * tiles_per_slice;
      ty = r / p.tiles_x; tx = r - ty * p.tiles_x;
      b = n / p.n_div; d = n - b * p.n_div;
    }
    for (int tile = tile_begin + g; tile < tile_end; tile += 2) {
      mbar_wait(tfull_bar(g), acc_phase);
      tc_fence_after();
      // phase 1: Y[row][pixel] = order 1 + order 0 (smallest first) -> shared memory
      if (32 * q < 72) {        // warp-uniform: the TMEM loads are .sync.aligned; rows >= 72 are zero padding
        const uint32_t t_base = tmem_base + ((uint32_t)(32 * q) << 16) + g * 256;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float v[32], u[32];
          tmem_ld_issue<32>(t_base + 128 + 32 * k, v);
          tmem_ld_issue<32>(t_base + 32 * k, u);
          tmem_ld_wait();
          tmem_ld_fence<32>(v);
          tmem_ld_fence<32>(u);
          if (row < 72) {
            const uint32_t dst = ysm + (row * kYPitch + 32 * k) * 4;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts_v4(dst + 16 * j, v[4 * j] + u[4 * j], v[4 * j + 1] + u[4 * j + 1], v[4 * j + 2] + u[4 * j + 2],
                     v[4 * j + 3] + u[4 * j + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(g));
      acc_phase ^= 1;
      named_barrier(1 + g, 128);
      // phase 2: out[co][oy][ox] = bias + sum over taps of Y[tap * 8 + co][(oy + dy) * 16 + ox + dx].  A thread
      // owns one (co, ox) column of the tile: every shared-memory offset below is a compile-time constant
      // (the first version derived (co, oy, ox) per output and was bound by that integer arithmetic)
      if (e < 8 * kOW) {
        const int x = tx * kOW + p2_ox, y0 = ty * kOH;
        float* o = p.out + (((size_t)b * 8 + p2_co) * p.n_div + d) * HW + (size_t)y0 * p.W + x;
        const uint32_t yb = ysm + (p2_co * kYPitch + p2_ox) * 4;
#pragma unroll
        for (int oy = 0; oy < kOH; ++oy) {
          float sum = 0.f;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) sum += lds_f32(yb + (tap * 8 * kYPitch + (oy + tap / 3) * kTW + tap % 3) * 4);
          if (x < p.W && y0 + oy < p.H) o[(size_t)oy * p.W] = fmaf(sum, p.inv_wscale, p2_bias);
        }
      }
      named_barrier(1 + g, 128);        // the group's Y buffer is free for its next tile
      tx += 2;
      while (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
      while (ty >= p.tiles_y) { ty -= p.tiles_y; ++n; if (++d == p.n_div) { d = 0; ++b; } }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn_last() {
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

}  // namespace

size_t tc_last_weight_bytes() { return kWBytes; }

bool tc_last_enabled() {
  const char* e = getenv("PDS_B200_LAST_TAPS");
  return !(e && atoi(e) == 0);
}

int tc_last_prepare(const float* w_oihw, uint16_t* packed, float wscale, cudaStream_t st) {
  PDS_KERNEL("tc_prepare_weights", st);
  const int total = kChunks * 2 * kRows * 8;
  last_prepare_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(w_oihw, packed, wscale);
  PDS_LAUNCH_CHECK("last_prepare_weights_kernel");
  return PDS_OK;
}

// in: activation planes [n_slices][2 terms][8 planes][H][W][8] fp16; out: (B, 8, D, H, W) fp32 signatures
int tc_conv_last(const uint16_t* packed_w, const float* bias, float wscale, const uint16_t* in, float* out,
                 int n_slices, int n_div, int H, int W, cudaStream_t st) {
  if (n_slices == 0) return PDS_OK;
  EncodeTiledFn enc = encode_fn_last();
  if (!enc) { set_error("conv_last: cuTensorMapEncodeTiled is not available from this driver"); return PDS_ERR_UNSUPPORTED; }
  LastParams p = {};
  // (pixel x 8 channels) as ONE dimension: a box row of 16 pixels is a single 256-byte request
  const size_t planes = (size_t)n_slices * 2 * 8;
  const cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)planes, 1};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)planes * H * W * 16};
  const cuuint32_t box[4] = {(cuuint32_t)kTW * 8, (cuuint32_t)kTH, 2, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&p.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<uint16_t*>(in), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv_last: cuTensorMapEncodeTiled failed with CUresult %d (W=%d H=%d planes=%zu)", (int)r, W, H, planes);
    return PDS_ERR_CUDA;
  }
  p.w = packed_w; p.bias = bias; p.out = out;
  p.n_slices = n_slices; p.n_div = n_div > 0 ? n_div : 1; p.H = H; p.W = W;
  p.tiles_x = (W + kOW - 1) / kOW; p.tiles_y = (H + kOH - 1) / kOH;
  p.inv_wscale = 1.0f / wscale;
  PDS_CUDA(allow_dynamic_smem(conv_last_kernel, (int)kSmemLast));
  const int total = p.tiles_x * p.tiles_y * n_slices;
  const int grid = total < num_sms() ? total : num_sms();
  PDS_KERNEL("conv_last<taps on M, S=2>", st);
  {
    const double px = (double)H * W * n_slices;
    PDS_KERNEL_WORK(2.0 * 9 * 64 * 8 * px, px * (4.0 * 8 + 2.0 * 2 * 64));
  }
  conv_last_kernel<<<grid, kThreadsLast, kSmemLast, st>>>(p);
  PDS_LAUNCH_CHECK("conv_last_kernel");
  return PDS_OK;
}

}  // namespace pds
