// tcgen05 implicit-GEMM 3x3 convolution (stride 1, zero padding 1) for the
// matching operation (reference matching.py:69-112; Conv2d semantics of
// network_blocks.py:19-24, 47-58), sm_100a only.
//
// GEMM view per disparity slice: D[pixel, cout] = sum over (tap, cin) of
// A[pixel + tap, cin] * W[tap, cin, cout].
//   * One MMA tile = 128 pixels = 8 (x) by 16 (y); a CTA tile is NT such tiles
//     side by side (8*NT x 16 pixels) sharing ONE haloed input tile in shared
//     memory: (8*NT + 2) x 18 pixels per 16-channel K chunk, loaded once by TMA
//     (OOB zero fill == the convolution's zero padding; for Matching's shifted
//     right descriptor the box is simply placed at x - d, matching.py:56-60).
//   * Operands are K-major, no-swizzle UMMA tiles.  Activations live in HBM as
//     [plane = 8-channel group][y][x][8] bf16 so that a TMA box lands in shared
//     memory as [plane][y][x][16 B]: 8 consecutive pixels x 16 B is exactly one
//     UMMA core matrix, the next image row is the next core-matrix group
//     (SBO = halo pitch), the second 8-channel group is LBO away -- and a tap
//     (dy, dx) is nothing but a start-address offset of (dy*pitch + dx)*16 B.
//     No im2col, no per-tap reload: 9 taps x NT tiles of tcgen05.mma per chunk.
//   * fp32 accuracy on bf16 tensor cores: operands are carried as S bf16 terms
//     (x = hi + mid + lo); the kernel issues the leading partial products
//     (S=3: 6 of them, S=2: 3, S=1: 1) into the same fp32 TMEM accumulator.
//   * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread),
//     warps 2-5 = epilogue (TMEM -> registers -> bias / LeakyReLU /
//     InstanceNorm partial sums -> HBM).  Accumulators are double-buffered in
//     TMEM so the epilogue of tile k overlaps the MMAs of tile k+1; CTAs are
//     persistent (grid = #SMs) and walk the tile list with a static stride.
#include <string>

#include "conv_tc.cuh"

namespace pds {
namespace {

constexpr int kPH = 18;          // halo rows of a CTA tile (16 + 2)
constexpr int kThreads = 192;
constexpr int kMaxStages = 6;
constexpr int kMaxProducts = 6;

struct TcKernelParams {
  const CUtensorMap* maps;       // [0] primary input, [1 + d] right descriptor at disparity d
  const __nv_bfloat16* w;        // [s][chunk][tap][2][N][8]
  const float* bias;             // [N]
  float* out_f32;
  __nv_bfloat16* out_ap;
  float* out_sig;
  double* stats;
  int n_slices, n_div, H, W, tiles_x, tiles_y;
  int S, nprod, nchunks, nchunks1;   // chunks [0, nchunks1) come from maps[0]
  int planes1, planes2;              // 8-channel planes per (slice, split) of the two inputs
  int N, Cout, epilogue, sig_D;
  int stages;
  uint32_t a_plane_bytes, w_plane_bytes, stage_bytes;
  unsigned char prod_a[kMaxProducts], prod_b[kMaxProducts];
};

// ---- PTX wrappers ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug traps after ~4 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = global_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && global_ns() - t0 > 4000000000ull) __trap();
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, no-swizzle shared-memory matrix descriptor (sm_100 format, version 1):
// core matrix = 8 rows x 16 B contiguous; SBO = byte distance between 8-row
// groups, LBO = byte distance between the two 16-byte K halves of one MMA.
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// x = t0 + t1 + t2 with bf16 terms (round-to-nearest residual splitting)
__device__ __forceinline__ void split_bf16(float x, int S, __nv_bfloat16 (&t)[3]) {
  t[0] = __float2bfloat16_rn(x);
  float r = x - __bfloat162float(t[0]);
  t[1] = __float2bfloat16_rn(r);
  r -= __bfloat162float(t[1]);
  t[2] = __float2bfloat16_rn(r);
  (void)S;
}

// Sum over the 32 lanes of each of 32 per-lane values; lane l ends with channel l.
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool upper = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < step; ++i) {
      const float send = upper ? v[i] : v[i + step];
      const float keep = upper ? v[i + step] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return v[0];
}

template <int NT>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_tc_kernel(const TcKernelParams p) {
  constexpr int PW = 8 * NT + 2;                         // halo pitch in pixels
  constexpr uint32_t TMEM_COLS = NT == 1 ? 128 : (NT == 2 ? 256 : 512);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + (uint32_t)p.stages * p.stage_bytes;
  // barriers: full[stages], empty[stages], tfull[2], tempty[2]; then tmem pointer; then bias
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + 2 + a); };
  uint32_t* tmem_slot = (uint32_t*)(smem + (size_t)p.stages * p.stage_bytes + 8 * (2 * kMaxStages + 4));
  float* sbias = (float*)(tmem_slot + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.N; i += kThreads) sbias[i] = p.bias[i];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_slice = p.tiles_x * p.tiles_y;
  const int total_tiles = tiles_per_slice * p.n_slices;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_slice, r = tile - n * tiles_per_slice;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        const int x0 = tx * 8 * NT, y0 = ty * 16;
        const int n_in = n / p.n_div, d = n - n_in * p.n_div;
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), p.S * (p.a_plane_bytes + p.w_plane_bytes));
          const uint32_t sa = smem_base + stage * p.stage_bytes;
          const uint32_t sw = sa + p.S * p.a_plane_bytes;
          const bool second = c >= p.nchunks1;
          for (int s = 0; s < p.S; ++s) {
            if (!second)
              tma_load_4d(sa + s * p.a_plane_bytes, p.maps, 0, x0 - 1, y0 - 1,
                          (n_in * p.S + s) * p.planes1 + 2 * c, full_bar(stage));
            else  // d >= W: the shifted image is all zeros -> park the box fully out of bounds
              tma_load_4d(sa + s * p.a_plane_bytes, p.maps + 1 + d, 0,
                          d >= p.W ? -(PW + 16) : x0 - 1 - d, y0 - 1,
                          (n_in * p.S + s) * p.planes2 + 2 * (c - p.nchunks1), full_bar(stage));
            bulk_load(sw + s * p.w_plane_bytes,
                      p.w + ((size_t)s * p.nchunks + c) * (p.w_plane_bytes / 2), p.w_plane_bytes,
                      full_bar(stage));
          }
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      const uint32_t a_lbo = kPH * PW * 16, a_sbo = PW * 16;
      const uint32_t b_lbo = p.N * 16, b_sbo = 128, tap_bytes = 2 * p.N * 16;
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_base = tmem_base + acc * (NT * 64);
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * p.stage_bytes;
          const uint32_t sw = sa + p.S * p.a_plane_bytes;
          for (int q = 0; q < p.nprod; ++q) {
            const uint32_t a0 = sa + p.prod_a[q] * p.a_plane_bytes;
            const uint32_t w0 = sw + p.prod_b[q] * p.w_plane_bytes;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int dy = tap / 3, dx = tap % 3;
              const uint64_t bdesc = umma_desc(w0 + tap * tap_bytes, b_lbo, b_sbo);
#pragma unroll
              for (int i = 0; i < NT; ++i) {
                const uint64_t adesc = umma_desc(a0 + (dy * PW + 8 * i + dx) * 16, a_lbo, a_sbo);
                tc_mma_bf16(d_base + i * 64, adesc, bdesc, idesc, (c | q | tap) != 0 ? 1u : 0u);
              }
            }
          }
          tc_commit(empty_bar(stage));       // frees the smem stage once these MMAs retire
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(tfull_bar(acc));           // accumulator ready for the epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===== epilogue warps (2..5): TMEM lanes 32*(warp%4) .. +31 =====
    const int q = warp & 3;
    const int row = 32 * q + lane;
    const int px = row & 7, py = row >> 3;
    const size_t HW = (size_t)p.H * p.W;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n = tile / tiles_per_slice, r = tile - n * tiles_per_slice;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int y = ty * 16 + py;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(32 * q) << 16) + acc * (NT * 64);
#pragma unroll 1
      for (int i = 0; i < NT; ++i) {
        const int x = tx * 8 * NT + 8 * i + px;
        const bool valid = (x < p.W) && (y < p.H);
        const size_t pix = (size_t)y * p.W + x;
        if (p.epilogue == TC_EPI_SIG) {
          float v[16];
          tmem_ld16(t_base + i * 64, v);
          if (valid) {
            const int b = n / p.sig_D, dd = n - b * p.sig_D;
#pragma unroll
            for (int ch = 0; ch < 16; ++ch)
              if (ch < p.Cout)
                p.out_sig[(((size_t)b * p.Cout + ch) * p.sig_D + dd) * HW + pix] = v[ch] + sbias[ch];
          }
          continue;
        }
#pragma unroll 1
        for (int col0 = 0; col0 < p.N; col0 += 32) {
          float v[32];
          tmem_ld32(t_base + i * 64 + col0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float t = v[j] + sbias[col0 + j];
            if (p.epilogue == TC_EPI_ACT) t = t > 0.f ? t : 0.1f * t;
            v[j] = t;
          }
          if (valid) {
            float4* o = reinterpret_cast<float4*>(p.out_f32) +
                        ((size_t)n * (p.N / 4) + col0 / 4) * HW + pix;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              o[(size_t)k * HW] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          }
          if (p.epilogue == TC_EPI_PLAIN) {
            if (valid) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                __nv_bfloat16 t[8][3];
#pragma unroll
                for (int e = 0; e < 8; ++e) split_bf16(v[8 * g + e], p.S, t[e]);
                for (int s = 0; s < p.S; ++s) {
                  union { __nv_bfloat16 h[8]; uint4 u; } pk;
#pragma unroll
                  for (int e = 0; e < 8; ++e) pk.h[e] = t[e][s];
                  reinterpret_cast<uint4*>(p.out_ap)[((size_t)(n * p.S + s) * (p.N / 8) + col0 / 8 + g) * HW + pix] = pk.u;
                }
              }
            }
          } else {
            // InstanceNorm partial sums over this warp's 32 pixels, one channel per lane
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (!valid) v[j] = 0.f;
              sq[j] = v[j] * v[j];
            }
            const float s1 = warp_transpose_reduce(v, lane);
            const float s2 = warp_transpose_reduce(sq, lane);
            double* dst = p.stats + ((size_t)n * p.N + col0 + lane) * 2;
            atomicAdd(dst, (double)s1);
            atomicAdd(dst + 1, (double)s2);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- auxiliary kernels ------------------------------------------------------------------

// (Cout, Cin, 3, 3) fp32 -> [s][chunk][tap][2][N][8] split bf16 (zero rows for cout >= Cout)
__global__ void tc_prepare_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                          int Cout, int Cin, int N, int S) {
  const int nchunks = Cin / 16;
  const size_t per_split = (size_t)nchunks * 9 * 2 * N * 8;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_split) return;
  const int e = i % 8;
  const int co = (i / 8) % N;
  const int j = (i / (8 * (size_t)N)) % 2;
  const int tap = (i / (16 * (size_t)N)) % 9;
  const int c = i / (144 * (size_t)N);
  const int ci = c * 16 + j * 8 + e;
  const float x = co < Cout ? w[((size_t)co * Cin + ci) * 9 + tap] : 0.f;
  __nv_bfloat16 t[3];
  split_bf16(x, S, t);
  for (int s = 0; s < S; ++s) out[s * per_split + i] = t[s];
}

__global__ void tc_pad_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int Cout, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = i < Cout ? b[i] : 0.f;
}

// (B, C, H, W) fp32 -> AP [B][S][C/8][H][W][8]
__global__ void tc_pack_nchw_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ ap,
                                    int C, size_t HW, int S, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t pix = i % HW;
  const int c8 = (i / HW) % (C / 8);
  const size_t b = i / (HW * (C / 8));
  __nv_bfloat16 t[8][3];
#pragma unroll
  for (int e = 0; e < 8; ++e) split_bf16(in[(b * C + c8 * 8 + e) * HW + pix], S, t[e]);
  for (int s = 0; s < S; ++s) {
    union { __nv_bfloat16 h[8]; uint4 u; } pk;
#pragma unroll
    for (int e = 0; e < 8; ++e) pk.h[e] = t[e][s];
    reinterpret_cast<uint4*>(ap)[((b * S + s) * (C / 8) + c8) * HW + pix] = pk.u;
  }
}

// InstanceNorm from accumulated sums on fp32 planes; one thread = 8 channels of a pixel.
__global__ void __launch_bounds__(256)
tc_norm_split_kernel(const float* __restrict__ y, const double* __restrict__ stats,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ residual, float* __restrict__ out_f32,
                     __nv_bfloat16* __restrict__ out_ap, int C, size_t HW, int S) {
  const int c8 = blockIdx.y, n = blockIdx.z;
  __shared__ float sc[8], sh[8];
  if (threadIdx.x < 8) {
    const int c = c8 * 8 + threadIdx.x;
    const double s = stats[((size_t)n * C + c) * 2], q = stats[((size_t)n * C + c) * 2 + 1];
    const double mean = s / (double)HW;
    double var = q / (double)HW - mean * mean;
    if (var < 0.0) var = 0.0;
    sc[threadIdx.x] = (float)mean;
    sh[threadIdx.x] = (float)(1.0 / sqrt(var + 1e-5));
  }
  __syncthreads();
  float g[8], b[8], m[8], r[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    g[e] = gamma[c8 * 8 + e]; b[e] = beta[c8 * 8 + e]; m[e] = sc[e]; r[e] = sh[e];
  }
  const float4* y4 = reinterpret_cast<const float4*>(y) + ((size_t)n * (C / 4) + 2 * c8) * HW;
  const float4* r4 = residual ? reinterpret_cast<const float4*>(residual) + ((size_t)n * (C / 4) + 2 * c8) * HW : nullptr;
  float4* o4 = out_f32 ? reinterpret_cast<float4*>(out_f32) + ((size_t)n * (C / 4) + 2 * c8) * HW : nullptr;
  for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < HW;
       pix += (size_t)gridDim.x * blockDim.x) {
    const float4 a = y4[pix], c = y4[HW + pix];
    float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (v[e] - m[e]) * r[e] * g[e] + b[e];
    if (r4) {
      const float4 ra = r4[pix], rc = r4[HW + pix];
      v[0] += ra.x; v[1] += ra.y; v[2] += ra.z; v[3] += ra.w;
      v[4] += rc.x; v[5] += rc.y; v[6] += rc.z; v[7] += rc.w;
    }
    if (o4) {
      o4[pix] = make_float4(v[0], v[1], v[2], v[3]);
      o4[HW + pix] = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (out_ap) {
      __nv_bfloat16 t[8][3];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_bf16(v[e], S, t[e]);
      for (int s = 0; s < S; ++s) {
        union { __nv_bfloat16 h[8]; uint4 u; } pk;
#pragma unroll
        for (int e = 0; e < 8; ++e) pk.h[e] = t[e][s];
        reinterpret_cast<uint4*>(out_ap)[((size_t)(n * S + s) * (C / 8) + c8) * HW + pix] = pk.u;
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

// AP tensor viewed as (8 ch, W_eff, H, planes) with row pitch W.
int encode_ap_map(CUtensorMap* map, const void* base, int W_eff, int W, int H, size_t planes,
                  int PW) {
  const cuuint64_t dims[4] = {8, (cuuint64_t)W_eff, (cuuint64_t)H, (cuuint64_t)planes};
  const cuuint64_t strides[3] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
  const cuuint32_t box[4] = {8, (cuuint32_t)PW, (cuuint32_t)kPH, 2};
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

template <int NT>
int launch_tc(TcKernelParams& p, cudaStream_t st) {
  constexpr int PW = 8 * NT + 2;
  p.a_plane_bytes = 2 * kPH * PW * 16;
  p.w_plane_bytes = 9 * 2 * p.N * 16;
  p.stage_bytes = (uint32_t)align_up((size_t)p.S * (p.a_plane_bytes + p.w_plane_bytes), 128);
  const size_t tail = 8 * (2 * kMaxStages + 4) + 16 + 64 * sizeof(float) + 256;
  const size_t budget = 227 * 1024 - tail;
  int stages = (int)(budget / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) {
    set_error("conv3x3_tc: stage of %u bytes does not fit twice in shared memory", p.stage_bytes);
    return PDS_ERR_UNSUPPORTED;
  }
  p.stages = stages;
  const size_t smem = (size_t)stages * p.stage_bytes + tail;
  PDS_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int total = p.tiles_x * p.tiles_y * p.n_slices;
  const int grid = total < num_sms() ? total : num_sms();
  static const std::string name = "conv3x3_tc<NT=" + std::to_string(NT) + ">";
  PDS_KERNEL(name.c_str(), st);
  conv3x3_tc_kernel<NT><<<grid, kThreads, smem, st>>>(p);
  PDS_LAUNCH_CHECK("conv3x3_tc_kernel");
  return PDS_OK;
}

}  // namespace

bool tc_available() { return encode_fn() != nullptr; }

size_t tc_conv_max_maps(int n_div) { return (size_t)1 + (n_div > 0 ? n_div : 0); }

int tc_prepare_weights(const TcLayer& l, const float* w_oihw, const float* bias, cudaStream_t st) {
  const size_t per_split = l.w_elems() / l.S;
  {
    PDS_KERNEL("tc_prepare_weights", st);
    tc_prepare_weights_kernel<<<(unsigned)((per_split + 255) / 256), 256, 0, st>>>(w_oihw, l.w, l.Cout, l.Cin, l.N, l.S);
    PDS_LAUNCH_CHECK("tc_prepare_weights_kernel");
  }
  PDS_KERNEL("tc_pad_bias", st);
  tc_pad_bias_kernel<<<1, 256, 0, st>>>(bias, l.bias, l.Cout, l.N);
  PDS_LAUNCH_CHECK("tc_pad_bias_kernel");
  return PDS_OK;
}

int tc_pack_nchw(const float* in, __nv_bfloat16* ap, int B, int C, int H, int W, int S, cudaStream_t st) {
  const size_t HW = (size_t)H * W, total = (size_t)B * (C / 8) * HW;
  if (total == 0) return PDS_OK;
  PDS_KERNEL("tc_pack_nchw", st);
  tc_pack_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, ap, C, HW, S, total);
  PDS_LAUNCH_CHECK("tc_pack_nchw_kernel");
  return PDS_OK;
}

int tc_norm_split(const float* y, const double* stats, const float* gamma, const float* beta,
                  const float* residual, float* out_f32, __nv_bfloat16* out_ap, int n_slices, int C,
                  int H, int W, int S, cudaStream_t st) {
  const size_t HW = (size_t)H * W;
  if (n_slices == 0 || HW == 0) return PDS_OK;
  unsigned gx = (unsigned)((HW + 255) / 256);
  if (gx > 64) gx = 64;
  dim3 grid(gx, (unsigned)(C / 8), (unsigned)n_slices);
  PDS_KERNEL("tc_norm_split", st);
  tc_norm_split_kernel<<<grid, 256, 0, st>>>(y, stats, gamma, beta, residual, out_f32, out_ap, C, HW, S);
  PDS_LAUNCH_CHECK("tc_norm_split_kernel");
  return PDS_OK;
}

int tc_conv3x3(const TcConvArgs& a, cudaStream_t st) {
  const TcLayer& l = *a.layer;
  if (!encode_fn()) {
    set_error("conv3x3_tc: cuTensorMapEncodeTiled is not available from this driver");
    return PDS_ERR_UNSUPPORTED;
  }
  if (a.n_slices == 0) return PDS_OK;
  TcKernelParams p = {};
  p.maps = a.maps_dev;
  p.w = l.w; p.bias = l.bias;
  p.out_f32 = a.out_f32; p.out_ap = a.out_ap; p.out_sig = a.out_sig; p.stats = a.stats;
  p.n_slices = a.n_slices; p.n_div = a.n_div > 0 ? a.n_div : 1; p.H = a.H; p.W = a.W;
  p.S = l.S;
  static const unsigned char pa[3][6] = {{0}, {0, 0, 1}, {0, 0, 1, 1, 0, 2}};
  static const unsigned char pb[3][6] = {{0}, {0, 1, 0}, {0, 1, 0, 1, 2, 0}};
  p.nprod = l.S == 1 ? 1 : (l.S == 2 ? 3 : 6);
  for (int i = 0; i < p.nprod; ++i) { p.prod_a[i] = pa[l.S - 1][i]; p.prod_b[i] = pb[l.S - 1][i]; }
  p.nchunks = l.Cin / 16;
  p.nchunks1 = a.in_C / 16;
  p.planes1 = a.in_C / 8;
  p.planes2 = a.in2 ? a.in2_C / 8 : 0;
  p.N = l.N; p.Cout = l.Cout; p.epilogue = a.epilogue;
  p.sig_D = a.epilogue == TC_EPI_SIG ? (a.n_div > 0 ? a.n_div : 1) : 1;
  if (a.epilogue == TC_EPI_SIG) p.n_div = 1;   // sig_D only drives the output indexing
  if ((a.in2 ? a.in_C + a.in2_C : a.in_C) != l.Cin || l.Cin % 16 || l.N % 16 || l.N > 64 ||
      (a.epilogue != TC_EPI_SIG && l.N % 32)) {
    set_error("conv3x3_tc: unsupported channel configuration (Cin=%d, N=%d)", l.Cin, l.N);
    return PDS_ERR_UNSUPPORTED;
  }
  // tile width: the widest NT in {3, 2, 1} that wastes the fewest columns; the split
  // planes of a stage must fit shared memory at least twice (S=3 -> NT <= 3)
  const int w8 = (a.W + 7) / 8;
  int NT = 1, best_waste = 1 << 30;
  for (int nt = 3; nt >= 1; --nt) {
    const int waste = ((w8 + nt - 1) / nt) * nt - w8;
    if (waste < best_waste) { best_waste = waste; NT = nt; }
  }
  p.tiles_x = (w8 + NT - 1) / NT;
  p.tiles_y = (a.H + 15) / 16;
  const int PW = 8 * NT + 2;
  // tensor maps: [0] = primary input; [1 + d] = right descriptors, width W - d so that
  // columns at or beyond the shifted image's right edge read as zero padding
  int rc = encode_ap_map(&a.maps_host[0], a.in, a.W, a.W, a.H, (size_t)a.in_slices * l.S * (a.in_C / 8), PW);
  if (rc != PDS_OK) return rc;
  size_t nmaps = 1;
  if (a.in2) {
    for (int d = 0; d < p.n_div; ++d) {
      const int weff = a.W - d > 0 ? a.W - d : 1;
      rc = encode_ap_map(&a.maps_host[1 + d], a.in2, weff, a.W, a.H,
                         (size_t)a.in_slices * l.S * (a.in2_C / 8), PW);
      if (rc != PDS_OK) return rc;
    }
    nmaps += p.n_div;
  }
  PDS_CUDA(cudaMemcpyAsync(a.maps_dev, a.maps_host, nmaps * sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
  switch (NT) {
    case 1: return launch_tc<1>(p, st);
    case 2: return launch_tc<2>(p, st);
    default: return launch_tc<3>(p, st);
  }
}

}  // namespace pds
