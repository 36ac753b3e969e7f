// Direct fp32 3-D convolutions for the FEW-CHANNEL levels of the hourglass
// (reference regularization.py:74-126; Conv3d / ConvTranspose3d semantics of
// network_blocks.py:61-85, 106-131).  At 4-16 channels an implicit GEMM has
// N = Cout <= 16: a tcgen05 MMA of that width is operand-fetch bound (measured,
// tools/mma_microbench.cu) and, with split operands for fp32 accuracy, no
// faster than the FFMA pipe -- so these layers run on CUDA cores, organised so
// that the FFMA pipe (not shared memory, not HBM) is the limit:
//
//   * a CTA owns a 32 (x) by TY (y) column of output voxels and MARCHES along
//     z; the input planes it needs live in a 4-slot shared-memory ring filled
//     by cp.async (zero fill == the convolution's padding) one plane ahead of
//     the arithmetic, so every input voxel is read from L2/HBM ~1.2 times;
//   * shared-memory planes are stored per channel quad, [quad][row][x][4]:
//     the 32 lanes of a warp (consecutive x) read consecutive 16-byte vectors
//     (conflict-free LDS.128), weights are read as warp-wide broadcasts;
//   * a thread owns PX rows of one x column and all Cout channels
//     (PX * Cout = 32 accumulators); rows slide through registers so each
//     LDS.128 of activations feeds 3 * 4 * Cout FFMAs.
//
// Epilogue: bias + LeakyReLU(0.1), channels-last store, InstanceNorm partial
// sums (fp32 per thread -> double per CTA -> 2 * Cout atomics per CTA).
#include <string>

#include "conv_layers.cuh"

namespace pds {
namespace {

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct DirectParams {
  const float* in;      // [N][D][H][W][CIN]
  const float* w;       // kernel layout of ConvLayer: [tap 27][CIN][COUT]
  const float* bias;
  float* out;           // [N][D][H][W][COUT]
  double* stats;        // [N][COUT][2] or null
  int N, D, H, W, zseg, nseg, lrelu;
};

// Conv3d 3x3x3, stride 1, padding 1.
template <int CIN, int COUT, int PX>
__global__ void __launch_bounds__(128)
conv3d_k3s1_direct_kernel(const DirectParams p) {
  constexpr int TY = 4 * PX, ROWS = TY + 2, COLS = 34, QI = CIN / 4, QO = COUT / 4;
  constexpr int PLANE_V4 = QI * ROWS * COLS;          // float4 per ring slot
  extern __shared__ __align__(16) float4 smem4[];
  float4* ring = smem4;                                // [4][QI][ROWS][COLS]
  float4* wsm = smem4 + 4 * PLANE_V4;                  // [27][CIN][QO]
  __shared__ double red[4][2 * COUT];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * TY;
  const int n = blockIdx.z / p.nseg, seg = blockIdx.z - n * p.nseg;
  const int z0 = seg * p.zseg, z1 = min(p.D, z0 + p.zseg);

  for (int i = threadIdx.x; i < 27 * CIN * QO; i += 128)
    wsm[i] = reinterpret_cast<const float4*>(p.w)[i];

  const float* in_n = p.in + (size_t)n * p.D * p.H * p.W * CIN;
  auto load_plane = [&](int zi) {
    float4* dst = ring + ((zi + 1) & 3) * PLANE_V4;
    const float* src_z = in_n + (size_t)zi * p.H * p.W * CIN;
    for (int idx = threadIdx.x; idx < ROWS * COLS * QI; idx += 128) {
      const int q = idx % QI, px = (idx / QI) % COLS, row = idx / (QI * COLS);
      const int gy = y0 + row - 1, gx = x0 + px - 1;
      const bool ok = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
      const float* src = ok ? src_z + ((size_t)gy * p.W + gx) * CIN + 4 * q : in_n;
      cp_async16((uint32_t)__cvta_generic_to_shared(dst + (q * ROWS + row) * COLS + px), src, ok);
    }
  };

  // prologue: planes z0-1, z0, z0+1
  for (int zi = z0 - 1; zi <= z0 + 1; ++zi) {
    if (zi >= 0 && zi < p.D) load_plane(zi);
    cp_async_commit();
  }

  float bias[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) bias[c] = __ldg(p.bias + c);
  float s1[COUT], s2[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) { s1[c] = 0.f; s2[c] = 0.f; }

  const int x = x0 + lane;
  const int yb = y0 + warp * PX;          // first of this thread's PX rows

  for (int z = z0; z < z1; ++z) {
    __syncthreads();                      // everyone is done with plane z-2's slot
    if (z + 2 < p.D && z + 2 <= z1) load_plane(z + 2);
    cp_async_commit();
    cp_async_wait<1>();                   // planes <= z+1 have landed (this thread's copies)
    __syncthreads();

    float acc[PX][COUT];
#pragma unroll
    for (int r = 0; r < PX; ++r)
#pragma unroll
      for (int c = 0; c < COUT; ++c) acc[r][c] = 0.f;

#pragma unroll 1
    for (int dz = 0; dz < 3; ++dz) {
      const int zi = z + dz - 1;
      if (zi < 0 || zi >= p.D) continue;
      const float4* plane = ring + ((zi + 1) & 3) * PLANE_V4;
#pragma unroll 1
      for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
        for (int q = 0; q < QI; ++q) {
          float4 a[PX + 2];
#pragma unroll
          for (int r = 0; r < PX + 2; ++r) a[r] = plane[(q * ROWS + warp * PX + r) * COLS + lane + dx];
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const float4* wt = wsm + (((dz * 3 + dy) * 3 + dx) * CIN + 4 * q) * QO;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
              for (int cq = 0; cq < QO; ++cq) {
                const float4 wv = wt[k * QO + cq];
#pragma unroll
                for (int r = 0; r < PX; ++r) {
                  const float4 av = a[r + dy];
                  const float ak = k == 0 ? av.x : (k == 1 ? av.y : (k == 2 ? av.z : av.w));
                  acc[r][4 * cq + 0] = fmaf(ak, wv.x, acc[r][4 * cq + 0]);
                  acc[r][4 * cq + 1] = fmaf(ak, wv.y, acc[r][4 * cq + 1]);
                  acc[r][4 * cq + 2] = fmaf(ak, wv.z, acc[r][4 * cq + 2]);
                  acc[r][4 * cq + 3] = fmaf(ak, wv.w, acc[r][4 * cq + 3]);
                }
              }
            }
          }
        }
      }
    }

    // epilogue of plane z
#pragma unroll
    for (int r = 0; r < PX; ++r) {
      const int y = yb + r;
      const bool ok = x < p.W && y < p.H;
      float v[COUT];
#pragma unroll
      for (int c = 0; c < COUT; ++c) {
        float t = acc[r][c] + bias[c];
        if (p.lrelu) t = t > 0.f ? t : 0.1f * t;
        v[c] = t;
        if (ok) { s1[c] += t; s2[c] = fmaf(t, t, s2[c]); }
      }
      if (ok) {
        float4* o = reinterpret_cast<float4*>(p.out + ((((size_t)n * p.D + z) * p.H + y) * p.W + x) * COUT);
#pragma unroll
        for (int cq = 0; cq < QO; ++cq) o[cq] = make_float4(v[4 * cq], v[4 * cq + 1], v[4 * cq + 2], v[4 * cq + 3]);
      }
    }
  }

  if (p.stats) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
      double a = (double)s1[c], b = (double)s2[c];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (lane == 0) { red[warp][2 * c] = a; red[warp][2 * c + 1] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * COUT) {
      const double t = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
      atomicAdd(p.stats + (size_t)n * COUT * 2 + threadIdx.x, t);
    }
  }
}

template <int CIN, int COUT, int PX>
int launch_k3s1(const ConvLayer& l, const ConvGeom& g, const float* in, float* out, double* stats,
                cudaStream_t st) {
  constexpr int TY = 4 * PX;
  DirectParams p;
  p.in = in; p.w = l.w; p.bias = l.bias; p.out = out; p.stats = stats;
  p.N = g.N; p.D = g.D; p.H = g.H; p.W = g.W; p.lrelu = l.lrelu ? 1 : 0;
  const int tiles = ((g.W + 31) / 32) * ((g.H + TY - 1) / TY) * g.N;
  // z segments: enough CTAs for ~3 per SM, at least 2 planes each
  int zseg = g.D;
  while (zseg > 2 && tiles * ((g.D + zseg - 1) / zseg) < 3 * num_sms()) zseg = (zseg + 1) / 2;
  p.zseg = zseg;
  p.nseg = (g.D + zseg - 1) / zseg;
  const size_t smem = (size_t)(4 * (CIN / 4) * (TY + 2) * 34 + 27 * CIN * (COUT / 4)) * sizeof(float4);
  PDS_CUDA(allow_dynamic_smem(conv3d_k3s1_direct_kernel<CIN, COUT, PX>, (int)smem));
  dim3 grid((unsigned)((g.W + 31) / 32), (unsigned)((g.H + TY - 1) / TY), (unsigned)(g.N * p.nseg));
  if (grid.z > 65535) { set_error("conv3d_direct: grid too large"); return PDS_ERR_UNSUPPORTED; }
  static const std::string name = "conv3d_k3s1_direct<" + std::to_string(CIN) + "," + std::to_string(COUT) + ">";
  PDS_KERNEL(name.c_str(), st);
  PDS_KERNEL_WORK(2.0 * 27 * CIN * COUT * g.N * g.D * g.H * g.W, 4.0 * g.N * g.D * g.H * g.W * (CIN + COUT));
  conv3d_k3s1_direct_kernel<CIN, COUT, PX><<<grid, 128, smem, st>>>(p);
  PDS_LAUNCH_CHECK("conv3d_k3s1_direct_kernel");
  return PDS_OK;
}

}  // namespace

// Returns PDS_ERR_UNSUPPORTED (without setting an error) when no direct kernel serves the layer.
int conv_forward_direct(const ConvLayer& l, const ConvGeom& g, const float* in, float* out,
                        double* stats, cudaStream_t st, bool* handled) {
  *handled = false;
  const bool k3s1 = l.dim[0].mode == DM_CONV3 && l.dim[1].mode == DM_CONV3 && l.dim[2].mode == DM_CONV3 &&
                    l.dim[0].stride == 1 && l.dim[1].stride == 1 && l.dim[2].stride == 1 && g.n_div == 1;
  if (k3s1 && l.Cin == 8 && l.Cout == 8) { *handled = true; return launch_k3s1<8, 8, 4>(l, g, in, out, stats, st); }
  if (k3s1 && l.Cin == 16 && l.Cout == 16) { *handled = true; return launch_k3s1<16, 16, 2>(l, g, in, out, stats, st); }
  return PDS_OK;
}

}  // namespace pds
