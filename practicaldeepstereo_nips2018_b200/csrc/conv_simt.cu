// fp32 CUDA-core implicit-GEMM convolution (PDS_PRECISION_FP32 path) plus the
// InstanceNorm-apply and layout kernels shared by every precision.
//
// Restates, for channels-last activations, the ATen operators the reference
// composes (network_blocks.py:9-85): Conv2d/Conv3d (k3 or k5, stride 1|2),
// ConvTranspose3d (k4 s2 p1 and (3,4,4)/(1,2,2) p1), LeakyReLU(0.1) and the
// statistics of InstanceNorm.  GEMM view: rows = output voxels of one parity
// class, columns = output channels, K = taps x input channels.  fp32 FFMA with
// fp32 accumulation; InstanceNorm sums are accumulated in double.
#include <string>

#include "conv_layers.cuh"

namespace pds {
namespace {

struct ConvParams {
  const float* in;
  const float* in2;
  const float* w;
  const float* bias;
  float* out;
  double* stats;
  int N, n_div, Cin, Cin1, Cout;
  int D, H, W, ZD, ZH, ZW, OD, OH, OW;
  DimSpec dim[3];
  int lrelu, tiles_n;
};

__device__ __forceinline__ void lin_form(const DimSpec& d, int cls, int& zs, int& ts, int& off) {
  switch (d.mode) {
    case DM_CONV3: zs = d.stride; ts = 1; off = -1; break;
    case DM_CONV5: zs = d.stride; ts = 1; off = -2; break;
    case DM_UNIT: zs = 1; ts = 0; off = 0; break;
    case DM_TCONV4: zs = 1; ts = -1; off = cls; break;
    default: zs = 1; ts = 1; off = -1; break;  // DM_TCONV3
  }
}

// TN couts x TM voxels per thread, NG cout groups per CTA, KC channels per
// K-chunk.  128 threads; BM = (128 / NG) * TM voxels, BN = NG * TN couts.
template <int TN, int NG, int TM, int KC>
__global__ void __launch_bounds__(128) conv_igemm_f32(const ConvParams p) {
  constexpr int MG = 128 / NG, BM = MG * TM, BN = NG * TN;
  constexpr int AV = (KC % 4 == 0) ? 4 : 1;            // floats per A load
  constexpr int A_LD = (BM * KC / AV + 127) / 128;      // A loads per thread per chunk
  constexpr int ROWS = BM >= 128 ? BM / 128 : 1;
  constexpr int W_LD = (KC * BN + 127) / 128;
  static_assert(TM % 4 == 0, "TM must be a multiple of 4");
  static_assert((BM * KC / AV) % 128 == 0 || BM * KC / AV < 128, "A chunk / thread mismatch");

  __shared__ __align__(16) float As[2][KC][BM];
  __shared__ __align__(16) float Ws[2][KC][BN];

  const int tid = threadIdx.x;
  const int mg = tid % MG, ng = tid / MG;
  const int tile = blockIdx.x;
  const int nt = blockIdx.y % p.tiles_n, cls = blockIdx.y / p.tiles_n;
  const int n = blockIdx.z;
  const int n0 = nt * BN;
  const int ncw = p.dim[2].nclass(), nch = p.dim[1].nclass();
  const int cls_w = cls % ncw, cls_h = (cls / ncw) % nch, cls_d = cls / (ncw * nch);
  const int ntw = p.dim[2].ntaps(), nth = p.dim[1].ntaps(), ntd = p.dim[0].ntaps();
  const int ntaps = ntw * nth * ntd;
  int zs[3], ts[3], off[3];
  lin_form(p.dim[0], cls_d, zs[0], ts[0], off[0]);
  lin_form(p.dim[1], cls_h, zs[1], ts[1], off[1]);
  lin_form(p.dim[2], cls_w, zs[2], ts[2], off[2]);

  const int Mz = p.ZD * p.ZH * p.ZW;
  const int n_in = n / p.n_div, shift = n % p.n_div;

  // z coordinates of the rows this thread stages
  int rzd[ROWS], rzh[ROWS], rzw[ROWS];
  bool rok[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int ml = (BM >= 128) ? tid + 128 * r : tid % BM;
    const int mz = tile * BM + ml;
    rok[r] = mz < Mz;
    const int mm = rok[r] ? mz : 0;
    rzw[r] = mm % p.ZW;
    rzh[r] = (mm / p.ZW) % p.ZH;
    rzd[r] = mm / (p.ZW * p.ZH);
  }

  const int cpt = p.Cin / KC;             // chunks per tap
  const int nchunks = ntaps * cpt;

  float areg[A_LD][AV];
  float wreg[W_LD];

  auto load_chunk = [&](int c) {
    const int tap = c / cpt, c0 = (c - tap * cpt) * KC;
    const int tw = tap % ntw, th = (tap / ntw) % nth, td = tap / (ntw * nth);
    const bool second = c0 >= p.Cin1;
    const float* src = second ? p.in2 : p.in;
    const int Cs = second ? p.Cin - p.Cin1 : p.Cin1;
    const int cb = second ? c0 - p.Cin1 : c0;
    const int sh = second ? shift : 0;
#pragma unroll
    for (int it = 0; it < A_LD; ++it) {
      const int f = it * 128 + tid;
      const int kq = f / BM;
      const int r = (BM >= 128) ? it % ROWS : 0;
      bool ok = rok[r] && (BM * KC / AV >= 128 || f < BM * KC / AV);
      const int id = rzd[r] * zs[0] + td * ts[0] + off[0];
      const int ih = rzh[r] * zs[1] + th * ts[1] + off[1];
      const int iw = rzw[r] * zs[2] + tw * ts[2] + off[2];
      ok = ok && id >= 0 && id < p.D && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W && iw - sh >= 0;
      if (ok) {
        const float* q = src + ((((size_t)n_in * p.D + id) * p.H + ih) * p.W + (iw - sh)) * Cs + cb + kq * AV;
        if (AV == 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(q));
          areg[it][0] = v.x; areg[it][1 % AV] = v.y; areg[it][2 % AV] = v.z; areg[it][3 % AV] = v.w;
        } else {
          areg[it][0] = __ldg(q);
        }
      } else {
#pragma unroll
        for (int e = 0; e < AV; ++e) areg[it][e] = 0.f;
      }
    }
    const float* wsrc = p.w + (((size_t)cls * ntaps + tap) * p.Cin + c0) * p.Cout + n0;
#pragma unroll
    for (int it = 0; it < W_LD; ++it) {
      const int e = it * 128 + tid;
      const int k = e / BN, j = e % BN;
      wreg[it] = (e < KC * BN && n0 + j < p.Cout) ? __ldg(wsrc + (size_t)k * p.Cout + j) : 0.f;
    }
  };
  auto store_chunk = [&](int buf) {
#pragma unroll
    for (int it = 0; it < A_LD; ++it) {
      const int f = it * 128 + tid;
      const int ml = f % BM, kq = f / BM;
      if (BM * KC / AV >= 128 || f < BM * KC / AV) {
#pragma unroll
        for (int e = 0; e < AV; ++e) As[buf][kq * AV + e][ml] = areg[it][e];
      }
    }
#pragma unroll
    for (int it = 0; it < W_LD; ++it) {
      const int e = it * 128 + tid;
      if (e < KC * BN) Ws[buf][e / BN][e % BN] = wreg[it];
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_chunk(0);
  store_chunk(0);
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) load_chunk(c + 1);
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < TM / 4; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][g * MG * 4 + mg * 4]);
        a[4 * g] = v.x; a[4 * g + 1] = v.y; a[4 * g + 2] = v.z; a[4 * g + 3] = v.w;
      }
      if (TN % 4 == 0) {
#pragma unroll
        for (int j = 0; j < TN / 4; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(&Ws[buf][k][ng * TN + 4 * j]);
          b[4 * j] = v.x; b[(4 * j + 1) % TN] = v.y; b[(4 * j + 2) % TN] = v.z; b[(4 * j + 3) % TN] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Ws[buf][k][ng * TN + j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (c + 1 < nchunks) store_chunk(buf ^ 1);
    __syncthreads();
  }

  // epilogue: bias, LeakyReLU, channels-last store, InstanceNorm partial sums
  float bias[TN];
  double s[TN], ss[TN];  // InstanceNorm sums: double from the first addition on
  const int co0 = n0 + ng * TN;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    bias[j] = (co0 + j < p.Cout) ? __ldg(p.bias + co0 + j) : 0.f;
    s[j] = 0.0; ss[j] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int ml = (i / 4) * MG * 4 + mg * 4 + (i % 4);
    const int mz = tile * BM + ml;
    if (mz >= Mz) continue;
    const int zw = mz % p.ZW, zh = (mz / p.ZW) % p.ZH, zd = mz / (p.ZW * p.ZH);
    const int od = p.dim[0].out_coord(zd, cls_d), oh = p.dim[1].out_coord(zh, cls_h),
              ow = p.dim[2].out_coord(zw, cls_w);
    float* o = p.out + ((((size_t)n * p.OD + od) * p.OH + oh) * p.OW + ow) * p.Cout + co0;
    float v[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      float x = acc[i][j] + bias[j];
      if (p.lrelu) x = x > 0.f ? x : 0.1f * x;
      v[j] = x; s[j] += (double)x; ss[j] += (double)x * (double)x;
    }
    if (TN % 4 == 0 && (p.Cout % 4 == 0)) {
#pragma unroll
      for (int j = 0; j < TN / 4; ++j)
        if (co0 + 4 * j < p.Cout)
          *reinterpret_cast<float4*>(o + 4 * j) =
              make_float4(v[4 * j], v[(4 * j + 1) % TN], v[(4 * j + 2) % TN], v[(4 * j + 3) % TN]);
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j) if (co0 + j < p.Cout) o[j] = v[j];
    }
  }
  if (p.stats) {
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      double ds = s[j], dss = ss[j];
#pragma unroll
      for (int o = (MG >= 32 ? 16 : MG / 2); o >= 1; o >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
        dss += __shfl_xor_sync(0xffffffffu, dss, o);
      }
      if ((tid % (MG >= 32 ? 32 : MG)) == 0 && co0 + j < p.Cout) {
        double* d = p.stats + ((size_t)n * p.Cout + co0 + j) * 2;
        atomicAdd(d, ds);
        atomicAdd(d + 1, dss);
      }
    }
  }
}

template <int TN, int NG, int TM, int KC>
int launch_conv(const ConvParams& p, cudaStream_t st) {
  constexpr int BM = (128 / NG) * TM, BN = NG * TN;
  ConvParams q = p;
  q.tiles_n = (p.Cout + BN - 1) / BN;
  const int Mz = p.ZD * p.ZH * p.ZW;
  const int ncls = p.dim[0].nclass() * p.dim[1].nclass() * p.dim[2].nclass();
  dim3 grid((unsigned)((Mz + BM - 1) / BM), (unsigned)(q.tiles_n * ncls), (unsigned)p.N);
  if (grid.y > 65535 || grid.z > 65535) {
    set_error("conv_forward_simt: grid too large");
    return PDS_ERR_UNSUPPORTED;
  }
  static const std::string name = "conv_igemm_f32<" + std::to_string(TN) + "," + std::to_string(NG) + "," +
                                  std::to_string(TM) + "," + std::to_string(KC) + ">";
  PDS_KERNEL(name.c_str(), st);
  {
    const double zvox = (double)Mz * ncls * p.N, taps = p.dim[0].ntaps() * p.dim[1].ntaps() * p.dim[2].ntaps();
    PDS_KERNEL_WORK(2.0 * taps * p.Cin * p.Cout * zvox,
                    4.0 * ((double)p.N / p.n_div * p.D * p.H * p.W * p.Cin + zvox * p.Cout));
  }
  conv_igemm_f32<TN, NG, TM, KC><<<grid, 128, 0, st>>>(q);
  PDS_LAUNCH_CHECK("conv_igemm_f32");
  return PDS_OK;
}

__global__ void relayout_weights_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                        DimSpec dd, DimSpec dh, DimSpec dw, int Cin, int Cout,
                                        int transposed, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = i % Cout;
  const int ci = (i / Cout) % Cin;
  size_t r = i / ((size_t)Cout * Cin);
  const int ntw = dw.ntaps(), nth = dh.ntaps(), ntd = dd.ntaps();
  const int ntaps = ntw * nth * ntd;
  const int tap = r % ntaps, cls = r / ntaps;
  const int tw = tap % ntw, th = (tap / ntw) % nth, td = tap / (ntw * nth);
  const int ncw = dw.nclass(), nch = dh.nclass();
  const int cw = cls % ncw, ch = (cls / ncw) % nch, cd = cls / (ncw * nch);
  const int kd = dd.kidx(td, cd), kh = dh.kidx(th, ch), kw = dw.kidx(tw, cw);
  const int KD = dd.ksize(), KH = dh.ksize(), KW = dw.ksize();
  (void)KD;
  const size_t a = transposed ? ((size_t)ci * Cout + co) : ((size_t)co * Cin + ci);
  dst[i] = src[((a * KD + kd) * KH + kh) * KW + kw];
}

// [N][S][C] channels-last InstanceNorm apply; 4 channels per thread.
template <int V>
__global__ void __launch_bounds__(256)
instance_norm_apply_kernel(const float* __restrict__ y, const double* __restrict__ stats,
                           const float* __restrict__ gamma, const float* __restrict__ beta,
                           const float* __restrict__ add, const float* __restrict__ add_bcast,
                           float* __restrict__ out, float* __restrict__ out2, size_t S, size_t HW,
                           int C) {
  extern __shared__ float sm[];  // mean[C], rstd[C], gamma[C], beta[C]
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double s = stats[((size_t)n * C + c) * 2], q = stats[((size_t)n * C + c) * 2 + 1];
    const double mean = s / (double)S;
    double var = q / (double)S - mean * mean;
    if (var < 0.0) var = 0.0;
    sm[c] = (float)mean;
    sm[C + c] = (float)(1.0 / sqrt(var + 1e-5));
    sm[2 * C + c] = gamma ? gamma[c] : 1.f;
    sm[3 * C + c] = beta ? beta[c] : 0.f;
  }
  __syncthreads();
  const size_t total = S * C / V;
  const size_t base = (size_t)n * S * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i * V;
    const int c = (int)(e % C);
    float v[V], a[V];
    if (V == 4) {
      const float4 t = *reinterpret_cast<const float4*>(y + base + e);
      v[0] = t.x; v[1 % V] = t.y; v[2 % V] = t.z; v[3 % V] = t.w;
    } else {
      v[0] = y[base + e];
    }
#pragma unroll
    for (int k = 0; k < V; ++k)
      v[k] = (v[k] - sm[c + k]) * sm[C + c + k] * sm[2 * C + c + k] + sm[3 * C + c + k];
    if (out) {
      if (V == 4) *reinterpret_cast<float4*>(out + base + e) = make_float4(v[0], v[1 % V], v[2 % V], v[3 % V]);
      else out[base + e] = v[0];
    }
    if (out2) {
#pragma unroll
      for (int k = 0; k < V; ++k) a[k] = v[k];
      if (add) {
        if (V == 4) {
          const float4 t = *reinterpret_cast<const float4*>(add + base + e);
          a[0] += t.x; a[1 % V] += t.y; a[2 % V] += t.z; a[3 % V] += t.w;
        } else {
          a[0] += add[base + e];
        }
      }
      if (add_bcast) {
        const size_t vox = e / C;
        const float* b = add_bcast + ((size_t)n * HW + vox % HW) * C + c;
#pragma unroll
        for (int k = 0; k < V; ++k) a[k] += b[k];
      }
      if (V == 4) *reinterpret_cast<float4*>(out2 + base + e) = make_float4(a[0], a[1 % V], a[2 % V], a[3 % V]);
      else out2[base + e] = a[0];
    }
  }
}

// in [N][C][S] -> out [N][S][C] (TO_CL) or the reverse, 32 positions per CTA.
template <bool TO_CL>
__global__ void __launch_bounds__(256)
layout_kernel(const float* __restrict__ in, float* __restrict__ out, int C, size_t S) {
  extern __shared__ float tile[];  // [C][33]
  const int n = blockIdx.y;
  const size_t s0 = (size_t)blockIdx.x * 32;
  const int ns = (int)((S - s0) < 32 ? (S - s0) : 32);
  const float* src = in + (size_t)n * C * S;
  float* dst = out + (size_t)n * C * S;
  if (TO_CL) {
    for (int i = threadIdx.x; i < C * 32; i += blockDim.x) {
      const int c = i / 32, s = i % 32;
      if (s < ns) tile[c * 33 + s] = src[(size_t)c * S + s0 + s];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * ns; i += blockDim.x) {
      const int s = i / C, c = i % C;
      dst[(s0 + s) * C + c] = tile[c * 33 + s];
    }
  } else {
    for (int i = threadIdx.x; i < C * ns; i += blockDim.x) {
      const int s = i / C, c = i % C;
      tile[c * 33 + s] = src[(s0 + s) * C + c];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * 32; i += blockDim.x) {
      const int c = i / 32, s = i % 32;
      if (s < ns) dst[(size_t)c * S + s0 + s] = tile[c * 33 + s];
    }
  }
}

template <bool TO_CL>
int launch_layout(const float* in, float* out, int N, int C, size_t S, cudaStream_t st) {
  if (N == 0 || S == 0) return PDS_OK;
  const size_t smem = (size_t)C * 33 * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("layout conversion: too many channels (%d)", C);
    return PDS_ERR_UNSUPPORTED;
  }
  dim3 grid((unsigned)((S + 31) / 32), (unsigned)N);
  PDS_KERNEL(TO_CL ? "layout_to_channels_last" : "layout_to_channels_first", st);
  PDS_KERNEL_WORK(0, 8.0 * N * C * S);
  layout_kernel<TO_CL><<<grid, 256, smem, st>>>(in, out, C, S);
  PDS_LAUNCH_CHECK("layout_kernel");
  return PDS_OK;
}

}  // namespace

int relayout_weights(const ConvLayer& l, const float* src, float* dst, cudaStream_t st) {
  const size_t total = l.weight_elems();
  PDS_KERNEL("relayout_weights", st);
  relayout_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      src, dst, l.dim[0], l.dim[1], l.dim[2], l.Cin, l.Cout, l.transposed ? 1 : 0, total);
  PDS_LAUNCH_CHECK("relayout_weights_kernel");
  return PDS_OK;
}

int conv_forward_simt(const ConvLayer& l, const ConvGeom& g, const float* in, const float* in2,
                      int Cin1, float* out, double* stats, cudaStream_t st) {
  ConvParams p;
  p.in = in; p.in2 = in2; p.w = l.w; p.bias = l.bias; p.out = out; p.stats = stats;
  p.N = g.N; p.n_div = g.n_div; p.Cin = l.Cin; p.Cin1 = in2 ? Cin1 : l.Cin; p.Cout = l.Cout;
  p.D = g.D; p.H = g.H; p.W = g.W;
  p.ZD = l.dim[0].zsize(g.D); p.ZH = l.dim[1].zsize(g.H); p.ZW = l.dim[2].zsize(g.W);
  p.OD = l.dim[0].out_size(g.D); p.OH = l.dim[1].out_size(g.H); p.OW = l.dim[2].out_size(g.W);
  for (int i = 0; i < 3; ++i) p.dim[i] = l.dim[i];
  p.lrelu = l.lrelu ? 1 : 0;
  p.tiles_n = 1;
  if (g.N == 0 || p.ZD * p.ZH * p.ZW == 0) return PDS_OK;
  const size_t Mz = (size_t)p.ZD * p.ZH * p.ZW;
  if (Mz > 0x7fffffff) {
    set_error("conv_forward_simt: more than 2^31 voxels per sample");
    return PDS_ERR_UNSUPPORTED;
  }
  const bool k8 = (p.Cin % 8 == 0) && (p.Cin1 % 8 == 0);
  const size_t work = Mz * g.N * l.nclass();
  if (k8 && l.Cout % 8 == 0) {
    const int bn = l.Cout >= 64 ? 64 : l.Cout;
    const bool big = work >= (size_t)148 * 8 * 256;   // enough rows for the larger tiles
    switch (bn) {
      case 8: return launch_conv<8, 1, 4, 8>(p, st);
      case 16: return launch_conv<8, 2, 4, 8>(p, st);
      case 32: return big ? launch_conv<8, 4, 8, 8>(p, st) : launch_conv<8, 4, 4, 8>(p, st);
      case 64: return big ? launch_conv<8, 8, 8, 8>(p, st) : launch_conv<8, 8, 4, 8>(p, st);
      default: break;
    }
  }
  if (k8 && l.Cout == 4) return launch_conv<4, 1, 4, 8>(p, st);
  if (p.Cin % 4 == 0 && p.Cin1 % 4 == 0 && l.Cout == 1) return launch_conv<1, 1, 4, 4>(p, st);
  return launch_conv<1, 1, 4, 1>(p, st);  // any channel count (slow, scalar)
}

int instance_norm_apply(const float* y, const double* stats, const float* gamma, const float* beta,
                        const float* add, const float* add_bcast, float* out, float* out2, int N,
                        size_t S, size_t HW, int C, cudaStream_t st) {
  if (N == 0 || S == 0) return PDS_OK;
  const size_t smem = (size_t)4 * C * sizeof(float);
  const bool v4 = (C % 4 == 0);
  const size_t total = S * C / (v4 ? 4 : 1);
  unsigned gx = (unsigned)((total + 255) / 256);
  const unsigned cap = (unsigned)(num_sms() * 16);
  if (gx > cap) gx = cap;
  dim3 grid(gx, (unsigned)N);
  PDS_KERNEL("instance_norm_apply", st);
  PDS_KERNEL_WORK(0, 4.0 * N * S * C * (1 + (out ? 1 : 0) + (out2 ? 1 : 0) + (add ? 1 : 0)));
  if (v4)
    instance_norm_apply_kernel<4><<<grid, 256, smem, st>>>(y, stats, gamma, beta, add, add_bcast, out, out2, S, HW, C);
  else
    instance_norm_apply_kernel<1><<<grid, 256, smem, st>>>(y, stats, gamma, beta, add, add_bcast, out, out2, S, HW, C);
  PDS_LAUNCH_CHECK("instance_norm_apply_kernel");
  return PDS_OK;
}

int nchw_to_nhwc(const float* in, float* out, int N, int C, size_t S, cudaStream_t st) {
  return launch_layout<true>(in, out, N, C, S, st);
}
int nhwc_to_nchw(const float* in, float* out, int N, int C, size_t S, cudaStream_t st) {
  return launch_layout<false>(in, out, N, C, S, st);
}

}  // namespace pds
