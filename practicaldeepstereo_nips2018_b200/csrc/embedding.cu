// a6 -- Embedding.forward (reference embedding.py:46-65) on the tcgen05 convolution engine:
//   InstanceNorm2d(3, non-affine) -> 2 x [conv5x5 stride 2 + LeakyReLU + IN] -> residual blocks
//   -> descriptor; shortcut = conv3x3(64 -> 8) + LeakyReLU + IN of the descriptor.
// The residual blocks are the same 64 -> 64 blocks as the matching operation's and run on its
// kernel (conv_tc.cu) with the images as slices; the other layers run on the generic engine.
//
// The image is normalised and written straight into phase-separated operand planes with its 3
// channels zero-padded to one 8-channel group, so the first convolution runs on the tensor cores
// as well (25 taps paired two per K=16 MMA).  Every later block is one convolution launch (fp32
// channels-last + InstanceNorm sums) and one normalisation pass that produces the next operand
// planes, fused with the residual additions (network_blocks.py:143-144).  Left and right images
// are just samples of one batch; the shortcut block runs on the first `n_shortcut` samples only
// (the reference computes and discards the right image's, network.py:40).
#include <algorithm>
#include <new>
#include <vector>

#include "conv_layers.cuh"
#include "conv_tc.cuh"
#include "conv_tcg.cuh"
#include "tc_ptx.cuh"

struct pds_embedding {
  int Cin, F, Fs, n_res, precision, split, fp16;
  float* raw = nullptr;                    // parameters, PyTorch layout, state_dict order
  std::vector<const float*> raw_params;
  std::vector<pds::TcgLayer> layers;       // conv1, conv2, shortcut (generic engine; planned per extent)
  std::vector<pds::TcLayer> res;           // 2 per residual block: the matching operation's 3x3 kernel
  char* res_blob = nullptr;
  char* blob = nullptr;
  int shape[2] = {0, 0};
  // plans and weight images of other extents seen by this handle (see pds_regularization::Saved)
  struct Saved { int shape[2]; std::vector<pds::TcgLayer> layers; char* blob; };
  std::vector<Saved> saved;
};

namespace pds {
namespace {

using namespace ptx;

// f3 -- the input path: the images as the caller holds them (dataset.py:67-72: cv2's interleaved
// uint8, converted to float and permuted on the host by the reference; or float / uint8 planar
// tensors), un-padded.  SizeAdapter.pad (size_adapter.py:29-43: zeros on top / left up to the
// padded extent H x W) and the first InstanceNorm2d (embedding.py:32) happen while the operand
// planes are written; the left and the right batch are two pointers (no torch.cat).
struct ImageSrc {
  const void* a; const void* b;   // samples 0 .. n_a-1 from a, the rest from b
  int n_a, layout, C, h, w, pad_top, pad_left;
};

template <int LAYOUT>
__device__ __forceinline__ float image_at(const ImageSrc& s, int n, int c, int y, int x) {
  const bool first = n < s.n_a;
  const void* base = first ? s.a : s.b;
  const size_t m = (size_t)(first ? n : n - s.n_a) * s.C * s.h * s.w;
  if (LAYOUT == PDS_IMAGE_F32_NCHW) return __ldg((const float*)base + m + ((size_t)c * s.h + y) * s.w + x);
  if (LAYOUT == PDS_IMAGE_U8_NCHW) return (float)__ldg((const unsigned char*)base + m + ((size_t)c * s.h + y) * s.w + x);
  return (float)__ldg((const unsigned char*)base + m + ((size_t)y * s.w + x) * s.C + c);
}

// sum / sum of squares of every (sample, channel) plane (the padding contributes zeros)
template <int LAYOUT>
__global__ void __launch_bounds__(256)
image_stats_kernel(const ImageSrc src, double* __restrict__ stats) {
  const int n = blockIdx.y / src.C, c = blockIdx.y - n * src.C;
  const size_t hw = (size_t)src.h * src.w;
  double s = 0.0, q = 0.0;
  if (LAYOUT == PDS_IMAGE_F32_NCHW && hw % 4 == 0) {
    const bool first = n < src.n_a;
    const float* p = (const float*)(first ? src.a : src.b) + ((size_t)(first ? n : n - src.n_a) * src.C + c) * hw;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < hw; i += (size_t)gridDim.x * blockDim.x * 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p + i));
      s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
      q += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
  } else if (LAYOUT == PDS_IMAGE_F32_NCHW) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
      const float v = image_at<LAYOUT>(src, n, c, (int)(i / src.w), (int)(i % src.w));
      s += v; q += (double)v * v;
    }
  } else {
    // uint8: integer sums are exact (and equal to the double sums of the float path); a thread's
    // share stays far below 2^32 / 255^2 elements, the totals are carried in 64 bits
    const bool first = n < src.n_a;
    const unsigned char* p = (const unsigned char*)(first ? src.a : src.b) + (size_t)(first ? n : n - src.n_a) * src.C * hw;
    const size_t stride = LAYOUT == PDS_IMAGE_U8_NHWC ? (size_t)src.C : 1;
    p += LAYOUT == PDS_IMAGE_U8_NHWC ? (size_t)c : (size_t)c * hw;
    unsigned long long ts = 0, tq = 0;
    unsigned int us = 0, uq = 0, cnt = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
      const unsigned int v = __ldg(p + i * stride);
      us += v; uq += v * v;
      if (++cnt == 32768) { ts += us; tq += uq; us = uq = cnt = 0; }
    }
    s = (double)(ts + us); q = (double)(tq + uq);
  }
  __shared__ double rs[8], rq[8];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { s += rs[w]; q += rq[w]; }
    atomicAdd(stats + 2 * blockIdx.y, s);
    atomicAdd(stats + 2 * blockIdx.y + 1, q);
  }
}

// InstanceNorm2d(C <= 8, affine=False, eps 1e-5) of the PADDED image (statistics over H x W, the
// pad region normalises to -mean * rstd) -> phase-separated AP planes
// [n][S][4 phases][1 plane][H/2][W/2][8] with channels C..7 zero.  One thread = one pixel.
template <bool FP16, int LAYOUT>
__global__ void __launch_bounds__(256)
image_norm_to_ap_kernel(const ImageSrc src, const double* __restrict__ stats,
                        uint16_t* __restrict__ out, int H, int W, int S) {
  const int n = blockIdx.y, C = src.C;
  const size_t HW = (size_t)H * W;
  __shared__ float sc[8], sh[8];
  if (threadIdx.x < 8) {
    float a = 0.f, b = 0.f;
    if ((int)threadIdx.x < C) {
      const double s = stats[((size_t)n * C + threadIdx.x) * 2], q = stats[((size_t)n * C + threadIdx.x) * 2 + 1];
      const double mean = s / (double)HW;
      double var = q / (double)HW - mean * mean;
      if (var < 0.0) var = 0.0;
      a = (float)(1.0 / sqrt(var + 1e-5));
      b = -(float)mean * a;
    }
    sc[threadIdx.x] = a; sh[threadIdx.x] = b;
  }
  __syncthreads();
  const int IY = H / 2, IX = W / 2;
  const size_t plane = (size_t)IY * IX;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)(i / W);
    const int sy = y - src.pad_top, sx = x - src.pad_left;
    const bool inside = sy >= 0 && sx >= 0;
    uint16_t t[8][3];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float v = c < C ? fmaf(inside ? image_at<LAYOUT>(src, n, c, sy, sx) : 0.f, sc[c], sh[c]) : 0.f;
      split_terms<FP16>(v, t[c]);
    }
    const int ph = (y & 1) * 2 + (x & 1);
    const size_t pos = (size_t)(y >> 1) * IX + (x >> 1);
    for (int s = 0; s < S; ++s) {
      union { uint16_t h[8]; uint4 u; } pk;
#pragma unroll
      for (int c = 0; c < 8; ++c) pk.h[c] = t[c][s];
      reinterpret_cast<uint4*>(out)[(((size_t)n * S + s) * 4 + ph) * plane + pos] = pk.u;
    }
  }
}

// split AP planes [n][S][C/8][HW][8] -> (n, C, HW) fp32 (terms summed smallest first)
template <bool FP16>
__global__ void __launch_bounds__(256)
ap_to_nchw_kernel(const uint16_t* __restrict__ ap, float* __restrict__ out, int C, size_t HW, int S) {
  const int c8 = blockIdx.y, n = blockIdx.z;
  for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += (size_t)gridDim.x * blockDim.x) {
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = S - 1; s >= 0; --s) {
      const uint4 q = reinterpret_cast<const uint4*>(ap)[((size_t)(n * S + s) * (C / 8) + c8) * HW + pix];
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint16_t h = (uint16_t)(w[e >> 1] >> (16 * (e & 1)));
        v[e] += FP16 ? __half2float(__ushort_as_half(h)) : __uint_as_float((uint32_t)h << 16);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) out[((size_t)n * C + c8 * 8 + e) * HW + pix] = v[e];
  }
}

std::vector<TcgShape> embedding_shapes(const pds_embedding* e, int H, int W) {
  std::vector<TcgShape> v;
  auto add = [&](int kind, int cin, int cout, int y, int x) {
    TcgShape s; s.kind = kind; s.nd = 2; s.Cin = cin; s.Cout = cout; s.Z = 1; s.Y = y; s.X = x; s.S = e->split;
    v.push_back(s);
  };
  add(TCG_CONV5_S2, 8, e->F, H, W);
  add(TCG_CONV5_S2, e->F, e->F, H / 2, W / 2);
  add(TCG_CONV3_S1, e->F, e->Fs, H / 4, W / 4);
  return v;
}

int prepare(pds_embedding* e, int H, int W, cudaStream_t st) {
  if (e->shape[0] == H && e->shape[1] == W && !e->layers.empty()) return PDS_OK;
  if (!e->layers.empty()) {     // park the current extent: a change of extent swaps, it does not free
    pds_embedding::Saved sv;
    sv.shape[0] = e->shape[0]; sv.shape[1] = e->shape[1];
    sv.layers = std::move(e->layers); sv.blob = e->blob;
    e->saved.push_back(std::move(sv));
    e->layers.clear(); e->blob = nullptr; e->shape[0] = e->shape[1] = 0;
  }
  for (size_t i = 0; i < e->saved.size(); ++i) {
    pds_embedding::Saved& sv = e->saved[i];
    if (sv.shape[0] == H && sv.shape[1] == W) {
      e->layers = std::move(sv.layers); e->blob = sv.blob;
      e->shape[0] = H; e->shape[1] = W;
      e->saved.erase(e->saved.begin() + i);
      return PDS_OK;
    }
  }
  constexpr size_t kMaxSavedShapes = 8;
  if (e->saved.size() > kMaxSavedShapes) {
    cudaFree(e->saved.front().blob);      // implicit device synchronisation
    e->saved.erase(e->saved.begin());
  }
  const std::vector<TcgShape> shapes = embedding_shapes(e, H, W);
  std::vector<TcgLayer> layers(shapes.size());
  size_t bytes = 0;
  for (size_t i = 0; i < shapes.size(); ++i) {
    int rc = tcg_plan(shapes[i], &layers[i].plan);
    if (rc != PDS_OK) return rc;
    layers[i].fp16 = e->fp16;
    layers[i].wscale = e->fp16 ? 256.f : 1.f;
    if (i == 0) layers[i].cin_src = e->Cin;
    bytes += tcg_layer_bytes(layers[i]);
  }
  PDS_CUDA(cudaMalloc(&e->blob, bytes));
  char* cur = e->blob;
  for (size_t i = 0; i < layers.size(); ++i) {
    size_t used = 0;
    const size_t block = i < 2 ? i : (size_t)(2 + 2 * e->n_res);     // parameter block of this layer
    const float* const* pp = &e->raw_params[4 * block];   // weight, bias, gamma, beta
    int rc = tcg_layer_init(layers[i], cur, pp[0], pp[1], st, &used);
    if (rc != PDS_OK) return rc;
    layers[i].gamma = pp[2]; layers[i].beta = pp[3];
    cur += used;
  }
  // the packing kernels ran on `st`; other streams that find the shape already prepared must not
  // race them (one-time cost per shape)
  PDS_CUDA(cudaStreamSynchronize(st));
  e->layers = layers;
  e->shape[0] = H; e->shape[1] = W;
  return PDS_OK;
}

struct Buffers { size_t ap, y1, yq, stats, total; };

Buffers buffers(const pds_embedding* e, int n, int H, int W) {
  Buffers b;
  auto buf = [&](size_t bytes) { return align_up(bytes, 256); };
  const size_t S = e->split, F = e->F;
  // operand planes: the image (8 ch at full size) and the half-size F-channel tensor are the largest
  b.ap = buf(std::max((size_t)n * S * H * W * 16, (size_t)n * S * (H / 2) * (W / 2) * F * 2));
  b.y1 = buf((size_t)n * (H / 2) * (W / 2) * F * 4);
  b.yq = buf((size_t)n * (H / 4) * (W / 4) * F * 4);
  b.stats = buf((size_t)(4 + 2 * e->n_res) * n * 64 * 2 * sizeof(double));   // conv1, conv2, residual convs, shortcut, image
  b.total = 2 * b.ap + b.y1 + 2 * b.yq + b.stats + 1024;
  return b;
}

}  // namespace
}  // namespace pds

extern "C" int pds_embedding_create(pds_embedding** out, const float* const* params, int n_params,
                                    int in_features, int features, int shortcut_features,
                                    int residual_blocks, int precision, void* stream) {
  using namespace pds;
  PDS_CHECK_ARG(out && params, "pds_embedding_create: null pointer");
  PDS_CHECK_ARG(in_features >= 1 && features >= 1 && shortcut_features >= 1 && residual_blocks >= 0,
                "pds_embedding_create: bad sizes");
  PDS_CHECK_ARG(n_params == 4 * (3 + 2 * residual_blocks),
                "pds_embedding_create: expected %d parameter tensors, got %d", 4 * (3 + 2 * residual_blocks), n_params);
  if (precision == PDS_PRECISION_FP32 || precision < 0 || precision > PDS_PRECISION_FP16) {
    set_error("pds_embedding_create: the embedding kernels are tensor-core only (precision != fp32)");
    return PDS_ERR_UNSUPPORTED;
  }
  if (!tcg_available() || in_features > 8 || features != 64 || shortcut_features % 4 || shortcut_features > 16) {
    set_error("pds_embedding_create: needs <= 8 input channels, 64 features, shortcut features in {4, 8, 12, 16}");
    return PDS_ERR_UNSUPPORTED;
  }
  pds_embedding* e = new (std::nothrow) pds_embedding();
  PDS_CHECK_ARG(e, "out of host memory");
  e->Cin = in_features; e->F = features; e->Fs = shortcut_features; e->n_res = residual_blocks;
  e->precision = precision;
  e->split = precision == PDS_PRECISION_BF16X3 ? 3
             : (precision == PDS_PRECISION_BF16X2 || precision == PDS_PRECISION_FP16X2) ? 2 : 1;
  e->fp16 = (precision == PDS_PRECISION_FP16X2 || precision == PDS_PRECISION_FP16) ? 1 : 0;
  // parameter sizes, state_dict order: [w, b, gamma, beta] per block
  std::vector<size_t> sizes;
  auto block = [&](int cin, int cout, int k) {
    sizes.push_back((size_t)cout * cin * k * k); sizes.push_back(cout); sizes.push_back(cout); sizes.push_back(cout);
  };
  block(in_features, features, 5);
  block(features, features, 5);
  for (int r = 0; r < 2 * residual_blocks; ++r) block(features, features, 3);
  block(features, shortcut_features, 3);
  size_t total = 0;
  for (size_t s : sizes) total += align_up(s, 64);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t err = cudaMalloc(&e->raw, total * sizeof(float));
  if (err != cudaSuccess) { delete e; return cuda_fail(err, "cudaMalloc(embedding parameters)"); }
  float* cur = e->raw;
  for (size_t i = 0; i < sizes.size(); ++i) {
    err = cudaMemcpyAsync(cur, params[i], sizes[i] * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (err != cudaSuccess) { cudaFree(e->raw); delete e; return cuda_fail(err, "cudaMemcpyAsync(parameters)"); }
    e->raw_params.push_back(cur);
    cur += align_up(sizes[i], 64);
  }
  // residual-block convolutions: weight images of the matching kernel (independent of the extent)
  e->res.resize(2 * residual_blocks);
  size_t rbytes = 0;
  for (TcLayer& l : e->res) {
    l.Cin = features; l.Cout = features; l.N = 64; l.S = e->split; l.fp16 = e->fp16; l.wscale = e->fp16 ? 256.f : 1.f;
    rbytes += align_up(l.w_elems() * 2, 256) + align_up(l.N * 4, 256);
  }
  if (rbytes) {
    err = cudaMalloc(&e->res_blob, rbytes);
    if (err != cudaSuccess) { cudaFree(e->raw); delete e; return cuda_fail(err, "cudaMalloc(embedding weights)"); }
    char* rc_cur = e->res_blob;
    for (size_t i = 0; i < e->res.size(); ++i) {
      TcLayer& l = e->res[i];
      const float* const* pp = &e->raw_params[4 * (2 + i)];
      l.w = (uint16_t*)rc_cur; rc_cur += align_up(l.w_elems() * 2, 256);
      l.bias = (float*)rc_cur; rc_cur += align_up(l.N * 4, 256);
      l.gamma = pp[2]; l.beta = pp[3];
      int rc = tc_prepare_weights(l, pp[0], pp[1], st);
      if (rc != PDS_OK) { cudaFree(e->raw); cudaFree(e->res_blob); delete e; return rc; }
    }
  }
  *out = e;
  return PDS_OK;
}

extern "C" void pds_embedding_destroy(pds_embedding* e) {
  if (!e) return;
  cudaFree(e->raw);
  cudaFree(e->res_blob);
  cudaFree(e->blob);
  for (auto& sv : e->saved) cudaFree(sv.blob);
  delete e;
}

extern "C" size_t pds_embedding_workspace_bytes(const pds_embedding* e, int n, int H, int W) {
  if (!e || n <= 0 || H <= 0 || W <= 0) return 0;
  return pds::buffers(e, n, H, W).total;
}

namespace pds {
namespace {
int embedding_forward(pds_embedding* e, const ImageSrc& img, float* descriptor, float* shortcut, int n,
                      int n_shortcut, int H, int W, void* workspace, size_t workspace_bytes, void* stream);
}
}  // namespace pds

extern "C" int pds_embedding_forward(pds_embedding* e, const float* images, float* descriptor,
                                     float* shortcut, int n, int n_shortcut, int H, int W,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  PDS_CHECK_ARG(e && images && descriptor, "pds_embedding_forward: null pointer");
  pds::ImageSrc img = {images, nullptr, n, PDS_IMAGE_F32_NCHW, e->Cin, H, W, 0, 0};
  return pds::embedding_forward(e, img, descriptor, shortcut, n, n_shortcut, H, W, workspace, workspace_bytes, stream);
}

extern "C" int pds_embedding_forward_images(pds_embedding* e, const void* images_a, int n_a,
                                            const void* images_b, int n_b, int layout, int h, int w,
                                            int pad_top, int pad_left, float* descriptor, float* shortcut,
                                            int n_shortcut, void* workspace, size_t workspace_bytes,
                                            void* stream) {
  PDS_CHECK_ARG(e && descriptor && n_a >= 0 && n_b >= 0 && (images_a || n_a == 0) && (images_b || n_b == 0),
                "pds_embedding_forward_images: null pointer");
  PDS_CHECK_ARG(layout == PDS_IMAGE_F32_NCHW || layout == PDS_IMAGE_U8_NCHW || layout == PDS_IMAGE_U8_NHWC,
                "pds_embedding_forward_images: unknown image layout");
  PDS_CHECK_ARG(h >= 1 && w >= 1 && pad_top >= 0 && pad_left >= 0,
                "pds_embedding_forward_images: bad image extent or padding");
  pds::ImageSrc img = {images_a, images_b, n_a, layout, e->Cin, h, w, pad_top, pad_left};
  return pds::embedding_forward(e, img, descriptor, shortcut, n_a + n_b, n_shortcut, h + pad_top, w + pad_left,
                                workspace, workspace_bytes, stream);
}

namespace pds {
namespace {
int embedding_forward(pds_embedding* e, const ImageSrc& img, float* descriptor, float* shortcut, int n,
                      int n_shortcut, int H, int W, void* workspace, size_t workspace_bytes, void* stream) {
  PDS_CHECK_ARG(n >= 0 && n_shortcut >= 0 && n_shortcut <= n && (shortcut || n_shortcut == 0),
                "pds_embedding_forward: bad sample counts");
  PDS_CHECK_ARG(H >= 4 && W >= 4 && H % 4 == 0 && W % 4 == 0,
                "pds_embedding_forward: (padded) height and width must be multiples of 4");
  if (n == 0) return PDS_OK;
  if (!workspace || workspace_bytes < pds_embedding_workspace_bytes(e, n, H, W) || ((uintptr_t)workspace & 255)) {
    set_error("pds_embedding_forward: workspace too small or not 256-byte aligned");
    return PDS_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int rc = prepare(e, H, W, st);
  if (rc != PDS_OK) return rc;
  const Buffers bs = buffers(e, n, H, W);
  Workspace ws(workspace, workspace_bytes);
  uint16_t* ap[2] = {(uint16_t*)ws.take<char>(bs.ap), (uint16_t*)ws.take<char>(bs.ap)};
  float* y1 = (float*)ws.take<char>(bs.y1);
  float* y2 = (float*)ws.take<char>(bs.yq);
  float* t = (float*)ws.take<char>(bs.yq);
  double* stats = (double*)ws.take<char>(bs.stats);
  if (ws.overflow) { set_error("pds_embedding_forward: workspace overflow"); return PDS_ERR_WORKSPACE; }
  PDS_CUDA(cudaMemsetAsync(stats, 0, bs.stats, st));
  const size_t stat_stride = (size_t)n * 64 * 2;
  auto st_of = [&](int i) { return stats + stat_stride * i; };
  const std::vector<TcgLayer>& L = e->layers;
  const int S = e->split, fp16 = e->fp16, F = e->F, Hh = H / 2, Wh = W / 2, Hq = H / 4, Wq = W / 4;
  auto src = [&](int layer, const float* y) {
    TcgNormSrc s; s.y = y; s.stats = st_of(layer); s.gamma = L[layer].gamma; s.beta = L[layer].beta;
    return s;
  };
  const int n_layers = (int)L.size();
  double* img_stats = st_of(3 + 2 * e->n_res);   // slot after the layers'

  // pad + InstanceNorm2d of the image -> phase-separated planes (size_adapter.py:42, embedding.py:32)
  {
    const size_t HW = (size_t)H * W, hw = (size_t)img.h * img.w;
    const double in_bytes = (double)n * e->Cin * hw * (img.layout == PDS_IMAGE_F32_NCHW ? 4.0 : 1.0);
    dim3 sgrid((unsigned)std::min<size_t>((hw / 4 + 255) / 256, img.layout == PDS_IMAGE_F32_NCHW ? 64 : 128),
               (unsigned)(n * e->Cin));
    dim3 grid((unsigned)std::min<size_t>((HW + 255) / 256, (size_t)num_sms() * 8), (unsigned)n);
    {
      PDS_KERNEL("image_stats", st);
      PDS_KERNEL_WORK(0, in_bytes);
      switch (img.layout) {
        case PDS_IMAGE_F32_NCHW: image_stats_kernel<PDS_IMAGE_F32_NCHW><<<sgrid, 256, 0, st>>>(img, img_stats); break;
        case PDS_IMAGE_U8_NCHW: image_stats_kernel<PDS_IMAGE_U8_NCHW><<<sgrid, 256, 0, st>>>(img, img_stats); break;
        default: image_stats_kernel<PDS_IMAGE_U8_NHWC><<<sgrid, 256, 0, st>>>(img, img_stats); break;
      }
      PDS_LAUNCH_CHECK("image_stats_kernel");
    }
    PDS_KERNEL("image_norm_to_ap", st);
    PDS_KERNEL_WORK(0, in_bytes + (double)n * HW * 16.0 * S);
#define PDS_IMG_CASE(FF, LL) \
    if ((fp16 != 0) == FF && img.layout == LL) image_norm_to_ap_kernel<FF, LL><<<grid, 256, 0, st>>>(img, img_stats, ap[0], H, W, S);
    PDS_IMG_CASE(true, PDS_IMAGE_F32_NCHW) PDS_IMG_CASE(true, PDS_IMAGE_U8_NCHW) PDS_IMG_CASE(true, PDS_IMAGE_U8_NHWC)
    PDS_IMG_CASE(false, PDS_IMAGE_F32_NCHW) PDS_IMG_CASE(false, PDS_IMAGE_U8_NCHW) PDS_IMG_CASE(false, PDS_IMAGE_U8_NHWC)
#undef PDS_IMG_CASE
    PDS_LAUNCH_CHECK("image_norm_to_ap_kernel");
  }
  // two stride-2 5x5 blocks
  if ((rc = tcg_conv_forward(L[0], n, ap[0], y1, st_of(0), 1, st)) != PDS_OK) return rc;
  if ((rc = tcg_norm_to_ap(src(0, y1), nullptr, nullptr, ap[1], n, F, 1, Hh, Wh, S, fp16, 4, st)) != PDS_OK) return rc;
  if ((rc = tcg_conv_forward(L[1], n, ap[1], y2, st_of(1), 1, st)) != PDS_OK) return rc;
  // residual stream x = IN(y2) as operand planes; every block: x <- IN(conv(IN(conv(x)))) + x, on
  // the matching operation's kernels with the images as slices (network_blocks.py:134-144)
  if ((rc = tcg_norm_to_ap(src(1, y2), nullptr, nullptr, ap[0], n, F, 1, Hq, Wq, S, fp16, 1, st)) != PDS_OK) return rc;
  TcConvArgs a = {};
  a.H = Hq; a.W = Wq; a.n_slices = n; a.n0 = 0; a.n_div = 1; a.in_slices = n; a.in_C = F;
  a.epilogue = TC_EPI_ACT; a.out_f32 = t;
  for (int r = 0; r < e->n_res; ++r) {
    const TcLayer& c1 = e->res[2 * r];
    const TcLayer& c2 = e->res[2 * r + 1];
    double* s1 = st_of(2 + 2 * r);
    double* s2 = st_of(3 + 2 * r);
    a.layer = &c1; a.in = ap[0]; a.stats = s1;
    if ((rc = tc_conv3x3(a, st)) != PDS_OK) return rc;
    if ((rc = tc_norm_split(t, s1, c1.gamma, c1.beta, nullptr, ap[1], n, F, Hq, Wq, S, fp16, st)) != PDS_OK) return rc;
    a.layer = &c2; a.in = ap[1]; a.stats = s2;
    if ((rc = tc_conv3x3(a, st)) != PDS_OK) return rc;
    if ((rc = tc_norm_split(t, s2, c2.gamma, c2.beta, ap[0], ap[0], n, F, Hq, Wq, S, fp16, st)) != PDS_OK) return rc;
  }
  {
    const size_t HWq = (size_t)Hq * Wq;
    PDS_KERNEL("ap_to_nchw", st);
    PDS_KERNEL_WORK(0, (double)n * F * HWq * (4.0 + 2.0 * S));
    dim3 grid((unsigned)std::min<size_t>((HWq + 255) / 256, 64), (unsigned)(F / 8), (unsigned)n);
    if (fp16) ap_to_nchw_kernel<true><<<grid, 256, 0, st>>>(ap[0], descriptor, F, HWq, S);
    else ap_to_nchw_kernel<false><<<grid, 256, 0, st>>>(ap[0], descriptor, F, HWq, S);
    PDS_LAUNCH_CHECK("ap_to_nchw_kernel");
  }
  if (n_shortcut > 0) {
    // _shortcut block on the descriptor planes (embedding.py:43-44,65)
    const int ls = n_layers - 1, ss = 2 + 2 * e->n_res;    // layer index, statistics slot
    if ((rc = tcg_conv_forward(L[ls], n_shortcut, ap[0], t, st_of(ss), 1, st)) != PDS_OK) return rc;
    if ((rc = instance_norm_apply(t, st_of(ss), L[ls].gamma, L[ls].beta, nullptr, nullptr, t, nullptr,
                                  n_shortcut, (size_t)Hq * Wq, (size_t)Hq * Wq, e->Fs, st)) != PDS_OK) return rc;
    if ((rc = nhwc_to_nchw(t, shortcut, n_shortcut, e->Fs, (size_t)Hq * Wq, st)) != PDS_OK) return rc;
  }
  return PDS_OK;
}
}  // namespace
}  // namespace pds
