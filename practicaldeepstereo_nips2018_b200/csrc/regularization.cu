// a3 -- Regularization.forward (reference regularization.py:94-126): the 3-D
// hourglass over the (B, 8, D, H, W) signature volume, plus the two blocks the
// reference tests individually (ContractionBlock3d :28-31, ExpansionBlock3d
// :54-57).  Activations are channels-last (NDHWC); every Conv -> LeakyReLU ->
// InstanceNorm block is one convolution launch whose epilogue applies bias and
// LeakyReLU and accumulates the InstanceNorm sums, followed by one
// normalise(+add) pass that also produces the sums the reference forms right
// after (shortcut + output, up-sampled + skip).
#include <stdlib.h>

#include <new>
#include <vector>

#include "conv_layers.cuh"
#include "conv_tc.cuh"
#include "conv_tcg.cuh"

struct pds_regularization {
  int F, precision;
  std::vector<pds::ConvLayer> layers;
  float* blob = nullptr;
  // tensor-core path (precision != fp32): raw parameters in PyTorch layout (the programs and the
  // weight images depend on the volume extent, so they are (re)built by the first forward of a shape)
  int split = 0, fp16 = 0;
  float* raw = nullptr;
  std::vector<const float*> raw_params;
  std::vector<pds::TcgLayer> tcg;       // all layers but _upsample_to_fullsize (the CURRENT extent)
  char* tcg_blob = nullptr;
  int tcg_shape[3] = {0, 0, 0};
  // plans and weight images of other extents seen by this handle: a change of extent swaps, it does
  // not free (CUDA graphs captured for an earlier extent keep pointing at valid memory); at most
  // kMaxSavedShapes are kept, the oldest is freed beyond that
  struct Saved { int shape[3]; std::vector<pds::TcgLayer> tcg; char* blob; };
  std::vector<Saved> saved;
  // host copies for the fused tail kernel (F == 8): last layer's weight (4,1,3,4,4) and bias,
  // InstanceNorm affine of _upsample_to_halfsize
  bool fused_tail = false;
  float tail_w[192], tail_bias = 0.f, tail_gamma[4], tail_beta[4];
};

namespace pds {
namespace {

const DimSpec kC3 = {DM_CONV3, 1}, kC3s2 = {DM_CONV3, 2}, kT4 = {DM_TCONV4, 1}, kT3 = {DM_TCONV3, 1};

ConvLayer block3(int cin, int cout, int stride) {
  const DimSpec d = stride == 2 ? kC3s2 : kC3;
  return make_layer(cin, cout, d, d, d, false, true);
}
ConvLayer tblock4(int cin, int cout) { return make_layer(cin, cout, kT4, kT4, kT4, true, true); }

// Layer order == parameter order of Regularization.state_dict().
std::vector<ConvLayer> hourglass_layers(int F) {
  std::vector<ConvLayer> v;
  v.push_back(block3(F, F, 1));                                   // _smoothing
  for (int s = 1; s <= 8; s *= 2) {                               // _contraction_blocks
    v.push_back(block3(F * s, 2 * F * s, 2));
    v.push_back(block3(2 * F * s, 2 * F * s, 1));
  }
  for (int s = 16; s >= 2; s /= 2) {                              // _expansion_blocks
    v.push_back(tblock4(F * s, F * s / 2));
    v.push_back(block3(F * s / 2, F * s / 2, 1));
  }
  v.push_back(tblock4(F, F / 2));                                 // _upsample_to_halfsize
  v.push_back(make_layer(F / 2, 1, kT3, kT4, kT4, true, false));  // _upsample_to_fullsize
  return v;
}

// Copies / re-lays-out parameters (state_dict order) into `blob`.
int load_layers(std::vector<ConvLayer>& layers, const float* const* params, float* blob,
                cudaStream_t st) {
  float* cur = blob;
  int pi = 0;
  for (auto& l : layers) {
    l.w = cur; cur += align_up(l.weight_elems(), 64);
    int rc = relayout_weights(l, params[pi++], l.w, st);
    if (rc != PDS_OK) return rc;
    const int nvec = l.lrelu ? 3 : 1;
    const float** dst[3] = {&l.bias, &l.gamma, &l.beta};
    for (int v = 0; v < nvec; ++v) {
      PDS_CUDA(cudaMemcpyAsync(cur, params[pi++], l.Cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
      *dst[v] = cur; cur += align_up(l.Cout, 64);
    }
  }
  return PDS_OK;
}

size_t blob_elems(const std::vector<ConvLayer>& layers) {
  size_t total = 0;
  for (auto& l : layers) total += align_up(l.weight_elems(), 64) + align_up(l.Cout, 64) * 3;
  return total;
}

int n_param_tensors(const std::vector<ConvLayer>& layers) {
  int n = 0;
  for (auto& l : layers) n += l.lrelu ? 4 : 2;
  return n;
}

// conv + LeakyReLU + stats, then InstanceNorm apply; out = IN(y), out2 = IN(y)+add+add_bcast.
int conv_block(const ConvLayer& l, const ConvGeom& g, const float* in, float* y, double* stats,
               const float* add, const float* add_bcast, float* out, float* out2, size_t bcast_hw,
               cudaStream_t st) {
  const size_t S = (size_t)l.dim[0].out_size(g.D) * l.dim[1].out_size(g.H) * l.dim[2].out_size(g.W);
  PDS_CUDA(cudaMemsetAsync(stats, 0, (size_t)g.N * l.Cout * 2 * sizeof(double), st));
  // direct FFMA kernels for the 8 / 16-channel 3x3x3 layers (4.64 -> 4.06 ms for the fp32 hourglass at C2)
  static const bool direct = !(getenv("PDS_B200_DIRECT3D") && atoi(getenv("PDS_B200_DIRECT3D")) == 0);
  bool handled = false;
  int rc = PDS_OK;
  if (direct) rc = conv_forward_direct(l, g, in, y, stats, st, &handled);
  if (rc == PDS_OK && !handled) rc = conv_forward_simt(l, g, in, nullptr, 0, y, stats, st);
  if (rc != PDS_OK) return rc;
  return instance_norm_apply(y, stats, l.gamma, l.beta, add, add_bcast, out, out2, g.N, S,
                             bcast_hw ? bcast_hw : S, l.Cout, st);
}

// Device copy of the parameters in PyTorch layout (state_dict order) for the tensor-core path.
int copy_raw_params(pds_regularization* reg, const float* const* params, cudaStream_t st) {
  size_t total = 0;
  std::vector<size_t> sizes;
  for (auto& l : reg->layers) {
    sizes.push_back(l.weight_elems_pytorch()); sizes.push_back(l.Cout);
    if (l.lrelu) { sizes.push_back(l.Cout); sizes.push_back(l.Cout); }
  }
  for (size_t n : sizes) total += align_up(n, 64);
  PDS_CUDA(cudaMalloc(&reg->raw, total * sizeof(float)));
  float* cur = reg->raw;
  for (size_t i = 0; i < sizes.size(); ++i) {
    PDS_CUDA(cudaMemcpyAsync(cur, params[i], sizes[i] * sizeof(float), cudaMemcpyDeviceToDevice, st));
    reg->raw_params.push_back(cur);
    cur += align_up(sizes[i], 64);
  }
  return PDS_OK;
}

// Shapes of the 18 tensor-core layers for a (D, H, W) signature volume, in layer order.
std::vector<TcgShape> hourglass_shapes(int F, int S, int D, int H, int W) {
  std::vector<TcgShape> v;
  auto add = [&](int kind, int cin, int cout, int z, int y, int x) {
    TcgShape s; s.kind = kind; s.nd = 3; s.Cin = cin; s.Cout = cout; s.Z = z; s.Y = y; s.X = x; s.S = S;
    v.push_back(s);
  };
  int c = F, z = D, y = H, x = W;
  // the two full-extent F -> F layers (98 % of the hourglass' voxels) put four voxels of a row on the
  // N axis (conv_tcg.cuh, TCG_CONV3_S1X4) when F == 8 and the width allows it; PDS_B200_TCG_X4=0: one
  // voxel per GEMM row as everywhere else
  static const bool x4_on = !(getenv("PDS_B200_TCG_X4") && atoi(getenv("PDS_B200_TCG_X4")) == 0);
  const int full_kind = (x4_on && F == 8 && W % 4 == 0) ? TCG_CONV3_S1X4 : TCG_CONV3_S1;
  add(full_kind, c, c, z, y, x);
  for (int k = 0; k < 4; ++k) {
    add(TCG_CONV3_S2, c, 2 * c, z, y, x);
    c *= 2; z /= 2; y /= 2; x /= 2;
    add(TCG_CONV3_S1, c, c, z, y, x);
  }
  // the 4-channel transposed layer runs with its parity classes merged along N (measured: 250 ->
  // 161 us at C2; with 8 output channels the merged N = 64 tile is no faster than 8 class passes)
  static const bool merge = !(getenv("PDS_B200_TCONV_MERGE") && atoi(getenv("PDS_B200_TCONV_MERGE")) == 0);
  auto tkind = [&](int cout) { return (merge && cout <= 4) ? TCG_TCONV4_S2M : TCG_TCONV4_S2; };
  for (int k = 0; k < 4; ++k) {
    add(tkind(c / 2), c, c / 2, z, y, x);
    c /= 2; z *= 2; y *= 2; x *= 2;
    add(k == 3 ? full_kind : TCG_CONV3_S1, c, c, z, y, x);
  }
  add(tkind(c / 2), c, c / 2, z, y, x);
  return v;
}

// (Re)plans the layers and rebuilds their weight images for a new volume extent.
int prepare_tcg(pds_regularization* reg, int D, int H, int W, cudaStream_t st) {
  if (reg->tcg_shape[0] == D && reg->tcg_shape[1] == H && reg->tcg_shape[2] == W && !reg->tcg.empty())
    return PDS_OK;
  if (!reg->tcg.empty()) {     // park the current extent
    pds_regularization::Saved sv;
    for (int i = 0; i < 3; ++i) sv.shape[i] = reg->tcg_shape[i];
    sv.tcg = std::move(reg->tcg); sv.blob = reg->tcg_blob;
    reg->saved.push_back(std::move(sv));
    reg->tcg.clear(); reg->tcg_blob = nullptr; reg->tcg_shape[0] = reg->tcg_shape[1] = reg->tcg_shape[2] = 0;
  }
  for (size_t i = 0; i < reg->saved.size(); ++i) {
    pds_regularization::Saved& sv = reg->saved[i];
    if (sv.shape[0] == D && sv.shape[1] == H && sv.shape[2] == W) {
      reg->tcg = std::move(sv.tcg); reg->tcg_blob = sv.blob;
      reg->tcg_shape[0] = D; reg->tcg_shape[1] = H; reg->tcg_shape[2] = W;
      reg->saved.erase(reg->saved.begin() + i);
      return PDS_OK;
    }
  }
  constexpr size_t kMaxSavedShapes = 8;
  if (reg->saved.size() > kMaxSavedShapes) {
    cudaFree(reg->saved.front().blob);      // implicit device synchronisation
    reg->saved.erase(reg->saved.begin());
  }
  const std::vector<TcgShape> shapes = hourglass_shapes(reg->F, reg->split, D, H, W);
  std::vector<TcgLayer> layers(shapes.size());
  size_t bytes = 0;
  for (size_t i = 0; i < shapes.size(); ++i) {
    int rc = tcg_plan(shapes[i], &layers[i].plan);
    if (rc != PDS_OK) return rc;
    layers[i].transposed = shapes[i].kind == TCG_TCONV4_S2 || shapes[i].kind == TCG_TCONV4_S2M;
    layers[i].fp16 = reg->fp16;
    layers[i].wscale = reg->fp16 ? 256.f : 1.f;
    bytes += tcg_layer_bytes(layers[i]);
  }
  PDS_CUDA(cudaMalloc(&reg->tcg_blob, bytes));
  char* cur = reg->tcg_blob;
  for (size_t i = 0; i < layers.size(); ++i) {
    size_t used = 0;
    const float* const* pp = &reg->raw_params[4 * i];   // weight, bias, gamma, beta
    int rc = tcg_layer_init(layers[i], cur, pp[0], pp[1], st, &used);
    if (rc != PDS_OK) return rc;
    layers[i].gamma = pp[2]; layers[i].beta = pp[3];
    cur += used;
  }
  // the packing kernels ran on `st`; other streams that find the shape already prepared must not
  // race them (one-time cost per shape)
  PDS_CUDA(cudaStreamSynchronize(st));
  reg->tcg = layers;
  reg->tcg_shape[0] = D; reg->tcg_shape[1] = H; reg->tcg_shape[2] = W;
  return PDS_OK;
}

// SubpixelMap + SizeAdapter.unpad fused into the last layer (disparity == null: plain cost volume)
struct TailFusion {
  float* disparity = nullptr;
  int64_t* argmax = nullptr;
  int R = 0, step = 1, crop_top = 0, crop_left = 0;
  float* state = nullptr;      // workspace for the segment states (taken by the forward functions)
};

struct TcgBuffers {
  size_t sc_cl, ap, y_l0, y_d[4], y_s[4], y_u[4], y_e[4], y_half, stats, tail_state, splitk, total;
};

// Scratch of the split-K layers (conv_tcg.cu): the largest need over the layers of this extent.  Uses
// the plans of the handle when it has seen the extent, otherwise plans it on the spot (host only).
size_t splitk_scratch_bytes(const pds_regularization* reg, int B, int D, int H, int W) {
  const std::vector<TcgLayer>* layers = nullptr;
  if (reg->tcg_shape[0] == D && reg->tcg_shape[1] == H && reg->tcg_shape[2] == W && !reg->tcg.empty()) layers = &reg->tcg;
  for (const auto& sv : reg->saved)
    if (!layers && sv.shape[0] == D && sv.shape[1] == H && sv.shape[2] == W) layers = &sv.tcg;
  std::vector<TcgLayer> planned;
  if (!layers) {
    const std::vector<TcgShape> shapes = hourglass_shapes(reg->F, reg->split, D, H, W);
    planned.resize(shapes.size());
    for (size_t i = 0; i < shapes.size(); ++i)
      if (tcg_plan(shapes[i], &planned[i].plan) != PDS_OK) return 0;
    layers = &planned;
  }
  size_t bytes = 0;
  for (const TcgLayer& l : *layers) bytes = std::max(bytes, tcg_splitk_bytes(l, B));
  return bytes;
}

TcgBuffers tcg_buffers(const pds_regularization* reg, int B, int D, int H, int W) {
  TcgBuffers b;
  const size_t vox = (size_t)D * H * W, F = reg->F, S = reg->split;
  auto buf = [&](size_t bytes) { return align_up(bytes, 256); };
  b.sc_cl = buf((size_t)B * H * W * F * 4);
  b.ap = buf((size_t)B * S * vox * F * 2);          // the widest AP tensor: F channels at full extent
  b.y_l0 = buf(B * vox * F * 4);
  size_t c = F, v = vox;
  for (int k = 0; k < 4; ++k) { c *= 2; v /= 8; b.y_d[k] = b.y_s[k] = buf(B * v * c * 4); }
  for (int k = 0; k < 4; ++k) { c /= 2; v *= 8; b.y_u[k] = b.y_e[k] = buf(B * v * c * 4); }
  b.y_half = buf(B * vox * 8 * (F / 2) * 4);
  b.stats = buf((size_t)18 * B * 16 * F * 2 * sizeof(double));
  b.tail_state = hourglass_tail_state_bytes(B, 2 * D, 2 * H, 2 * W);
  b.splitk = splitk_scratch_bytes(reg, B, D, H, W);
  b.total = b.sc_cl + 2 * b.ap + b.y_l0 + b.y_half + b.stats + b.tail_state + b.splitk + 1024;
  for (int k = 0; k < 4; ++k) b.total += b.y_d[k] + b.y_s[k] + b.y_u[k] + b.y_e[k];
  return b;
}

// Regularization.forward on the tcgen05 engine.  Every Conv -> LeakyReLU -> InstanceNorm block is one
// convolution launch (fp32 channels-last output + sums) and one normalisation pass that writes the
// NEXT layer's operand planes, fused with the additions the reference performs in between
// (regularization.py:117-123); skip tensors are never materialised: the pass that needs one
// re-normalises the stored convolution output.
int tcg_forward(pds_regularization* reg, const float* signatures, const float* shortcut, float* cost, int B,
                int D, int H, int W, void* workspace, size_t workspace_bytes, cudaStream_t st,
                const TailFusion& tf) {
  int rc = prepare_tcg(reg, D, H, W, st);
  if (rc != PDS_OK) return rc;
  const int F = reg->F, S = reg->split, fp16 = reg->fp16;
  const TcgBuffers bs = tcg_buffers(reg, B, D, H, W);
  Workspace ws(workspace, workspace_bytes);
  float* sc_cl = (float*)ws.take<char>(bs.sc_cl);
  uint16_t* ap[2] = {(uint16_t*)ws.take<char>(bs.ap), (uint16_t*)ws.take<char>(bs.ap)};
  float* y_l0 = (float*)ws.take<char>(bs.y_l0);
  float *y_d[4], *y_s[4], *y_u[4], *y_e[4];
  for (int k = 0; k < 4; ++k) { y_d[k] = (float*)ws.take<char>(bs.y_d[k]); y_s[k] = (float*)ws.take<char>(bs.y_s[k]); }
  for (int k = 0; k < 4; ++k) { y_u[k] = (float*)ws.take<char>(bs.y_u[k]); y_e[k] = (float*)ws.take<char>(bs.y_e[k]); }
  float* y_half = (float*)ws.take<char>(bs.y_half);
  double* stats = (double*)ws.take<char>(bs.stats);
  float* tail_state = (float*)ws.take<char>(bs.tail_state);
  float* splitk = (float*)ws.take<char>(bs.splitk);
  if (ws.overflow) { set_error("pds_regularization_forward: workspace overflow"); return PDS_ERR_WORKSPACE; }
  PDS_CUDA(cudaMemsetAsync(stats, 0, bs.stats, st));
  const size_t stat_stride = (size_t)B * 16 * F * 2;
  auto st_of = [&](int layer) { return stats + stat_stride * layer; };
  const std::vector<TcgLayer>& L = reg->tcg;
  auto src = [&](int layer, const float* y) {
    TcgNormSrc s; s.y = y; s.stats = st_of(layer); s.gamma = L[layer].gamma; s.beta = L[layer].beta;
    return s;
  };

  if ((rc = nchw_to_nhwc(shortcut, sc_cl, B, F, (size_t)H * W, st)) != PDS_OK) return rc;
  if ((rc = tc_pack_nchw(signatures, ap[0], B, F, 1, (int)((size_t)D * H * W), S, fp16, st,
                         L[0].plan.phase_arg() == TCG_PHASES_X4 ? W : 0)) != PDS_OK) return rc;
  // layer 0: output = smoothing(signatures); level-0 input = shortcut (broadcast over D) + output
  if ((rc = tcg_conv_forward(L[0], B, ap[0], y_l0, st_of(0), 1, st, splitk, bs.splitk)) != PDS_OK) return rc;
  if ((rc = tcg_norm_to_ap(src(0, y_l0), nullptr, sc_cl, ap[1], B, F, D, H, W, S, fp16, 8, st)) != PDS_OK) return rc;
  int c = F, z = D, y = H, x = W, li = 1;
  for (int k = 0; k < 4; ++k) {
    // ContractionBlock3d (regularization.py:28-31): down = block_s2(in); smooth = block(down)
    const int ld = li++, lsm = li++;
    if ((rc = tcg_conv_forward(L[ld], B, ap[1], y_d[k], st_of(ld), 1, st, splitk, bs.splitk)) != PDS_OK) return rc;
    c *= 2; z /= 2; y /= 2; x /= 2;
    if ((rc = tcg_norm_to_ap(src(ld, y_d[k]), nullptr, nullptr, ap[0], B, c, z, y, x, S, fp16, 1, st)) != PDS_OK) return rc;
    if ((rc = tcg_conv_forward(L[lsm], B, ap[0], y_s[k], st_of(lsm), 1, st, splitk, bs.splitk)) != PDS_OK) return rc;
    if (k < 3) {   // next level input = down_k + smooth_k, phase-separated for the stride-2 layer
      const TcgNormSrc d = src(ld, y_d[k]);
      if ((rc = tcg_norm_to_ap(src(lsm, y_s[k]), &d, nullptr, ap[1], B, c, z, y, x, S, fp16, 8, st)) != PDS_OK) return rc;
    } else {
      if ((rc = tcg_norm_to_ap(src(lsm, y_s[k]), nullptr, nullptr, ap[1], B, c, z, y, x, S, fp16, 1, st)) != PDS_OK) return rc;
    }
  }
  for (int k = 0; k < 4; ++k) {
    // ExpansionBlock3d (regularization.py:54-57): smoothing(up(out) + skip); skip = the smoothing
    // output pushed before contraction 3 - k (layer 2 * (3 - k) of the contraction part, 0 for k = 3)
    const int lu = li++, lsm = li++;
    if ((rc = tcg_conv_forward(L[lu], B, ap[1], y_u[k], st_of(lu), 1, st, splitk, bs.splitk)) != PDS_OK) return rc;
    c /= 2; z *= 2; y *= 2; x *= 2;
    const int skip_layer = k < 3 ? 2 * (3 - k) : 0;
    const TcgNormSrc skip = src(skip_layer, k < 3 ? y_s[2 - k] : y_l0);
    if ((rc = tcg_norm_to_ap(src(lu, y_u[k]), &skip, nullptr, ap[0], B, c, z, y, x, S, fp16, L[lsm].plan.phase_arg(), st)) != PDS_OK) return rc;
    if ((rc = tcg_conv_forward(L[lsm], B, ap[0], y_e[k], st_of(lsm), 1, st, splitk, bs.splitk)) != PDS_OK) return rc;
    if ((rc = tcg_norm_to_ap(src(lsm, y_e[k]), nullptr, nullptr, ap[1], B, c, z, y, x, S, fp16, 1, st)) != PDS_OK) return rc;
  }
  // _upsample_to_halfsize (its InstanceNorm is applied by the tail kernel) + _upsample_to_fullsize
  if ((rc = tcg_conv_forward(L[li], B, ap[1], y_half, st_of(li), 1, st, splitk, bs.splitk)) != PDS_OK) return rc;
  return hourglass_tail_forward(y_half, cost, st_of(li), reg->tail_gamma, reg->tail_beta, reg->tail_w,
                                reg->tail_bias, B, 2 * D, 2 * H, 2 * W, st, tf.disparity, tf.argmax, tf.R,
                                tf.step, tf.crop_top, tf.crop_left, tail_state);
}

}  // namespace
}  // namespace pds

extern "C" int pds_regularization_create(pds_regularization** out, const float* const* params,
                                         int n_params, int F, int precision, void* stream) {
  using namespace pds;
  PDS_CHECK_ARG(out && params, "pds_regularization_create: null pointer");
  PDS_CHECK_ARG(F >= 2 && F % 2 == 0, "pds_regularization_create: number_of_features must be even");
  PDS_CHECK_ARG(precision >= PDS_PRECISION_FP32 && precision <= PDS_PRECISION_FP16,
                "pds_regularization_create: bad precision");
  pds_regularization* reg = new (std::nothrow) pds_regularization();
  PDS_CHECK_ARG(reg, "out of host memory");
  reg->F = F; reg->precision = precision;
  reg->layers = hourglass_layers(F);
  if (n_params != n_param_tensors(reg->layers)) {
    set_error("pds_regularization_create: expected %d parameter tensors, got %d",
              n_param_tensors(reg->layers), n_params);
    delete reg;
    return PDS_ERR_INVALID_ARGUMENT;
  }
  cudaError_t e = cudaMalloc(&reg->blob, blob_elems(reg->layers) * sizeof(float));
  if (e != cudaSuccess) { delete reg; return cuda_fail(e, "cudaMalloc(regularization weights)"); }
  int rc = load_layers(reg->layers, params, reg->blob, (cudaStream_t)stream);
  if (rc == PDS_OK && F == 8) {
    // parameter order (state_dict): ..., _upsample_to_halfsize.{0.weight, 0.bias, 2.weight, 2.bias},
    // _upsample_to_fullsize.{weight, bias}
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e1 = cudaMemcpyAsync(reg->tail_gamma, params[n_params - 4], 4 * sizeof(float), cudaMemcpyDeviceToHost, st);
    cudaError_t e2 = cudaMemcpyAsync(reg->tail_beta, params[n_params - 3], 4 * sizeof(float), cudaMemcpyDeviceToHost, st);
    cudaError_t e3 = cudaMemcpyAsync(reg->tail_w, params[n_params - 2], 192 * sizeof(float), cudaMemcpyDeviceToHost, st);
    cudaError_t e4 = cudaMemcpyAsync(&reg->tail_bias, params[n_params - 1], sizeof(float), cudaMemcpyDeviceToHost, st);
    cudaError_t e5 = cudaStreamSynchronize(st);
    for (cudaError_t e : {e1, e2, e3, e4, e5})
      if (e != cudaSuccess && rc == PDS_OK) rc = cuda_fail(e, "copy of the hourglass tail parameters");
    reg->fused_tail = rc == PDS_OK;
  }
  if (rc == PDS_OK && precision != PDS_PRECISION_FP32 && F == 8 && tcg_available()) {
    reg->split = precision == PDS_PRECISION_BF16X3 ? 3
                 : (precision == PDS_PRECISION_BF16X2 || precision == PDS_PRECISION_FP16X2) ? 2 : 1;
    reg->fp16 = (precision == PDS_PRECISION_FP16X2 || precision == PDS_PRECISION_FP16) ? 1 : 0;
    rc = copy_raw_params(reg, params, (cudaStream_t)stream);
  }
  if (rc != PDS_OK) { cudaFree(reg->blob); cudaFree(reg->raw); delete reg; return rc; }
  *out = reg;
  return PDS_OK;
}

extern "C" void pds_regularization_destroy(pds_regularization* reg) {
  if (!reg) return;
  cudaFree(reg->blob);
  cudaFree(reg->raw);
  cudaFree(reg->tcg_blob);
  for (auto& sv : reg->saved) cudaFree(sv.blob);
  delete reg;
}

extern "C" size_t pds_regularization_workspace_bytes(const pds_regularization* reg, int B, int D,
                                                     int H, int W) {
  using namespace pds;
  if (!reg || B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
  const size_t vox = (size_t)D * H * W, F = reg->F;
  if (reg->raw) return tcg_buffers(reg, B, D, H, W).total;
  auto buf = [&](size_t elems) { return align_up(elems * 4, 256); };
  size_t bytes = buf(B * vox * F) + buf((size_t)B * H * W * F);      // sig_cl, shortcut_cl
  bytes += 2 * buf(B * vox * F);                                     // out0 (skip), sum0
  size_t c = F, v = vox;
  for (int k = 0; k < 4; ++k) {                                      // down, smooth(skip), sum
    c *= 2; v /= 8;
    bytes += 3 * buf(B * v * c);
  }
  for (int k = 0; k < 4; ++k) {                                      // up(+skip), smooth
    c /= 2; v *= 8;
    bytes += 2 * buf(B * v * c);
  }
  bytes += buf(B * vox * 8 * (F / 2));                               // half-size volume
  bytes += align_up((size_t)B * 16 * F * 2 * sizeof(double), 256);   // stats (largest layer)
  bytes += hourglass_tail_state_bytes(B, 2 * D, 2 * H, 2 * W);       // fused tail + estimator
  return bytes + 1024;
}

namespace pds {
namespace {
int regularization_forward(pds_regularization* reg, const float* signatures, const float* shortcut,
                           float* cost, int B, int D, int H, int W, void* workspace,
                           size_t workspace_bytes, void* stream, const TailFusion& tf) {
  PDS_CHECK_ARG(reg && signatures && shortcut && (cost || tf.disparity), "pds_regularization_forward: null pointer");
  PDS_CHECK_ARG(B >= 0 && D >= 16 && H >= 16 && W >= 16 && D % 16 == 0 && H % 16 == 0 && W % 16 == 0,
                "pds_regularization_forward: D, H, W must be positive multiples of 16");
  if (B == 0) return PDS_OK;
  if (!workspace || workspace_bytes < pds_regularization_workspace_bytes(reg, B, D, H, W) ||
      ((uintptr_t)workspace & 255)) {
    set_error("pds_regularization_forward: workspace too small or not 256-byte aligned");
    return PDS_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (tf.disparity && !reg->fused_tail) {
    set_error("pds_regularization_forward_disparity: needs the 8-feature hourglass");
    return PDS_ERR_UNSUPPORTED;
  }
  if (reg->raw) return tcg_forward(reg, signatures, shortcut, cost, B, D, H, W, workspace, workspace_bytes, st, tf);
  const int F = reg->F;
  const size_t vox = (size_t)D * H * W, hw = (size_t)H * W;
  Workspace ws(workspace, workspace_bytes);
  float* sig_cl = ws.take<float>(B * vox * F);
  float* sc_cl = ws.take<float>((size_t)B * hw * F);
  float* skip[4];
  float* sum = nullptr;
  skip[0] = ws.take<float>(B * vox * F);
  sum = ws.take<float>(B * vox * F);
  double* stats = nullptr;
  // remaining buffers are taken as the pipeline advances; stats last would break the
  // accounting, so reserve it now
  stats = ws.take<double>((size_t)B * 16 * F * 2);
  float* tail_state = (float*)ws.take<char>(hourglass_tail_state_bytes(B, 2 * D, 2 * H, 2 * W));
  if (ws.overflow) { set_error("pds_regularization_forward: workspace overflow"); return PDS_ERR_WORKSPACE; }

  int rc;
  if ((rc = nchw_to_nhwc(signatures, sig_cl, B, F, vox, st)) != PDS_OK) return rc;
  if ((rc = nchw_to_nhwc(shortcut, sc_cl, B, F, hw, st)) != PDS_OK) return rc;

  const std::vector<ConvLayer>& L = reg->layers;
  int li = 0;
  ConvGeom g; g.N = B; g.n_div = 1; g.D = D; g.H = H; g.W = W;
  // output = smoothing(signatures); level-0 input = shortcut (broadcast over D) + output
  // (regularization.py:115-119)
  if ((rc = conv_block(L[li++], g, sig_cl, skip[0], stats, nullptr, sc_cl, skip[0], sum, hw, st)) != PDS_OK) return rc;
  int c = F;
  float* out = nullptr;
  for (int k = 0; k < 4; ++k) {
    const size_t n_out = (size_t)B * (g.D / 2) * (g.H / 2) * (g.W / 2) * 2 * c;
    float* down = ws.take<float>(n_out);
    float* smooth = ws.take<float>(n_out);
    float* next_sum = ws.take<float>(n_out);
    if (ws.overflow) { set_error("pds_regularization_forward: workspace overflow"); return PDS_ERR_WORKSPACE; }
    // ContractionBlock3d (regularization.py:28-31)
    if ((rc = conv_block(L[li++], g, sum, down, stats, nullptr, nullptr, down, nullptr, 0, st)) != PDS_OK) return rc;
    g.D /= 2; g.H /= 2; g.W /= 2; c *= 2;
    // smooth_k is the next skip; next level input = down_k + smooth_k
    if ((rc = conv_block(L[li++], g, down, smooth, stats, down, nullptr, smooth, k < 3 ? next_sum : nullptr, 0, st)) != PDS_OK) return rc;
    if (k < 3) skip[k + 1] = smooth;
    sum = next_sum;
    out = smooth;
  }
  for (int k = 0; k < 4; ++k) {
    // ExpansionBlock3d (regularization.py:54-57): smoothing(up(out) + skip)
    const size_t n_out = (size_t)B * g.D * g.H * g.W * 8 * (c / 2);
    float* up = ws.take<float>(n_out);
    float* sm = ws.take<float>(n_out);
    if (ws.overflow) { set_error("pds_regularization_forward: workspace overflow"); return PDS_ERR_WORKSPACE; }
    if ((rc = conv_block(L[li++], g, out, up, stats, skip[3 - k], nullptr, nullptr, up, 0, st)) != PDS_OK) return rc;
    g.D *= 2; g.H *= 2; g.W *= 2; c /= 2;
    if ((rc = conv_block(L[li++], g, up, sm, stats, nullptr, nullptr, sm, nullptr, 0, st)) != PDS_OK) return rc;
    out = sm;
  }
  float* half = ws.take<float>((size_t)B * vox * 8 * (F / 2));
  if (ws.overflow) { set_error("pds_regularization_forward: workspace overflow"); return PDS_ERR_WORKSPACE; }
  if (reg->fused_tail) {
    // _upsample_to_halfsize: convolution + LeakyReLU + sums; its InstanceNorm is applied by the
    // tail kernel while it reads the volume (saves one 424 MB pass at C2)
    PDS_CUDA(cudaMemsetAsync(stats, 0, (size_t)B * L[li].Cout * 2 * sizeof(double), st));
    if ((rc = conv_forward_simt(L[li++], g, out, nullptr, 0, half, stats, st)) != PDS_OK) return rc;
    g.D *= 2; g.H *= 2; g.W *= 2;
    return hourglass_tail_forward(half, cost, stats, reg->tail_gamma, reg->tail_beta, reg->tail_w,
                                  reg->tail_bias, B, g.D, g.H, g.W, st, tf.disparity, tf.argmax, tf.R,
                                  tf.step, tf.crop_top, tf.crop_left, tail_state);
  }
  if ((rc = conv_block(L[li++], g, out, half, stats, nullptr, nullptr, half, nullptr, 0, st)) != PDS_OK) return rc;
  g.D *= 2; g.H *= 2; g.W *= 2;
  // _upsample_to_fullsize: 1 output channel, channels-last == (B, 2D, 4H, 4W) after squeeze(1)
  return conv_forward_simt(L[li], g, half, nullptr, 0, cost, nullptr, st);
}
}  // namespace
}  // namespace pds

extern "C" int pds_regularization_forward(pds_regularization* reg, const float* signatures,
                                          const float* shortcut, float* cost, int B, int D, int H,
                                          int W, void* workspace, size_t workspace_bytes,
                                          void* stream) {
  return pds::regularization_forward(reg, signatures, shortcut, cost, B, D, H, W, workspace,
                                     workspace_bytes, stream, pds::TailFusion());
}

extern "C" int pds_regularization_forward_disparity(pds_regularization* reg, const float* signatures,
                                                    const float* shortcut, float* disparity,
                                                    int64_t* argmax, int B, int D, int H, int W,
                                                    int half_support_window, int disparity_step,
                                                    int crop_top, int crop_left, void* workspace,
                                                    size_t workspace_bytes, void* stream) {
  using namespace pds;
  PDS_CHECK_ARG(disparity, "pds_regularization_forward_disparity: null pointer");
  // estimator.py:34-41
  PDS_CHECK_ARG(disparity_step >= 1, "\"disparity_step\" should be positive integer.");
  PDS_CHECK_ARG(half_support_window >= 1, "\"half_support_window\" should be positive integer.");
  PDS_CHECK_ARG(half_support_window % disparity_step == 0,
                "\"half_support_window\" should be multiple of the\"disparity_step\"");
  PDS_CHECK_ARG(crop_top >= 0 && crop_top < 4 * H && crop_left >= 0 && crop_left < 4 * W,
                "pds_regularization_forward_disparity: crop outside the image");
  TailFusion tf;
  tf.disparity = disparity; tf.argmax = argmax; tf.R = half_support_window / disparity_step;
  tf.step = disparity_step; tf.crop_top = crop_top; tf.crop_left = crop_left;
  if (tf.R > 4) {
    set_error("pds_regularization_forward_disparity: window radius above 4 is not fused");
    return PDS_ERR_UNSUPPORTED;
  }
  return regularization_forward(reg, signatures, shortcut, nullptr, B, D, H, W, workspace,
                                workspace_bytes, stream, tf);
}

// ---- individually tested blocks ------------------------------------------------

namespace pds {
namespace {
size_t block_blob_elems(std::vector<ConvLayer>& layers) { return blob_elems(layers); }
}  // namespace
}  // namespace pds

extern "C" size_t pds_contraction_block_workspace_bytes(int B, int C, int D, int H, int W) {
  using namespace pds;
  if (B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
  std::vector<ConvLayer> layers = {block3(C, 2 * C, 2), block3(2 * C, 2 * C, 1)};
  const size_t vin = (size_t)B * D * H * W * C;
  const size_t vout = (size_t)B * ((D - 1) / 2 + 1) * ((H - 1) / 2 + 1) * ((W - 1) / 2 + 1) * 2 * C;
  return align_up(block_blob_elems(layers) * 4, 256) + align_up(vin * 4, 256) +
         2 * align_up(vout * 4, 256) + align_up((size_t)B * 2 * C * 2 * 8, 256) + 1024;
}

extern "C" int pds_contraction_block_forward(const float* const* params, const float* in,
                                             float* down, float* smooth, int B, int C, int D,
                                             int H, int W, void* workspace, size_t workspace_bytes,
                                             void* stream) {
  using namespace pds;
  PDS_CHECK_ARG(params && in && down && smooth, "pds_contraction_block_forward: null pointer");
  PDS_CHECK_ARG(B >= 0 && C >= 1 && D >= 1 && H >= 1 && W >= 1, "pds_contraction_block_forward: bad shape");
  if (B == 0) return PDS_OK;
  if (!workspace || workspace_bytes < pds_contraction_block_workspace_bytes(B, C, D, H, W) ||
      ((uintptr_t)workspace & 255)) {
    set_error("pds_contraction_block_forward: workspace too small or not 256-byte aligned");
    return PDS_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<ConvLayer> layers = {block3(C, 2 * C, 2), block3(2 * C, 2 * C, 1)};
  Workspace ws(workspace, workspace_bytes);
  float* blob = ws.take<float>(block_blob_elems(layers));
  const int OD = (D - 1) / 2 + 1, OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  const size_t vin = (size_t)D * H * W, vout = (size_t)OD * OH * OW;
  float* in_cl = ws.take<float>(B * vin * C);
  float* d_cl = ws.take<float>(B * vout * 2 * C);
  float* s_cl = ws.take<float>(B * vout * 2 * C);
  double* stats = ws.take<double>((size_t)B * 2 * C * 2);
  if (ws.overflow) { set_error("pds_contraction_block_forward: workspace overflow"); return PDS_ERR_WORKSPACE; }
  int rc;
  if ((rc = load_layers(layers, params, blob, st)) != PDS_OK) return rc;
  if ((rc = nchw_to_nhwc(in, in_cl, B, C, vin, st)) != PDS_OK) return rc;
  ConvGeom g; g.N = B; g.D = D; g.H = H; g.W = W;
  if ((rc = conv_block(layers[0], g, in_cl, d_cl, stats, nullptr, nullptr, d_cl, nullptr, 0, st)) != PDS_OK) return rc;
  g.D = OD; g.H = OH; g.W = OW;
  if ((rc = conv_block(layers[1], g, d_cl, s_cl, stats, nullptr, nullptr, s_cl, nullptr, 0, st)) != PDS_OK) return rc;
  if ((rc = nhwc_to_nchw(d_cl, down, B, 2 * C, vout, st)) != PDS_OK) return rc;
  return nhwc_to_nchw(s_cl, smooth, B, 2 * C, vout, st);
}

extern "C" size_t pds_expansion_block_workspace_bytes(int B, int C, int D, int H, int W) {
  using namespace pds;
  if (B <= 0 || C <= 1 || D <= 0 || H <= 0 || W <= 0) return 0;
  std::vector<ConvLayer> layers = {tblock4(C, C / 2), block3(C / 2, C / 2, 1)};
  const size_t vin = (size_t)B * D * H * W * C, vout = (size_t)B * D * H * W * 8 * (C / 2);
  return align_up(block_blob_elems(layers) * 4, 256) + align_up(vin * 4, 256) +
         3 * align_up(vout * 4, 256) + align_up((size_t)B * C * 2 * 8, 256) + 1024;
}

extern "C" int pds_expansion_block_forward(const float* const* params, const float* in,
                                           const float* skip, float* out, int B, int C, int D,
                                           int H, int W, void* workspace, size_t workspace_bytes,
                                           void* stream) {
  using namespace pds;
  PDS_CHECK_ARG(params && in && skip && out, "pds_expansion_block_forward: null pointer");
  PDS_CHECK_ARG(B >= 0 && C >= 2 && C % 2 == 0 && D >= 1 && H >= 1 && W >= 1,
                "pds_expansion_block_forward: bad shape");
  if (B == 0) return PDS_OK;
  if (!workspace || workspace_bytes < pds_expansion_block_workspace_bytes(B, C, D, H, W) ||
      ((uintptr_t)workspace & 255)) {
    set_error("pds_expansion_block_forward: workspace too small or not 256-byte aligned");
    return PDS_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<ConvLayer> layers = {tblock4(C, C / 2), block3(C / 2, C / 2, 1)};
  Workspace ws(workspace, workspace_bytes);
  float* blob = ws.take<float>(block_blob_elems(layers));
  const size_t vin = (size_t)D * H * W, vout = vin * 8;
  float* in_cl = ws.take<float>(B * vin * C);
  float* skip_cl = ws.take<float>(B * vout * (C / 2));
  float* up = ws.take<float>(B * vout * (C / 2));
  float* sm = ws.take<float>(B * vout * (C / 2));
  double* stats = ws.take<double>((size_t)B * C * 2);
  if (ws.overflow) { set_error("pds_expansion_block_forward: workspace overflow"); return PDS_ERR_WORKSPACE; }
  int rc;
  if ((rc = load_layers(layers, params, blob, st)) != PDS_OK) return rc;
  if ((rc = nchw_to_nhwc(in, in_cl, B, C, vin, st)) != PDS_OK) return rc;
  if ((rc = nchw_to_nhwc(skip, skip_cl, B, C / 2, vout, st)) != PDS_OK) return rc;
  ConvGeom g; g.N = B; g.D = D; g.H = H; g.W = W;
  if ((rc = conv_block(layers[0], g, in_cl, up, stats, skip_cl, nullptr, nullptr, up, 0, st)) != PDS_OK) return rc;
  g.D *= 2; g.H *= 2; g.W *= 2;
  if ((rc = conv_block(layers[1], g, up, sm, stats, nullptr, nullptr, sm, nullptr, 0, st)) != PDS_OK) return rc;
  return nhwc_to_nchw(sm, out, B, C / 2, vout, st);
}
