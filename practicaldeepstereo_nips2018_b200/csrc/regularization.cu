// a3 -- Regularization.forward (reference regularization.py:94-126): the 3-D
// hourglass over the (B, 8, D, H, W) signature volume, plus the two blocks the
// reference tests individually (ContractionBlock3d :28-31, ExpansionBlock3d
// :54-57).  Activations are channels-last (NDHWC); every Conv -> LeakyReLU ->
// InstanceNorm block is one convolution launch whose epilogue applies bias and
// LeakyReLU and accumulates the InstanceNorm sums, followed by one
// normalise(+add) pass that also produces the sums the reference forms right
// after (shortcut + output, up-sampled + skip).
#include <new>
#include <vector>

#include "conv_layers.cuh"

struct pds_regularization {
  int F, precision;
  std::vector<pds::ConvLayer> layers;
  float* blob = nullptr;
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
  int rc = conv_forward_simt(l, g, in, nullptr, 0, y, stats, st);
  if (rc != PDS_OK) return rc;
  return instance_norm_apply(y, stats, l.gamma, l.beta, add, add_bcast, out, out2, g.N, S,
                             bcast_hw ? bcast_hw : S, l.Cout, st);
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
  if (rc != PDS_OK) { cudaFree(reg->blob); delete reg; return rc; }
  *out = reg;
  return PDS_OK;
}

extern "C" void pds_regularization_destroy(pds_regularization* reg) {
  if (!reg) return;
  cudaFree(reg->blob);
  delete reg;
}

extern "C" size_t pds_regularization_workspace_bytes(const pds_regularization* reg, int B, int D,
                                                     int H, int W) {
  using namespace pds;
  if (!reg || B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
  const size_t vox = (size_t)D * H * W, F = reg->F;
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
  return bytes + 1024;
}

extern "C" int pds_regularization_forward(pds_regularization* reg, const float* signatures,
                                          const float* shortcut, float* cost, int B, int D, int H,
                                          int W, void* workspace, size_t workspace_bytes,
                                          void* stream) {
  using namespace pds;
  PDS_CHECK_ARG(reg && signatures && shortcut && cost, "pds_regularization_forward: null pointer");
  PDS_CHECK_ARG(B >= 0 && D >= 16 && H >= 16 && W >= 16 && D % 16 == 0 && H % 16 == 0 && W % 16 == 0,
                "pds_regularization_forward: D, H, W must be positive multiples of 16");
  if (B == 0) return PDS_OK;
  if (!workspace || workspace_bytes < pds_regularization_workspace_bytes(reg, B, D, H, W) ||
      ((uintptr_t)workspace & 255)) {
    set_error("pds_regularization_forward: workspace too small or not 256-byte aligned");
    return PDS_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
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
                                  reg->tail_bias, B, g.D, g.H, g.W, st);
  }
  if ((rc = conv_block(L[li++], g, out, half, stats, nullptr, nullptr, half, nullptr, 0, st)) != PDS_OK) return rc;
  g.D *= 2; g.H *= 2; g.W *= 2;
  // _upsample_to_fullsize: 1 output channel, channels-last == (B, 2D, 4H, 4W) after squeeze(1)
  return conv_forward_simt(L[li], g, half, nullptr, 0, cost, nullptr, st);
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
