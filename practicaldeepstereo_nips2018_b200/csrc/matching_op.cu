// a2 -- MatchingOperation.forward over all disparities (reference
// matching.py:69-112 driven by the loop of matching.py:53-62) as one fused
// pipeline.  The concatenated volume cat[left, shift_d(right)] is never
// materialised: the first convolution gathers its two halves straight from the
// left / right descriptors (the right one at x - d with zero fill).
//
// Every disparity is an independent "slice" n = b*D + d (InstanceNorm
// statistics are per sample and channel, so disparities stack exactly on the
// batch axis).  Per residual block:
//   y = lrelu(conv(x)) [+ sum, sum^2]  ->  y = IN(y)
//   t = lrelu(conv(y)) [+ sum, sum^2]  ->  x = IN(t) + x
// (network_blocks.py:134-144).
#include <stdlib.h>

#include <new>
#include <vector>

#include "conv_layers.cuh"
#include "conv_tc.cuh"

struct pds_matching_op {
  int C, F, S, n_res, precision;
  std::vector<pds::ConvLayer> layers;  // conv0, 2 per residual block, conv_last
  float* blob = nullptr;               // all parameters, kernel layout (fp32 path)
  // tensor-core path (precision != fp32)
  int split = 0;                       // bf16 terms per value (1, 2 or 3)
  std::vector<pds::TcLayer> tc;
  void* tc_blob = nullptr;
  CUtensorMap* maps_dev = nullptr;
  CUtensorMap* maps_host = nullptr;
  int maps_cap = 0;
};

namespace pds {
namespace {

const DimSpec kUnit = {DM_UNIT, 1};
const DimSpec kConv3 = {DM_CONV3, 1};

}  // namespace
}  // namespace pds

extern "C" int pds_matching_op_create(pds_matching_op** out, const float* const* params,
                                      int n_params, int C, int F, int S, int n_res, int precision,
                                      void* stream) {
  using namespace pds;
  PDS_CHECK_ARG(out && params, "pds_matching_op_create: null pointer");
  PDS_CHECK_ARG(C >= 1 && F >= 1 && S >= 1 && n_res >= 0, "pds_matching_op_create: bad sizes");
  PDS_CHECK_ARG(n_params == 4 + 8 * n_res,
                "pds_matching_op_create: expected %d parameter tensors, got %d", 4 + 8 * n_res, n_params);
  PDS_CHECK_ARG(precision >= PDS_PRECISION_FP32 && precision <= PDS_PRECISION_BF16,
                "pds_matching_op_create: bad precision");
  cudaStream_t st = (cudaStream_t)stream;
  if (precision != PDS_PRECISION_FP32) {
    if (!tc_available()) {
      set_error("pds_matching_op_create: the driver does not export cuTensorMapEncodeTiled");
      return PDS_ERR_UNSUPPORTED;
    }
    if (C % 16 || F != 64 || S > 16) {
      set_error("pds_matching_op_create: tensor-core path needs C %% 16 == 0, F == 64, S <= 16");
      return PDS_ERR_UNSUPPORTED;
    }
  }
  pds_matching_op* op = new (std::nothrow) pds_matching_op();
  PDS_CHECK_ARG(op, "out of host memory");
  op->C = C; op->F = F; op->S = S; op->n_res = n_res; op->precision = precision;
  if (precision != PDS_PRECISION_FP32) {
    op->split = precision == PDS_PRECISION_BF16X3 ? 3 : (precision == PDS_PRECISION_BF16X2 ? 2 : 1);
    const int nl = 2 + 2 * n_res;
    op->tc.resize(nl);
    size_t bytes = 0;
    for (int i = 0; i < nl; ++i) {
      TcLayer& l = op->tc[i];
      l.Cin = i == 0 ? 2 * C : F;
      l.Cout = i == nl - 1 ? S : F;
      l.N = (int)align_up(l.Cout, 16);
      l.S = op->split;
      bytes += align_up(l.w_elems() * 2, 256) + align_up(l.N * 4, 256) + 2 * align_up(F * 4, 256);
    }
    cudaError_t e = cudaMalloc(&op->tc_blob, bytes);
    if (e != cudaSuccess) { delete op; return cuda_fail(e, "cudaMalloc(matching tensor-core weights)"); }
    char* cur = (char*)op->tc_blob;
    int pi = 0, rc = PDS_OK;
    for (int i = 0; i < nl && rc == PDS_OK; ++i) {
      TcLayer& l = op->tc[i];
      l.w = (__nv_bfloat16*)cur; cur += align_up(l.w_elems() * 2, 256);
      l.bias = (float*)cur; cur += align_up(l.N * 4, 256);
      rc = tc_prepare_weights(l, params[pi], params[pi + 1], st);
      pi += 2;
      if (i > 0 && i < nl - 1) {   // Conv -> LeakyReLU -> InstanceNorm block: gamma, beta
        float* g = (float*)cur; cur += align_up(F * 4, 256);
        float* b = (float*)cur; cur += align_up(F * 4, 256);
        cudaError_t e1 = cudaMemcpyAsync(g, params[pi], F * 4, cudaMemcpyDeviceToDevice, st);
        cudaError_t e2 = cudaMemcpyAsync(b, params[pi + 1], F * 4, cudaMemcpyDeviceToDevice, st);
        if (e1 != cudaSuccess || e2 != cudaSuccess) rc = cuda_fail(e1 != cudaSuccess ? e1 : e2, "cudaMemcpyAsync(parameters)");
        l.gamma = g; l.beta = b;
        pi += 2;
      }
    }
    if (rc != PDS_OK) { cudaFree(op->tc_blob); delete op; return rc; }
    *out = op;
    return PDS_OK;
  }
  op->layers.push_back(make_layer(2 * C, F, kUnit, kConv3, kConv3, false, false));
  for (int i = 0; i < 2 * n_res; ++i) op->layers.push_back(make_layer(F, F, kUnit, kConv3, kConv3, false, true));
  op->layers.push_back(make_layer(F, S, kUnit, kConv3, kConv3, false, false));
  size_t total = 0;
  for (auto& l : op->layers) total += align_up(l.weight_elems(), 64) + align_up(l.Cout, 64) * 3;
  cudaError_t e = cudaMalloc(&op->blob, total * sizeof(float));
  if (e != cudaSuccess) { delete op; return cuda_fail(e, "cudaMalloc(matching weights)"); }
  float* cur = op->blob;
  int pi = 0, rc = PDS_OK;
  for (auto& l : op->layers) {
    l.w = cur; cur += align_up(l.weight_elems(), 64);
    rc = relayout_weights(l, params[pi++], l.w, st);
    if (rc != PDS_OK) break;
    const int nvec = l.lrelu ? 3 : 1;  // bias [, gamma, beta]
    const float** dst[3] = {&l.bias, &l.gamma, &l.beta};
    for (int v = 0; v < nvec; ++v) {
      e = cudaMemcpyAsync(cur, params[pi++], l.Cout * sizeof(float), cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemcpyAsync(parameters)"); break; }
      *dst[v] = cur; cur += align_up(l.Cout, 64);
    }
    if (rc != PDS_OK) break;
  }
  if (rc != PDS_OK) { cudaFree(op->blob); delete op; return rc; }
  *out = op;
  return PDS_OK;
}

extern "C" void pds_matching_op_destroy(pds_matching_op* op) {
  if (!op) return;
  cudaFree(op->blob);
  cudaFree(op->tc_blob);
  cudaFree(op->maps_dev);
  free(op->maps_host);
  delete op;
}

namespace pds {
namespace {

// Tensor-core pipeline buffers (all sizes in bytes, 256-aligned).
struct TcPlan {
  size_t lap, rap, x, y, xap, aap, stats, total;
};

TcPlan tc_plan(const pds_matching_op* op, int B, int H, int W, int D) {
  const size_t hw = (size_t)H * W, n = (size_t)B * D, S = op->split;
  TcPlan p;
  p.lap = align_up((size_t)B * S * op->C * hw * 2, 256);
  p.rap = p.lap;
  p.x = align_up(n * op->F * hw * 4, 256);
  p.y = p.x;
  p.xap = align_up(n * S * op->F * hw * 2, 256);
  p.aap = p.xap;
  p.stats = align_up(n * op->F * 2 * sizeof(double) * 2 * (op->n_res > 0 ? op->n_res : 1), 256);
  p.total = p.lap + p.rap + p.x + p.y + p.xap + p.aap + p.stats + 1024;
  return p;
}

int tc_forward(pds_matching_op* op, const float* left, const float* right, float* signatures, int B,
               int H, int W, int D, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const size_t hw = (size_t)H * W;
  const int N = B * D, S = op->split;
  if (op->maps_cap < 1 + D) {
    cudaFree(op->maps_dev); free(op->maps_host);
    op->maps_dev = nullptr; op->maps_host = nullptr; op->maps_cap = 0;
    const int cap = 1 + (D > 64 ? D : 64);
    PDS_CUDA(cudaMalloc(&op->maps_dev, cap * sizeof(CUtensorMap)));
    if (posix_memalign((void**)&op->maps_host, 64, cap * sizeof(CUtensorMap)) != 0) {
      set_error("out of host memory"); return PDS_ERR_CUDA;
    }
    op->maps_cap = cap;
  }
  Workspace ws(workspace, workspace_bytes);
  const TcPlan pl = tc_plan(op, B, H, W, D);
  __nv_bfloat16* lap = (__nv_bfloat16*)ws.take<char>(pl.lap);
  __nv_bfloat16* rap = (__nv_bfloat16*)ws.take<char>(pl.rap);
  float* x = (float*)ws.take<char>(pl.x);
  float* y = (float*)ws.take<char>(pl.y);
  __nv_bfloat16* xap = (__nv_bfloat16*)ws.take<char>(pl.xap);
  __nv_bfloat16* aap = (__nv_bfloat16*)ws.take<char>(pl.aap);
  double* stats = (double*)ws.take<char>(pl.stats);
  if (ws.overflow) { set_error("pds_matching_op_forward: workspace overflow"); return PDS_ERR_WORKSPACE; }
  const size_t stat_elems = (size_t)N * op->F * 2;
  PDS_CUDA(cudaMemsetAsync(stats, 0, pl.stats, st));
  int rc;
  if ((rc = tc_pack_nchw(left, lap, B, op->C, H, W, S, st)) != PDS_OK) return rc;
  if ((rc = tc_pack_nchw(right, rap, B, op->C, H, W, S, st)) != PDS_OK) return rc;

  TcConvArgs a = {};
  a.maps_dev = op->maps_dev; a.maps_host = op->maps_host;
  a.H = H; a.W = W; a.n_slices = N;
  // conv0: cat[left, shift_d(right)] gathered straight from the descriptors
  a.layer = &op->tc[0]; a.epilogue = TC_EPI_PLAIN; a.n_div = D;
  a.in = lap; a.in_slices = B; a.in_C = op->C; a.in2 = rap; a.in2_C = op->C;
  a.out_f32 = x; a.out_ap = xap;
  if ((rc = tc_conv3x3(a, st)) != PDS_OK) return rc;
  a.in2 = nullptr; a.in2_C = 0; a.n_div = 1; a.in_slices = N; a.in_C = op->F;
  for (int r = 0; r < op->n_res; ++r) {
    const TcLayer& c1 = op->tc[1 + 2 * r];
    const TcLayer& c2 = op->tc[2 + 2 * r];
    double* s1 = stats + stat_elems * (2 * r);
    double* s2 = stats + stat_elems * (2 * r + 1);
    a.layer = &c1; a.epilogue = TC_EPI_ACT; a.in = xap; a.out_f32 = y; a.out_ap = nullptr; a.stats = s1;
    if ((rc = tc_conv3x3(a, st)) != PDS_OK) return rc;
    if ((rc = tc_norm_split(y, s1, c1.gamma, c1.beta, nullptr, nullptr, aap, N, op->F, H, W, S, st)) != PDS_OK) return rc;
    a.layer = &c2; a.in = aap; a.stats = s2;
    if ((rc = tc_conv3x3(a, st)) != PDS_OK) return rc;
    // x = IN(t) + x (ResidualBlock.forward, network_blocks.py:143-144)
    if ((rc = tc_norm_split(y, s2, c2.gamma, c2.beta, x, x, xap, N, op->F, H, W, S, st)) != PDS_OK) return rc;
  }
  a.layer = &op->tc.back(); a.epilogue = TC_EPI_SIG; a.in = xap; a.n_div = D;
  a.out_f32 = nullptr; a.out_ap = nullptr; a.stats = nullptr; a.out_sig = signatures;
  return tc_conv3x3(a, st);
}

}  // namespace
}  // namespace pds

extern "C" size_t pds_matching_op_workspace_bytes(const pds_matching_op* op, int B, int H, int W,
                                                  int D) {
  using namespace pds;
  if (!op || B <= 0 || H <= 0 || W <= 0 || D <= 0) return 0;
  if (op->precision != PDS_PRECISION_FP32) return tc_plan(op, B, H, W, D).total;
  const size_t hw = (size_t)H * W, n = (size_t)B * D;
  size_t bytes = 2 * align_up((size_t)B * hw * op->C * 4, 256);       // descriptors, channels-last
  bytes += 3 * align_up(n * hw * op->F * 4, 256);                    // x, y, t
  bytes += align_up(n * hw * op->S * 4, 256);                        // signatures, channels-last
  bytes += align_up(n * op->F * 2 * sizeof(double) * 2 * (op->n_res > 0 ? op->n_res : 1), 256);
  return bytes + 1024;
}

extern "C" int pds_matching_op_forward(pds_matching_op* op, const float* left, const float* right,
                                       float* signatures, int B, int H, int W, int D,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  using namespace pds;
  PDS_CHECK_ARG(op && left && right && signatures, "pds_matching_op_forward: null pointer");
  PDS_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && D >= 1, "pds_matching_op_forward: bad shape");
  if (B == 0) return PDS_OK;
  if (!workspace || workspace_bytes < pds_matching_op_workspace_bytes(op, B, H, W, D) ||
      ((uintptr_t)workspace & 255)) {
    set_error("pds_matching_op_forward: workspace too small or not 256-byte aligned");
    return PDS_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (op->precision != PDS_PRECISION_FP32)
    return tc_forward(op, left, right, signatures, B, H, W, D, workspace, workspace_bytes, st);
  const size_t hw = (size_t)H * W;
  const int N = B * D;
  Workspace ws(workspace, workspace_bytes);
  float* lcl = ws.take<float>((size_t)B * hw * op->C);
  float* rcl = ws.take<float>((size_t)B * hw * op->C);
  float* x = ws.take<float>((size_t)N * hw * op->F);
  float* y = ws.take<float>((size_t)N * hw * op->F);
  float* t = ws.take<float>((size_t)N * hw * op->F);
  float* sig = ws.take<float>((size_t)N * hw * op->S);
  const size_t stat_elems = (size_t)N * op->F * 2;
  const int n_stat = 2 * (op->n_res > 0 ? op->n_res : 1);
  double* stats = ws.take<double>(stat_elems * n_stat);
  if (ws.overflow) { set_error("pds_matching_op_forward: workspace overflow"); return PDS_ERR_WORKSPACE; }
  PDS_CUDA(cudaMemsetAsync(stats, 0, stat_elems * n_stat * sizeof(double), st));

  int rc;
  if ((rc = nchw_to_nhwc(left, lcl, B, op->C, hw, st)) != PDS_OK) return rc;
  if ((rc = nchw_to_nhwc(right, rcl, B, op->C, hw, st)) != PDS_OK) return rc;

  ConvGeom g0; g0.N = N; g0.n_div = D; g0.D = 1; g0.H = H; g0.W = W;   // reads the B descriptors
  ConvGeom g;  g.N = N;  g.n_div = 1;  g.D = 1;  g.H = H;  g.W = W;
  if ((rc = conv_forward_simt(op->layers[0], g0, lcl, rcl, op->C, x, nullptr, st)) != PDS_OK) return rc;
  for (int r = 0; r < op->n_res; ++r) {
    const ConvLayer& c1 = op->layers[1 + 2 * r];
    const ConvLayer& c2 = op->layers[2 + 2 * r];
    double* s1 = stats + stat_elems * (2 * r);
    double* s2 = stats + stat_elems * (2 * r + 1);
    if ((rc = conv_forward_simt(c1, g, x, nullptr, 0, y, s1, st)) != PDS_OK) return rc;
    if ((rc = instance_norm_apply(y, s1, c1.gamma, c1.beta, nullptr, nullptr, y, nullptr, N, hw, hw, op->F, st)) != PDS_OK) return rc;
    if ((rc = conv_forward_simt(c2, g, y, nullptr, 0, t, s2, st)) != PDS_OK) return rc;
    // x = IN(t) + x  (ResidualBlock.forward, network_blocks.py:143-144)
    if ((rc = instance_norm_apply(t, s2, c2.gamma, c2.beta, x, nullptr, nullptr, x, N, hw, hw, op->F, st)) != PDS_OK) return rc;
  }
  if ((rc = conv_forward_simt(op->layers.back(), g, x, nullptr, 0, sig, nullptr, st)) != PDS_OK) return rc;
  // [b][d][h][w][S] -> (B, S, D, H, W): th.stack(dim=2) of matching.py:63
  return nhwc_to_nchw(sig, signatures, B, op->S, (size_t)D * hw, st);
}
