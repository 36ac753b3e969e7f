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
  int split = 0;                       // 16-bit terms per value (1, 2 or 3)
  int fp16 = 0;                        // term type (1 = half, 0 = bfloat16)
  int group = 0;                       // disparity slices per pass (0 = all of them)
  std::vector<pds::TcLayer> tc;
  // factorised first convolution (conv_tc.cu, tc_compose_first): left half + bias, right half,
  // right half's kx = 2 taps in place
  pds::TcLayer first[3];
  int factor = 0;
  // second level (matching_factor.cu): block 1's first convolution without bias, its weights as
  // [kx][dy][ci][co] fp32 for the column corrections
  pds::TcLayer second;
  float* wt1 = nullptr;
  int factor2 = 0;
  uint16_t* last_w = nullptr;          // last convolution with the taps on the M axis (conv_last.cu): packed weights
  int two_pass = 1;                    // tc_compose_second as sums + normalised planes (no fp32 round trip)
  int fuse_norm = 0;                   // InstanceNorm passes inside the convolution launches (conv_tc.cu, FUSE): opt-in
  int dynamic_conv = 0;                // 64 -> 64 layers on the dynamically scheduled kernel (eight epilogue warps): opt-in
  void* tc_blob = nullptr;
  // tensor maps of the shifted right descriptors (un-factorised first convolution only): host staging
  // here, the device copy lives in the CALLER'S workspace (one per stream: no sharing between streams)
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
  PDS_CHECK_ARG(precision >= PDS_PRECISION_FP32 && precision <= PDS_PRECISION_FP16,
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
    op->split = precision == PDS_PRECISION_BF16X3 ? 3
                : (precision == PDS_PRECISION_BF16X2 || precision == PDS_PRECISION_FP16X2) ? 2 : 1;
    op->fp16 = (precision == PDS_PRECISION_FP16X2 || precision == PDS_PRECISION_FP16) ? 1 : 0;
    if (const char* g = getenv("PDS_B200_MATCH_GROUP")) op->group = atoi(g);
    const int nl = 2 + 2 * n_res;
    op->tc.resize(nl);
    size_t bytes = 0;
    for (int i = 0; i < nl; ++i) {
      TcLayer& l = op->tc[i];
      l.Cin = i == 0 ? 2 * C : F;
      l.Cout = i == nl - 1 ? S : F;
      l.N = l.Cout <= 16 ? 16 : 64;
      l.S = op->split;
      l.fp16 = op->fp16;
      // fp16 terms: scale the weights by 2^8 so that the low term of typical (|w| << 1)
      // weights stays a normal half; undone exactly in the epilogue
      l.wscale = op->fp16 ? 256.f : 1.f;
      bytes += align_up(l.w_elems() * 2, 256) + align_up(l.N * 4, 256) + 2 * align_up(F * 4, 256);
    }
    op->factor = !(getenv("PDS_B200_MATCH_FACTOR") && atoi(getenv("PDS_B200_MATCH_FACTOR")) == 0);
    op->factor2 = op->factor && n_res >= 1 && C == F &&
                  !(getenv("PDS_B200_MATCH_FACTOR2") && atoi(getenv("PDS_B200_MATCH_FACTOR2")) == 0);
    op->two_pass = !(getenv("PDS_B200_COMPOSE_TWO_PASS") && atoi(getenv("PDS_B200_COMPOSE_TWO_PASS")) == 0);
    op->fuse_norm = tc_fused_norm_enabled() ? 1 : 0;
    op->dynamic_conv = tc_dynamic_conv_enabled() ? 1 : 0;
    {
      TcLayer& l = op->second;
      l.Cin = F; l.Cout = F; l.N = 64; l.S = op->split; l.fp16 = op->fp16; l.wscale = op->fp16 ? 256.f : 1.f;
      bytes += align_up(l.w_elems() * 2, 256) + align_up(l.N * 4, 256) + align_up((size_t)9 * F * F * 4, 256);
    }
    for (int i = 0; i < 3; ++i) {
      TcLayer& l = op->first[i];
      l.Cin = C; l.Cout = F; l.N = 64; l.S = op->split; l.fp16 = op->fp16; l.wscale = op->fp16 ? 256.f : 1.f;
      bytes += align_up(l.w_elems() * 2, 256) + align_up(l.N * 4, 256);
    }
    const bool last_taps = tc_last_enabled() && op->split == 2 && op->fp16 && S == 8 && F == 64;
    if (last_taps) bytes += align_up(tc_last_weight_bytes(), 256);
    cudaError_t e = cudaMalloc(&op->tc_blob, bytes);
    if (e != cudaSuccess) { delete op; return cuda_fail(e, "cudaMalloc(matching tensor-core weights)"); }
    char* cur = (char*)op->tc_blob;
    int pi = 0, rc = PDS_OK;
    for (int i = 0; i < nl && rc == PDS_OK; ++i) {
      TcLayer& l = op->tc[i];
      l.w = (uint16_t*)cur; cur += align_up(l.w_elems() * 2, 256);
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
    for (int i = 0; i < 3 && rc == PDS_OK; ++i) {
      TcLayer& l = op->first[i];
      l.w = (uint16_t*)cur; cur += align_up(l.w_elems() * 2, 256);
      l.bias = (float*)cur; cur += align_up(l.N * 4, 256);
      rc = tc_prepare_weights(l, params[0], i == 0 ? params[1] : nullptr, st, 2 * C, i == 0 ? 0 : C, i == 2);
    }
    if (rc == PDS_OK && op->factor2) {
      TcLayer& l = op->second;
      l.w = (uint16_t*)cur; cur += align_up(l.w_elems() * 2, 256);
      l.bias = (float*)cur; cur += align_up(l.N * 4, 256);
      op->wt1 = (float*)cur; cur += align_up((size_t)9 * F * F * 4, 256);
      rc = tc_prepare_weights(l, params[2], nullptr, st);
      if (rc == PDS_OK) rc = tc_transpose_weights(params[2], op->wt1, F, st);
    }
    if (rc == PDS_OK && last_taps) {
      op->last_w = (uint16_t*)cur; cur += align_up(tc_last_weight_bytes(), 256);
      rc = tc_last_prepare(params[n_params - 2], op->last_w, op->tc.back().wscale, st);
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
  free(op->maps_host);
  delete op;
}

namespace pds {
namespace {

// Tensor-core pipeline buffers (all sizes in bytes, 256-aligned).  The disparity
// slices n = b * D + d are processed in groups of `G`: the three activation
// buffers of a group (residual stream planes, fp32 convolution output,
// normalised planes) are sized to stay resident in the 126 MB L2 between the
// convolution that writes them and the pass that reads them.
struct TcPlan {
  int G;
  size_t lap, rap, xa, t, ya, stats, sched, first, ap2, cols, maps, total;
};

TcPlan tc_plan(const pds_matching_op* op, int B, int H, int W, int D) {
  const size_t hw = (size_t)H * W, n = (size_t)B * D, S = op->split;
  TcPlan p;
  const size_t per_slice = (size_t)op->F * hw * (2 * S + 4 + 2 * S);
  (void)per_slice;
  // Measured on B200 (C2, fp16x2): one pass over all 48 slices 3.3 ms, groups of 12 / 3 / 1
  // slices 4.4 / 7.9 / 12.2 ms -- per-launch fill/drain and the tail wave cost more than L2
  // residency returns, so the default is a single group; PDS_B200_MATCH_GROUP overrides.
  int G = op->group > 0 ? op->group : (int)n;
  if ((size_t)G > n) G = (int)n;
  p.G = G;
  p.lap = align_up((size_t)B * S * op->C * hw * 2, 256);
  p.rap = p.lap;
  p.xa = align_up((size_t)G * S * op->F * hw * 2, 256);
  p.ya = p.xa;
  p.t = align_up((size_t)G * op->F * hw * 4, 256);
  p.stats = align_up(n * op->F * 2 * sizeof(double) * 2 * (op->n_res > 0 ? op->n_res : 1), 256);
  // scheduler words of the fused convolution + normalisation launches (two per residual block)
  p.sched = tc_sched_bytes((int)n) * 2 * (op->n_res > 0 ? op->n_res : 1);
  p.first = align_up((size_t)B * op->F * hw * 4, 256);     // A, Bf, Q of the factorised first convolution
  p.ap2 = align_up((size_t)2 * B * S * op->F * hw * 2, 256);                         // planes of A and Bf
  p.cols = align_up((size_t)B * tc_column_jobs(D) * H * op->F * 4, 256);             // column corrections
  p.maps = align_up((size_t)(D > 64 ? D : 64) * sizeof(CUtensorMap), 256);
  p.total = p.lap + p.rap + p.xa + p.t + p.ya + p.stats + p.sched + 5 * p.first + p.ap2 + p.cols + p.maps + 1024;
  return p;
}

int tc_forward(pds_matching_op* op, const float* left, const float* right, float* signatures, int B,
               int H, int W, int D, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int N = B * D, S = op->split, fp16 = op->fp16;
  if (op->maps_cap < D) {
    free(op->maps_host);
    op->maps_host = nullptr; op->maps_cap = 0;
    const int cap = D > 64 ? D : 64;
    if (posix_memalign((void**)&op->maps_host, 64, cap * sizeof(CUtensorMap)) != 0) {
      set_error("out of host memory"); return PDS_ERR_CUDA;
    }
    op->maps_cap = cap;
  }
  Workspace ws(workspace, workspace_bytes);
  const TcPlan pl = tc_plan(op, B, H, W, D);
  uint16_t* lap = (uint16_t*)ws.take<char>(pl.lap);
  uint16_t* rap = (uint16_t*)ws.take<char>(pl.rap);
  uint16_t* xa = (uint16_t*)ws.take<char>(pl.xa);
  float* t = (float*)ws.take<char>(pl.t);
  uint16_t* ya = (uint16_t*)ws.take<char>(pl.ya);
  double* stats = (double*)ws.take<char>(pl.stats);
  char* sched = ws.take<char>(pl.sched);                 // directly behind the sums: one memset clears both
  float* fa = (float*)ws.take<char>(2 * pl.first);    // A then Bf, contiguous (one 2B-slice tensor)
  float* fb = fa + (size_t)B * op->F * H * W;
  float* fq = (float*)ws.take<char>(pl.first);
  float* fp = (float*)ws.take<char>(2 * pl.first);    // PA then PB
  uint16_t* ap2 = (uint16_t*)ws.take<char>(pl.ap2);
  float* cols = (float*)ws.take<char>(pl.cols);
  CUtensorMap* maps_dev = (CUtensorMap*)ws.take<char>(pl.maps);
  if (ws.overflow) { set_error("pds_matching_op_forward: workspace overflow"); return PDS_ERR_WORKSPACE; }
  const size_t stat_elems = (size_t)N * op->F * 2;
  PDS_CUDA(cudaMemsetAsync(stats, 0, pl.stats + pl.sched, st));
  int rc;
  if ((rc = tc_pack_nchw(left, lap, B, op->C, H, W, S, fp16, st)) != PDS_OK) return rc;
  if ((rc = tc_pack_nchw(right, rap, B, op->C, H, W, S, fp16, st)) != PDS_OK) return rc;
  // shifted-read tensor maps of the right descriptors (only the un-factorised first convolution reads
  // them): encoded per call into this call's workspace -- the copy from pageable host memory is staged
  // before cudaMemcpyAsync returns, so the host buffer is free again
  const bool need_maps = !(op->factor && pl.G == N);
  if (need_maps) {
    if ((rc = tc_encode_shift_maps(op->maps_host, rap, B, S, op->C, H, W, D)) != PDS_OK) return rc;
    PDS_CUDA(cudaMemcpyAsync(maps_dev, op->maps_host, D * sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
  }

  for (int n0 = 0; n0 < N; n0 += pl.G) {
    const int g = N - n0 < pl.G ? N - n0 : pl.G;
    TcConvArgs a = {};
    a.maps_dev = maps_dev; a.maps_host = op->maps_host;
    a.H = H; a.W = W; a.n_slices = g; a.n0 = n0; a.n_div = D;
    if (op->factor && pl.G == N) {
      // conv0 is linear in the concatenation: three convolutions over the B descriptors (instead
      // of one over B * D slices) and one composition pass that writes every slice's planes
      TcConvArgs f = a;
      f.n_slices = B; f.n0 = 0; f.n_div = 1; f.epilogue = TC_EPI_F32; f.in_slices = B; f.in_C = op->C;
      f.layer = &op->first[0]; f.in = lap; f.out_f32 = fa;
      if ((rc = tc_conv3x3(f, st)) != PDS_OK) return rc;
      f.layer = &op->first[1]; f.in = rap; f.out_f32 = fb;
      if ((rc = tc_conv3x3(f, st)) != PDS_OK) return rc;
      f.layer = &op->first[2]; f.out_f32 = fq;
      if ((rc = tc_conv3x3(f, st)) != PDS_OK) return rc;
      if (!(op->factor2 && W >= 4))
        if ((rc = tc_compose_first(fa, fb, fq, xa, B, op->F, H, W, D, S, fp16, st)) != PDS_OK) return rc;
    } else {
      // conv0: cat[left, shift_d(right)] gathered straight from the descriptors
      a.layer = &op->tc[0]; a.epilogue = TC_EPI_PLAIN;
      a.in = lap; a.in_slices = B; a.in_C = op->C; a.in2 = rap; a.in2_C = op->C;
      a.out_ap = xa;
      if ((rc = tc_conv3x3(a, st)) != PDS_OK) return rc;
    }
    a.in2 = nullptr; a.in2_C = 0; a.in_slices = g; a.in_C = op->F; a.out_ap = nullptr;
    const bool second = op->factor && pl.G == N && op->factor2 && W >= 4;
    for (int r = 0; r < op->n_res; ++r) {
      const TcLayer& c1 = op->tc[1 + 2 * r];
      const TcLayer& c2 = op->tc[2 + 2 * r];
      double* s1 = stats + stat_elems * (2 * r);
      double* s2 = stats + stat_elems * (2 * r + 1);
      const size_t soff = (size_t)n0 * op->F * 2;
      if (r == 0 && second) {
        // block 1's first convolution is linear too: conv1(A), conv1(Bf) over the descriptors, a few
        // single-column corrections, and one pass that writes every slice's activation + sums
        // (matching_factor.cu); x0 itself is never materialised
        if ((rc = tc_planes_to_ap(fa, ap2, 2 * B, op->F, H, W, S, fp16, st)) != PDS_OK) return rc;
        TcConvArgs f = a;
        f.n_slices = 2 * B; f.n0 = 0; f.n_div = 1; f.epilogue = TC_EPI_F32; f.in_slices = 2 * B; f.in_C = op->F;
        f.layer = &op->second; f.in = ap2; f.out_f32 = fp;
        if ((rc = tc_conv3x3(f, st)) != PDS_OK) return rc;
        if ((rc = tc_column_ops(fb, fq, op->wt1, cols, B, op->F, H, W, D, st)) != PDS_OK) return rc;
        const float* pb = fp + (size_t)B * op->F * H * W;
        if (op->two_pass) {
          // sums, then the values recomputed + normalised straight into the operand planes: the
          // fp32 activation (4 B/element out, 4 B/element back in) never touches memory
          if ((rc = tc_compose_second_stats(fp, pb, cols, c1.bias, s1, B, op->F, H, W, D, st)) != PDS_OK) return rc;
          if ((rc = tc_compose_second_norm(fp, pb, cols, c1.bias, s1, c1.gamma, c1.beta, ya, B, op->F, H, W, D, S,
                                           fp16, st)) != PDS_OK) return rc;
        } else {
          if ((rc = tc_compose_second(fp, pb, cols, c1.bias, t, s1, B, op->F, H, W, D, st)) != PDS_OK) return rc;
          if ((rc = tc_norm_split(t, s1 + soff, c1.gamma, c1.beta, nullptr, ya, g, op->F, H, W, S, fp16, st)) != PDS_OK) return rc;
        }
      } else {
        // y = IN(lrelu(conv(x))): the normalisation runs behind the convolution inside its launch
        a.layer = &c1; a.epilogue = TC_EPI_ACT; a.in = xa; a.out_f32 = t; a.stats = s1;
        a.norm_mode = TC_NORM_PLAIN; a.norm_out = ya; a.res_ap = nullptr;
        a.sched = (op->dynamic_conv && pl.G == N) ? sched + tc_sched_bytes(N) * (2 * r) : nullptr; a.fuse = op->fuse_norm;
        if ((rc = tc_conv3x3(a, st)) != PDS_OK) return rc;
      }
      // x = IN(lrelu(conv(y))) + x (ResidualBlock.forward, network_blocks.py:143-144); block 1's x0 is
      // rebuilt from the per-sample terms where the factorisation never materialised it
      a.layer = &c2; a.epilogue = TC_EPI_ACT; a.in = ya; a.out_f32 = t; a.stats = s2;
      a.norm_out = xa;
      a.sched = (op->dynamic_conv && pl.G == N) ? sched + tc_sched_bytes(N) * (2 * r + 1) : nullptr; a.fuse = op->fuse_norm;
      if (r == 0 && second) { a.norm_mode = TC_NORM_RESIDUAL_FIRST; a.fA = fa; a.fB = fb; a.fQ = fq; a.res_ap = nullptr; }
      else { a.norm_mode = TC_NORM_RESIDUAL; a.res_ap = xa; }
      if ((rc = tc_conv3x3(a, st)) != PDS_OK) return rc;
      a.norm_mode = TC_NORM_NONE; a.norm_out = nullptr; a.res_ap = nullptr; a.sched = nullptr; a.fuse = 0;
    }
    if (op->last_w && pl.G == N) {     // taps on the M axis: 12 MMAs of 64 cycles per 84 pixels
      if ((rc = tc_conv_last(op->last_w, op->tc.back().bias, op->tc.back().wscale, xa, signatures, N, D, H, W, st)) != PDS_OK)
        return rc;
      continue;
    }
    a.layer = &op->tc.back(); a.epilogue = TC_EPI_SIG; a.in = xa;
    a.out_f32 = nullptr; a.out_ap = nullptr; a.stats = nullptr; a.out_sig = signatures;
    if ((rc = tc_conv3x3(a, st)) != PDS_OK) return rc;
  }
  return PDS_OK;
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
