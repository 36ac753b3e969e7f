/*
 * pds_oracle.c -- CPU restatement of the PdsNetwork.forward hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under practicaldeepstereo_nips2018_b200/
 * may import, link or call this file; it is the checker for tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * The reference (tlkvstepan/PracticalDeepStereo_NIPS2018) is pure Python on
 * top of PyTorch: every number on the path is produced by ATen operators
 * (third-party dependency: "pytorch 1.0", README.md:5; the container runs
 * torch 2.11.0).  This file restates the published semantics of those
 * operators (Conv2d / Conv3d / ConvTranspose3d cross-correlation with zero
 * padding, LeakyReLU, InstanceNorm with biased variance and eps inside the
 * square root, max with lowest-index tie break, softmax) and the reference's
 * own composition of them.  Each function cites the reference file:line it
 * follows.
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against
 *   (1) the reference's known-answer tests (test/test_matching.py:17-32,
 *       test/test_estimator.py:14-27), and
 *   (2) tests/golden/*.npz -- outputs of the UNMODIFIED reference modules run
 *       in the build container by oracle/make_golden.py.
 *
 * Layouts are PyTorch's: NCHW / NCDHW, float32, contiguous.
 * Parameters of composite modules are passed as ONE flat float array holding
 * the tensors of the module's state_dict() in state_dict() order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define PDS_LRELU_SLOPE 0.1f /* network_blocks.py:57 */
#define PDS_IN_EPS 1e-5      /* torch.nn.InstanceNorm default */

int pds_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void pds_oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------ */
/* Primitive operators (torch.nn semantics)                                   */
/* ------------------------------------------------------------------------ */

/* nn.Conv3d (network_blocks.py:9-16, 61-72); Conv2d is the D == 1 case
 * (network_blocks.py:19-34, 47-58).  w is (Cout, Cin, kD, kH, kW). */
void pds_oracle_conv3d(const float* in, const float* w, const float* bias,
                       float* out, int N, int Cin, int D, int H, int W,
                       int Cout, int kD, int kH, int kW, int sD, int sH,
                       int sW, int pD, int pH, int pW) {
  const int OD = (D + 2 * pD - kD) / sD + 1;
  const int OH = (H + 2 * pH - kH) / sH + 1;
  const int OW = (W + 2 * pW - kW) / sW + 1;
  const long in_plane = (long)D * H * W;
  const long out_plane = (long)OD * OH * OW;
#pragma omp parallel for collapse(2) schedule(dynamic)
  for (int n = 0; n < N; ++n) {
    for (int co = 0; co < Cout; ++co) {
      float* o = out + ((long)n * Cout + co) * out_plane;
      const float b = bias ? bias[co] : 0.0f;
      for (long i = 0; i < out_plane; ++i) o[i] = b;
      for (int ci = 0; ci < Cin; ++ci) {
        const float* x = in + ((long)n * Cin + ci) * in_plane;
        const float* wk = w + ((long)co * Cin + ci) * kD * kH * kW;
        for (int kd = 0; kd < kD; ++kd)
          for (int kh = 0; kh < kH; ++kh)
            for (int kw = 0; kw < kW; ++kw) {
              const float wv = wk[(kd * kH + kh) * kW + kw];
              /* valid output range along w for this tap */
              int ow_lo = 0, ow_hi = OW;
              while (ow_lo < OW && ow_lo * sW - pW + kw < 0) ++ow_lo;
              while (ow_hi > ow_lo && (ow_hi - 1) * sW - pW + kw >= W) --ow_hi;
              for (int od = 0; od < OD; ++od) {
                const int id = od * sD - pD + kd;
                if (id < 0 || id >= D) continue;
                for (int oh = 0; oh < OH; ++oh) {
                  const int ih = oh * sH - pH + kh;
                  if (ih < 0 || ih >= H) continue;
                  const float* xr = x + ((long)id * H + ih) * W - pW + kw;
                  float* orow = o + ((long)od * OH + oh) * OW;
                  if (sW == 1) {
                    for (int ow = ow_lo; ow < ow_hi; ++ow)
                      orow[ow] += wv * xr[ow];
                  } else {
                    for (int ow = ow_lo; ow < ow_hi; ++ow)
                      orow[ow] += wv * xr[ow * sW];
                  }
                }
              }
            }
      }
    }
  }
}

/* nn.ConvTranspose3d (network_blocks.py:37-44, 75-85).  w is
 * (Cin, Cout, kD, kH, kW); out = (in-1)*s - 2p + k. */
void pds_oracle_conv_transpose3d(const float* in, const float* w,
                                 const float* bias, float* out, int N, int Cin,
                                 int D, int H, int W, int Cout, int kD, int kH,
                                 int kW, int sD, int sH, int sW, int pD,
                                 int pH, int pW) {
  const int OD = (D - 1) * sD - 2 * pD + kD;
  const int OH = (H - 1) * sH - 2 * pH + kH;
  const int OW = (W - 1) * sW - 2 * pW + kW;
  const long in_plane = (long)D * H * W;
  const long out_plane = (long)OD * OH * OW;
#pragma omp parallel for collapse(2) schedule(dynamic)
  for (int n = 0; n < N; ++n) {
    for (int co = 0; co < Cout; ++co) {
      float* o = out + ((long)n * Cout + co) * out_plane;
      const float b = bias ? bias[co] : 0.0f;
      for (long i = 0; i < out_plane; ++i) o[i] = b;
      for (int ci = 0; ci < Cin; ++ci) {
        const float* x = in + ((long)n * Cin + ci) * in_plane;
        const float* wk = w + ((long)ci * Cout + co) * kD * kH * kW;
        for (int kd = 0; kd < kD; ++kd)
          for (int kh = 0; kh < kH; ++kh)
            for (int kw = 0; kw < kW; ++kw) {
              const float wv = wk[(kd * kH + kh) * kW + kw];
              for (int id = 0; id < D; ++id) {
                const int od = id * sD - pD + kd;
                if (od < 0 || od >= OD) continue;
                for (int ih = 0; ih < H; ++ih) {
                  const int oh = ih * sH - pH + kh;
                  if (oh < 0 || oh >= OH) continue;
                  const float* xr = x + ((long)id * H + ih) * W;
                  float* orow = o + ((long)od * OH + oh) * OW - pW + kw;
                  for (int iw = 0; iw < W; ++iw) {
                    const int ow = iw * sW - pW + kw;
                    if (ow < 0 || ow >= OW) continue;
                    orow[iw * sW] += wv * xr[iw];
                  }
                }
              }
            }
      }
    }
  }
}

/* nn.LeakyReLU(0.1, inplace=True) (network_blocks.py:57,71,84). */
void pds_oracle_leaky_relu(float* x, long n) {
#pragma omp parallel for
  for (long i = 0; i < n; ++i) x[i] = x[i] > 0.0f ? x[i] : PDS_LRELU_SLOPE * x[i];
}

/* nn.InstanceNorm{2,3}d(affine) in eval and train mode alike (no running
 * stats): per (n, c) over S spatial elements, biased variance, eps = 1e-5
 * (network_blocks.py:58,72,85; embedding.py:32 with gamma == NULL). */
void pds_oracle_instance_norm(float* x, const float* gamma, const float* beta,
                              int N, int C, long S) {
#pragma omp parallel for collapse(2)
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c) {
      float* p = x + ((long)n * C + c) * S;
      double s = 0.0;
      for (long i = 0; i < S; ++i) s += p[i];
      const double mean = s / (double)S;
      double v = 0.0;
      for (long i = 0; i < S; ++i) {
        const double d = p[i] - mean;
        v += d * d;
      }
      const float rstd = (float)(1.0 / sqrt(v / (double)S + PDS_IN_EPS));
      const float g = gamma ? gamma[c] : 1.0f;
      const float b = beta ? beta[c] : 0.0f;
      const float m = (float)mean;
      for (long i = 0; i < S; ++i) p[i] = (p[i] - m) * rstd * g + b;
    }
}

static void add_inplace(float* a, const float* b, long n) {
#pragma omp parallel for
  for (long i = 0; i < n; ++i) a[i] += b[i];
}

/* ------------------------------------------------------------------------ */
/* 2-D blocks (D == 1)                                                        */
/* ------------------------------------------------------------------------ */

static void conv2d(const float* in, const float* w, const float* b, float* out,
                   int N, int Cin, int H, int W, int Cout, int k, int s) {
  pds_oracle_conv3d(in, w, b, out, N, Cin, 1, H, W, Cout, 1, k, k, 1, s, s, 0,
                    k / 2, k / 2);
}

void pds_oracle_conv2d(const float* in, const float* w, const float* b,
                       float* out, int N, int Cin, int H, int W, int Cout,
                       int k, int s) {
  conv2d(in, w, b, out, N, Cin, H, W, Cout, k, s);
}

/* convolution_block_2D_with_relu_and_instance_norm (network_blocks.py:47-58) */
static const float* block2d(const float* in, const float* p, float* out, int N,
                            int Cin, int H, int W, int Cout, int k, int s) {
  const float* w = p;
  const float* b = w + (long)Cout * Cin * k * k;
  const float* g = b + Cout;
  const float* be = g + Cout;
  const int OH = (H + 2 * (k / 2) - k) / s + 1, OW = (W + 2 * (k / 2) - k) / s + 1;
  conv2d(in, w, b, out, N, Cin, H, W, Cout, k, s);
  pds_oracle_leaky_relu(out, (long)N * Cout * OH * OW);
  pds_oracle_instance_norm(out, g, be, N, Cout, (long)OH * OW);
  return be + Cout;
}

/* network_blocks.ResidualBlock (network_blocks.py:134-144): two 3x3 blocks
 * plus identity, no activation after the sum. x is updated in place. */
static const float* residual_block2d(float* x, const float* p, float* t0,
                                     float* t1, int N, int C, int H, int W) {
  p = block2d(x, p, t0, N, C, H, W, C, 3, 1);
  p = block2d(t0, p, t1, N, C, H, W, C, 3, 1);
  add_inplace(x, t1, (long)N * C * H * W);
  return p;
}

/* MatchingOperation.forward (matching.py:69-112).
 * in (N, Cin, H, W) -> out (N, Csig, H, W); defaults Cin=128, F=64, Csig=8,
 * 2 residual blocks.  params: state_dict order of MatchingOperation. */
long pds_oracle_matching_operation(const float* in, const float* params,
                                   float* out, int N, int Cin, int F, int Csig,
                                   int n_res, int H, int W) {
  const long plane = (long)N * F * H * W;
  float* x = (float*)malloc(sizeof(float) * plane);
  float* t0 = (float*)malloc(sizeof(float) * plane);
  float* t1 = (float*)malloc(sizeof(float) * plane);
  const float* p = params;
  conv2d(in, p, p + (long)F * Cin * 9, x, N, Cin, H, W, F, 3, 1);
  p += (long)F * Cin * 9 + F;
  for (int r = 0; r < n_res; ++r) p = residual_block2d(x, p, t0, t1, N, F, H, W);
  conv2d(x, p, p + (long)Csig * F * 9, out, N, F, H, W, Csig, 3, 1);
  p += (long)Csig * F * 9 + Csig;
  free(x);
  free(t0);
  free(t1);
  return (long)(p - params);
}

/* Matching.forward, data movement part (matching.py:50-62): for every
 * disparity d the tensor handed to `operation`:
 *   cat[left, shift_d(right)], shift_d(right)[x] = right[x-d], 0 for x < d.
 * out is (B, D, 2C, H, W): slice [b, d] is the operation input of disparity d
 * (D = maximum_disparity + 1). */
void pds_oracle_matching_concat(const float* left, const float* right,
                                float* out, int B, int C, int H, int W, int D) {
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int d = 0; d < D; ++d) {
      float* o = out + ((long)b * D + d) * 2 * C * H * W;
      memcpy(o, left + (long)b * C * H * W, sizeof(float) * C * H * W);
      o += (long)C * H * W;
      const float* r = right + (long)b * C * H * W;
      for (long row = 0; row < (long)C * H; ++row)
        for (int x = 0; x < W; ++x)
          o[row * W + x] = x >= d ? r[row * W + x - d] : 0.0f;
    }
}

/* Matching.forward with MatchingOperation (matching.py:34-63): output
 * (B, Csig, D, H, W) -- the th.stack(dim=2) of the per-disparity results. */
void pds_oracle_matching(const float* left, const float* right,
                         const float* params, float* out, int B, int C, int F,
                         int Csig, int n_res, int H, int W, int D) {
  const long hw = (long)H * W;
  float* cat = (float*)malloc(sizeof(float) * B * 2 * C * hw);
  float* sig = (float*)malloc(sizeof(float) * B * Csig * hw);
  for (int d = 0; d < D; ++d) {
    for (int b = 0; b < B; ++b) {
      float* o = cat + (long)b * 2 * C * hw;
      memcpy(o, left + (long)b * C * hw, sizeof(float) * C * hw);
      o += C * hw;
      const float* r = right + (long)b * C * hw;
      for (long row = 0; row < (long)C * H; ++row)
        for (int x = 0; x < W; ++x)
          o[row * W + x] = x >= d ? r[row * W + x - d] : 0.0f;
    }
    pds_oracle_matching_operation(cat, params, sig, B, 2 * C, F, Csig, n_res, H, W);
    for (int b = 0; b < B; ++b)
      for (int c = 0; c < Csig; ++c)
        memcpy(out + (((long)b * Csig + c) * D + d) * hw,
               sig + ((long)b * Csig + c) * hw, sizeof(float) * hw);
  }
  free(cat);
  free(sig);
}

/* ------------------------------------------------------------------------ */
/* 3-D blocks and the hourglass (regularization.py)                           */
/* ------------------------------------------------------------------------ */

/* convolution_block_3D_with_relu_and_instance_norm (network_blocks.py:61-72),
 * k = 3, stride s in all three dims. */
static const float* block3d(const float* in, const float* p, float* out, int N,
                            int Cin, int D, int H, int W, int Cout, int s) {
  const float* w = p;
  const float* b = w + (long)Cout * Cin * 27;
  const float* g = b + Cout;
  const float* be = g + Cout;
  const int OD = (D + 2 - 3) / s + 1, OH = (H + 2 - 3) / s + 1,
            OW = (W + 2 - 3) / s + 1;
  pds_oracle_conv3d(in, w, b, out, N, Cin, D, H, W, Cout, 3, 3, 3, s, s, s, 1, 1, 1);
  pds_oracle_leaky_relu(out, (long)N * Cout * OD * OH * OW);
  pds_oracle_instance_norm(out, g, be, N, Cout, (long)OD * OH * OW);
  return be + Cout;
}

/* transposed_convolutional_block_4x4x4_stride_2 (network_blocks.py:75-85,
 * 124-131): ConvT k4 s2 p1 -> LReLU -> IN. Doubles D, H, W. */
static const float* tblock3d(const float* in, const float* p, float* out,
                             int N, int Cin, int D, int H, int W, int Cout) {
  const float* w = p;
  const float* b = w + (long)Cin * Cout * 64;
  const float* g = b + Cout;
  const float* be = g + Cout;
  pds_oracle_conv_transpose3d(in, w, b, out, N, Cin, D, H, W, Cout, 4, 4, 4, 2,
                              2, 2, 1, 1, 1);
  pds_oracle_leaky_relu(out, (long)N * Cout * 8 * D * H * W);
  pds_oracle_instance_norm(out, g, be, N, Cout, (long)8 * D * H * W);
  return be + Cout;
}

/* Regularization.forward (regularization.py:94-126).
 * sig (B, F, D, H, W), shortcut (B, F, H, W) -> cost (B, 2D, 4H, 4W).
 * D, H, W must be divisible by 16.  params: state_dict order. */
long pds_oracle_regularization(const float* sig, const float* shortcut,
                               const float* params, float* cost, int B, int F,
                               int D, int H, int W) {
  const long vox = (long)D * H * W;
  const float* p = params;
  float* skips[4];
  float* cur = (float*)malloc(sizeof(float) * B * F * vox);
  /* output = self._smoothing(matching_signatures)  (regularization.py:116) */
  p = block3d(sig, p, cur, B, F, D, H, W, F, 1);
  /* shortcut = shortcut_from_left_image.unsqueeze(2): broadcast along D */
  float* sc = (float*)malloc(sizeof(float) * B * F * vox);
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < F; ++c)
      for (int d = 0; d < D; ++d)
        memcpy(sc + (((long)b * F + c) * D + d) * H * W,
               shortcut + ((long)b * F + c) * H * W, sizeof(float) * H * W);
  int c = F, d = D, h = H, w = W;
  for (int k = 0; k < 4; ++k) {
    /* shortcuts.append(output); shortcut, output = block(shortcut + output)
     * (regularization.py:117-119); block = ContractionBlock3d (:28-31). */
    skips[k] = cur;
    const long n_in = (long)B * c * d * h * w;
    float* sum = (float*)malloc(sizeof(float) * n_in);
    memcpy(sum, sc, sizeof(float) * n_in);
    add_inplace(sum, cur, n_in);
    free(sc);
    const long n_out = (long)B * 2 * c * (d / 2) * (h / 2) * (w / 2);
    float* down = (float*)malloc(sizeof(float) * n_out);
    float* smooth = (float*)malloc(sizeof(float) * n_out);
    p = block3d(sum, p, down, B, c, d, h, w, 2 * c, 2);
    free(sum);
    c *= 2; d /= 2; h /= 2; w /= 2;
    p = block3d(down, p, smooth, B, c, d, h, w, c, 1);
    sc = down;
    cur = smooth;
  }
  free(sc);
  for (int k = 0; k < 4; ++k) {
    /* ExpansionBlock3d.forward (regularization.py:54-57) */
    const long n_out = (long)B * (c / 2) * 8 * d * h * w;
    float* up = (float*)malloc(sizeof(float) * n_out);
    p = tblock3d(cur, p, up, B, c, d, h, w, c / 2);
    free(cur);
    c /= 2; d *= 2; h *= 2; w *= 2;
    add_inplace(up, skips[3 - k], n_out);
    free(skips[3 - k]);
    cur = (float*)malloc(sizeof(float) * n_out);
    p = block3d(up, p, cur, B, c, d, h, w, c, 1);
    free(up);
  }
  /* _upsample_to_halfsize then _upsample_to_fullsize (regularization.py:125) */
  float* half = (float*)malloc(sizeof(float) * B * (F / 2) * 8 * vox);
  p = tblock3d(cur, p, half, B, F, D, H, W, F / 2);
  free(cur);
  const float* wf = p;
  const float* bf = wf + (long)(F / 2) * 1 * 3 * 4 * 4;
  pds_oracle_conv_transpose3d(half, wf, bf, cost, B, F / 2, 2 * D, 2 * H, 2 * W,
                              1, 3, 4, 4, 1, 2, 2, 1, 1, 1);
  p = bf + 1;
  free(half);
  return (long)(p - params);
}

/* ContractionBlock3d.forward (regularization.py:28-31): returns both the
 * down-sampled tensor and its smoothed version. */
long pds_oracle_contraction_block(const float* in, const float* params,
                                  float* down, float* smooth, int N, int C,
                                  int D, int H, int W) {
  const float* p = block3d(in, params, down, N, C, D, H, W, 2 * C, 2);
  p = block3d(down, p, smooth, N, 2 * C, (D - 1) / 2 + 1, (H - 1) / 2 + 1,
              (W - 1) / 2 + 1, 2 * C, 1);
  return (long)(p - params);
}

/* ExpansionBlock3d.forward (regularization.py:54-57). C = input features. */
long pds_oracle_expansion_block(const float* in, const float* skip,
                                const float* params, float* out, int N, int C,
                                int D, int H, int W) {
  const long n_out = (long)N * (C / 2) * 8 * D * H * W;
  float* up = (float*)malloc(sizeof(float) * n_out);
  const float* p = tblock3d(in, params, up, N, C, D, H, W, C / 2);
  add_inplace(up, skip, n_out);
  p = block3d(up, p, out, N, C / 2, 2 * D, 2 * H, 2 * W, C / 2, 1);
  free(up);
  return (long)(p - params);
}

/* ------------------------------------------------------------------------ */
/* Embedding (embedding.py:46-65) -- adjacent to the hot path                 */
/* ------------------------------------------------------------------------ */

/* image (N, 3, H, W) -> descriptor (N, F, H/4, W/4), shortcut (N, Fs, H/4, W/4) */
long pds_oracle_embedding(const float* image, const float* params,
                          float* descriptor, float* shortcut, int N, int Cimg,
                          int F, int Fs, int n_res, int H, int W) {
  const long n_img = (long)N * Cimg * H * W;
  float* x = (float*)malloc(sizeof(float) * n_img);
  memcpy(x, image, sizeof(float) * n_img);
  /* nn.InstanceNorm2d(3), non-affine (embedding.py:32) */
  pds_oracle_instance_norm(x, NULL, NULL, N, Cimg, (long)H * W);
  const int H2 = (H + 4 - 5) / 2 + 1, W2 = (W + 4 - 5) / 2 + 1;
  const int H4 = (H2 + 4 - 5) / 2 + 1, W4 = (W2 + 4 - 5) / 2 + 1;
  float* a = (float*)malloc(sizeof(float) * N * F * H2 * W2);
  const float* p = block2d(x, params, a, N, Cimg, H, W, F, 5, 2);
  free(x);
  p = block2d(a, p, descriptor, N, F, H2, W2, F, 5, 2);
  free(a);
  const long plane = (long)N * F * H4 * W4;
  float* t0 = (float*)malloc(sizeof(float) * plane);
  float* t1 = (float*)malloc(sizeof(float) * plane);
  for (int r = 0; r < n_res; ++r)
    p = residual_block2d(descriptor, p, t0, t1, N, F, H4, W4);
  free(t0);
  free(t1);
  p = block2d(descriptor, p, shortcut, N, F, H4, W4, Fs, 3, 1);
  return (long)(p - params);
}

/* ------------------------------------------------------------------------ */
/* SubpixelMap.__call__ (estimator.py:45-91)                                  */
/* ------------------------------------------------------------------------ */

/* sim (B, D, H, W) -> disparity (B, H, W); argmax (B, H, W) int64 optional.
 * th.max semantics: first maximum wins; a NaN beats everything and the first
 * NaN wins.  Window shifts run over range((-hsw)//step, hsw//step + 1)
 * (estimator.py:66-68; Python floor division).  Out-of-range taps have
 * similarity -inf (weight 0) and disparity 0 (estimator.py:71-83). */
void pds_oracle_subpixel_map(const float* sim, float* disparity,
                             int64_t* argmax, int B, int D, int H, int W,
                             int half_support_window, int disparity_step) {
  const long hw = (long)H * W;
  /* Python: -hsw // step == floor(-hsw / step) */
  int lo = -half_support_window / disparity_step;
  if ((-half_support_window) % disparity_step != 0) --lo;
  const int hi = half_support_window / disparity_step;
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (long i = 0; i < hw; ++i) {
      const float* s = sim + (long)b * D * hw + i;
      float best = s[0];
      int idx = 0;
      for (int d = 1; d < D; ++d) {
        const float v = s[(long)d * hw];
        if (best != best) break; /* first NaN already holds */
        if (v > best || v != v) {
          best = v;
          idx = d;
        }
      }
      if (argmax) argmax[(long)b * hw + i] = idx;
      /* softmax over the window (estimator.py:88-90), max-subtracted */
      float e[64], dsp[64];
      int n = 0;
      float sum = 0.0f;
      for (int sft = lo; sft <= hi; ++sft, ++n) {
        const int j = idx + sft;
        if (j < 0 || j >= D) {
          e[n] = 0.0f;
          dsp[n] = 0.0f;
        } else {
          e[n] = expf(s[(long)j * hw] - best);
          dsp[n] = (float)(disparity_step * j);
        }
        sum += e[n];
      }
      float acc = 0.0f;
      for (int k = 0; k < n; ++k) acc += (e[k] / sum) * dsp[k];
      disparity[(long)b * hw + i] = acc;
    }
}

/* ------------------------------------------------------------------------ */
/* SizeAdapter (size_adapter.py:29-52) and PdsNetwork.forward (network.py)    */
/* ------------------------------------------------------------------------ */

static int round_up(int v, int m) { return ((v + m - 1) / m) * m; }

/* Top/left zero pad to a multiple of `minimum_size` (size_adapter.py:29-43). */
void pds_oracle_pad(const float* in, float* out, int N, int C, int H, int W,
                    int minimum_size) {
  const int HP = round_up(H, minimum_size), WP = round_up(W, minimum_size);
  const int ph = HP - H, pw = WP - W;
  memset(out, 0, sizeof(float) * (long)N * C * HP * WP);
  for (long nc = 0; nc < (long)N * C; ++nc)
    for (int y = 0; y < H; ++y)
      memcpy(out + (nc * HP + y + ph) * WP + pw, in + (nc * H + y) * W,
             sizeof(float) * W);
}

/* PdsNetwork.forward in eval mode (network.py:38-52) with the default
 * modules (network.py:54-65).  left/right (B, 3, H, W) -> disparity (B, H, W).
 * params_* are the flat state_dicts of _embedding, _matching._operation and
 * _regularization.  If cost_out != NULL it receives the padded cost volume
 * (B, (md+1)/2, HP, WP). */
int pds_oracle_network_forward(const float* left, const float* right,
                               const float* params_embedding,
                               const float* params_matching,
                               const float* params_regularization,
                               float* disparity, float* cost_out, int B, int H,
                               int W, int maximum_disparity) {
  if ((maximum_disparity + 1) % 64 != 0) return 1; /* network.py:28-31 */
  const int HP = round_up(H, 64), WP = round_up(W, 64);
  const int Hq = HP / 4, Wq = WP / 4;
  const int Dq = (maximum_disparity + 1) / 4; /* network.py:36 */
  const long nimg = (long)B * 3 * HP * WP;
  float* lp = (float*)malloc(sizeof(float) * nimg);
  float* rp = (float*)malloc(sizeof(float) * nimg);
  pds_oracle_pad(left, lp, B, 3, H, W, 64);
  pds_oracle_pad(right, rp, B, 3, H, W, 64);
  const long ndesc = (long)B * 64 * Hq * Wq, nsc = (long)B * 8 * Hq * Wq;
  float* ld = (float*)malloc(sizeof(float) * ndesc);
  float* rd = (float*)malloc(sizeof(float) * ndesc);
  float* lsc = (float*)malloc(sizeof(float) * nsc);
  float* rsc = (float*)malloc(sizeof(float) * nsc);
  pds_oracle_embedding(lp, params_embedding, ld, lsc, B, 3, 64, 8, 2, HP, WP);
  pds_oracle_embedding(rp, params_embedding, rd, rsc, B, 3, 64, 8, 2, HP, WP);
  free(lp); free(rp); free(rsc);
  float* sig = (float*)malloc(sizeof(float) * B * 8 * Dq * Hq * Wq);
  pds_oracle_matching(ld, rd, params_matching, sig, B, 64, 64, 8, 2, Hq, Wq, Dq);
  free(ld); free(rd);
  const long ncost = (long)B * 2 * Dq * HP * WP;
  float* cost = cost_out ? cost_out : (float*)malloc(sizeof(float) * ncost);
  pds_oracle_regularization(sig, lsc, params_regularization, cost, B, 8, Dq, Hq, Wq);
  free(sig); free(lsc);
  float* dpad = (float*)malloc(sizeof(float) * B * HP * WP);
  pds_oracle_subpixel_map(cost, dpad, NULL, B, 2 * Dq, HP, WP, 4, 2);
  if (!cost_out) free(cost);
  /* unpad: crop [pad_h:, pad_w:] (size_adapter.py:51-52) */
  const int ph = HP - H, pw = WP - W;
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < H; ++y)
      memcpy(disparity + ((long)b * H + y) * W,
             dpad + ((long)b * HP + y + ph) * WP + pw, sizeof(float) * W);
  free(dpad);
  return 0;
}
