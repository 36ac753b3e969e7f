// Host-side layer descriptions shared by the convolution back ends
// (conv_simt.cu: fp32 CUDA cores; conv_tc.cu: tcgen05 tensor cores) and the
// two stage pipelines (matching_op.cu, regularization.cu).
//
// Internal activation layout is channels-last ("NDHWC", 2-D tensors have D=1):
// the contraction axis of the implicit GEMM (taps x channels) is then made of
// contiguous channel runs.  PyTorch-layout tensors are converted once at the
// stage boundary (layout.cu).
#pragma once
#include "pds_common.cuh"

namespace pds {

// How one spatial dimension of a layer maps output positions to input taps.
//   z     : position on the layer's "z-grid" (one GEMM row per z per class)
//   t     : tap index in [0, ntaps)
//   cls   : output parity class (only TCONV4S2 has two)
// DM_CONV3   Conv k3 p1 stride s : in = z*s + t - 1, k = t,        out = z
// DM_UNIT    kernel 1            : in = z,           k = 0,        out = z
// DM_TCONV4  ConvT k4 s2 p1      : in = z + cls - t, k = 1-cls+2t, out = 2z+cls
// DM_TCONV3  ConvT k3 s1 p1      : in = z + t - 1,   k = 2 - t,    out = z
// DM_CONV5   Conv k5 p2 stride s : in = z*s + t - 2, k = t,        out = z
enum DimMode { DM_CONV3 = 0, DM_UNIT = 1, DM_TCONV4 = 2, DM_TCONV3 = 3, DM_CONV5 = 4 };

struct DimSpec {
  int mode, stride;
  __host__ __device__ int ntaps() const {
    return mode == DM_UNIT ? 1 : (mode == DM_TCONV4 ? 2 : (mode == DM_CONV5 ? 5 : 3));
  }
  __host__ __device__ int nclass() const { return mode == DM_TCONV4 ? 2 : 1; }
  __host__ __device__ int ksize() const {
    return mode == DM_UNIT ? 1 : (mode == DM_TCONV4 ? 4 : (mode == DM_CONV5 ? 5 : 3));
  }
  __host__ __device__ int kidx(int t, int cls) const {
    return mode == DM_UNIT ? 0 : (mode == DM_TCONV4 ? 1 - cls + 2 * t : (mode == DM_TCONV3 ? 2 - t : t));
  }
  __host__ __device__ int in_coord(int z, int t, int cls) const {
    switch (mode) {
      case DM_CONV3: return z * stride + t - 1;
      case DM_CONV5: return z * stride + t - 2;
      case DM_UNIT: return z;
      case DM_TCONV4: return z + cls - t;
      default: return z + t - 1;
    }
  }
  __host__ __device__ int zsize(int in) const {
    return (mode == DM_CONV3 || mode == DM_CONV5) ? (in - 1) / stride + 1 : in;
  }
  __host__ __device__ int out_size(int in) const { return mode == DM_TCONV4 ? 2 * in : zsize(in); }
  __host__ __device__ int out_coord(int z, int cls) const { return mode == DM_TCONV4 ? 2 * z + cls : z; }
};

// One convolution (+ optional LeakyReLU + InstanceNorm parameters) with its
// weights already in kernel layout: w[class][tap][Cin][Cout], fp32.
struct ConvLayer {
  DimSpec dim[3];          // d, h, w
  int Cin = 0, Cout = 0;
  bool transposed = false; // source weight layout (Cin, Cout, k..) instead of (Cout, Cin, k..)
  bool lrelu = false;      // Conv -> LeakyReLU(0.1) -> InstanceNorm block?
  float* w = nullptr;      // device, kernel layout
  const float* bias = nullptr;
  const float* gamma = nullptr;  // InstanceNorm affine (null for bare convs)
  const float* beta = nullptr;
  int ntaps() const { return dim[0].ntaps() * dim[1].ntaps() * dim[2].ntaps(); }
  int nclass() const { return dim[0].nclass() * dim[1].nclass() * dim[2].nclass(); }
  size_t weight_elems() const { return (size_t)nclass() * ntaps() * Cin * Cout; }
  size_t weight_elems_pytorch() const {
    return (size_t)Cin * Cout * dim[0].ksize() * dim[1].ksize() * dim[2].ksize();
  }
};

inline ConvLayer make_layer(int cin, int cout, DimSpec d, DimSpec h, DimSpec w, bool transposed,
                            bool lrelu) {
  ConvLayer l;
  l.dim[0] = d; l.dim[1] = h; l.dim[2] = w;
  l.Cin = cin; l.Cout = cout; l.transposed = transposed; l.lrelu = lrelu;
  return l;
}

// Geometry of one launch.
struct ConvGeom {
  int N = 0;      // output samples
  int n_div = 1;  // input sample = n / n_div; disparity shift = n % n_div (second source only)
  int D = 1, H = 1, W = 1;  // input spatial size
};

// src: PyTorch-layout weights; dst: kernel layout (see ConvLayer::w).
int relayout_weights(const ConvLayer& l, const float* src, float* dst, cudaStream_t st);

// fp32 CUDA-core implicit GEMM.  in / in2 / out are channels-last.  in2 (may be
// null) supplies channels [Cin1, Cin) and is read at x - (n % n_div) with zero
// fill (Matching's shifted right descriptor, matching.py:56-60).  If stats is
// non-null the epilogue accumulates sum / sum of squares of the activated
// output per (n, cout) into stats[n][cout][2] (double), which must be zeroed.
int conv_forward_simt(const ConvLayer& l, const ConvGeom& g, const float* in, const float* in2,
                      int Cin1, float* out, double* stats, cudaStream_t st);

// Direct fp32 kernels for the few-channel 3x3x3 stride-1 layers (conv3d_direct.cu); *handled is
// false when the layer is not one of them (nothing is launched then).
int conv_forward_direct(const ConvLayer& l, const ConvGeom& g, const float* in, float* out,
                        double* stats, cudaStream_t st, bool* handled);

// InstanceNorm (biased variance, eps 1e-5) applied from accumulated statistics,
// fused with the additions that follow it in the reference:
//   norm = (y - mean) * rstd * gamma + beta
//   out  = norm                       (may alias y)
//   out2 = norm + add + add_bcast     (optional; add_bcast is [N][HW][C], broadcast over D)
// y / out / add: [N][S][C] channels-last, S = voxels per sample.
int instance_norm_apply(const float* y, const double* stats, const float* gamma, const float* beta,
                        const float* add, const float* add_bcast, float* out, float* out2, int N,
                        size_t S, size_t HW, int C, cudaStream_t st);

// Hourglass tail (hourglass_tail.cu): InstanceNorm of the 4-channel half-size volume
// fused into ConvTranspose3d(4 -> 1, (3,4,4), stride (1,2,2), padding 1).  in:
// [B][D][H][W][4] post-LeakyReLU; stats [B][4][2] (null: already normalised);
// gamma / beta / w (PyTorch layout) are HOST pointers; out: (B, D, 2H, 2W).  With `disparity`
// the layer is fused with SubpixelMap (estimator.py:59-91, window radius R, disparity step) and
// the SizeAdapter crop: the cost volume is never written, `out` is ignored.
int hourglass_tail_forward(const float* in, float* out, const double* stats, const float* gamma_host,
                           const float* beta_host, const float* w_host, float bias, int B, int D,
                           int H, int W, cudaStream_t st, float* disparity = nullptr,
                           int64_t* argmax = nullptr, int R = 0, int step = 1, int crop_top = 0,
                           int crop_left = 0, float* state = nullptr);
// Scratch for the fused form (partial SubpixelMap states of the disparity-axis segments); with it
// the fused kernel keeps the segmentation of the plain one and a merge kernel finishes.
size_t hourglass_tail_state_bytes(int B, int D, int H, int W);

// [N][C][S] <-> [N][S][C]
int nchw_to_nhwc(const float* in, float* out, int N, int C, size_t S, cudaStream_t st);
int nhwc_to_nchw(const float* in, float* out, int N, int C, size_t S, cudaStream_t st);

// Bump allocator over a caller-provided workspace (256-byte aligned blocks).
struct Workspace {
  char* base;
  size_t size, used = 0;
  bool overflow = false;
  Workspace(void* p, size_t n) : base((char*)p), size(n) {}
  template <typename T>
  T* take(size_t count) {
    const size_t bytes = align_up(count * sizeof(T), 256);
    if (used + bytes > size) { overflow = true; return nullptr; }
    T* r = (T*)(base + used);
    used += bytes;
    return r;
  }
};

}  // namespace pds
