// tcgen05 (5th-gen tensor core) implicit-GEMM 3x3 convolution for the matching
// operation, with split-bf16 operands ("bf16xS": every fp32 value is carried as
// S bf16 terms, hi + mid + lo; the products of the leading terms are
// accumulated in fp32 in TMEM -- S=3 gives fp32-equivalent products, S=1 is
// plain bf16).  See conv_tc.cu for the kernel and DESIGN.md for the data layout.
#pragma once
#include <cuda.h>

#include "pds_common.cuh"

namespace pds {

// Activation planes ("AP"): [slice][split s][C/8][H][W][8] bf16 -- every
// 8-channel group of every pixel is one 16-byte vector, pixels of an image row
// are contiguous.  This is exactly the K-major no-swizzle core-matrix layout of
// a UMMA operand, so a TMA box (8 ch, x, y, planes) lands in shared memory ready
// for tcgen05.mma, and a tap shift is just a start-address offset.
struct TcLayer {
  int Cin = 0;     // input channels (multiple of 16)
  int Cout = 0;    // real output channels
  int N = 0;       // Cout padded to a multiple of 16 (UMMA N)
  int S = 1;       // split terms
  __nv_bfloat16* w = nullptr;   // [s][chunk][tap 9][2][N][8]
  float* bias = nullptr;        // [N]
  const float* gamma = nullptr; // InstanceNorm affine of the block (may be null)
  const float* beta = nullptr;
  size_t w_elems() const { return (size_t)S * (Cin / 16) * 9 * 2 * N * 8; }
};

enum TcEpilogue {
  TC_EPI_ACT = 0,   // bias + LeakyReLU + InstanceNorm sums -> fp32 planes
  TC_EPI_PLAIN = 1, // bias -> fp32 planes + split bf16 planes (conv0)
  TC_EPI_SIG = 2    // bias -> (B, Cout, D, H, W) fp32 signatures (last conv)
};

struct TcConvArgs {
  const TcLayer* layer;
  int epilogue;
  int n_slices;            // output slices (B * D)
  int n_div;               // conv0: input slice = n / n_div, disparity = n % n_div; else 1
  int H, W;
  // inputs: AP tensors; `in2` only for conv0 (right descriptors, read at x - d)
  const __nv_bfloat16* in;
  int in_slices;           // number of slices in `in`
  int in_C;                // channels of `in`
  const __nv_bfloat16* in2;
  int in2_C;
  // outputs (by epilogue)
  float* out_f32;          // [n][N/4][H][W][4]
  __nv_bfloat16* out_ap;   // [n][S][N/8][H][W][8]
  float* out_sig;          // (B, Cout, D, H, W)
  double* stats;           // [n][N][2], pre-zeroed
  // scratch: device array of CUtensorMap (>= 1 + n_div entries), 64-byte aligned
  CUtensorMap* maps_dev;
  CUtensorMap* maps_host;  // host staging of the same size
};

size_t tc_conv_max_maps(int n_div);

// weights (Cout, Cin, 3, 3) fp32 + bias (Cout) -> TcLayer buffers (pre-allocated)
int tc_prepare_weights(const TcLayer& l, const float* w_oihw, const float* bias, cudaStream_t st);

// (B, C, H, W) fp32 -> AP [B][S][C/8][H][W][8]
int tc_pack_nchw(const float* in, __nv_bfloat16* ap, int B, int C, int H, int W, int S,
                 cudaStream_t st);

int tc_conv3x3(const TcConvArgs& a, cudaStream_t st);

// InstanceNorm apply on fp32 planes [n][C/4][H][W][4] (+ optional residual, same
// layout), writing fp32 planes (optional) and split AP planes (optional).
int tc_norm_split(const float* y, const double* stats, const float* gamma, const float* beta,
                  const float* residual, float* out_f32, __nv_bfloat16* out_ap, int n_slices,
                  int C, int H, int W, int S, cudaStream_t st);

bool tc_available();  // driver entry point for cuTensorMapEncodeTiled resolved?

}  // namespace pds
