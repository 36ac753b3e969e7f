// tcgen05 (5th-gen tensor core) implicit-GEMM 3x3 convolution with split
// 16-bit operands: every fp32 value x is carried as S terms t0 + t1 (+ t2) of
// fp16 (11-bit significands, "fp16x2" = 22 bits) or bf16 (8-bit significands,
// "bf16x3" = 24 bits); the products whose orders of magnitude sum to < S are
// accumulated in fp32 in TMEM, ONE ACCUMULATOR PER ORDER OF MAGNITUDE, and added
// in fp32 registers in the epilogue.  See conv_tc.cu for the kernel and
// DESIGN.md for the data layout and the measured tensor-pipe model behind it.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "pds_common.cuh"

namespace pds {

// Activation planes ("AP"): [slice][term s][C/8][H][W][8] 16-bit -- every
// 8-channel group of every pixel is one 16-byte vector and the pixels of an
// image row are contiguous.  This is exactly the K-major no-swizzle core-matrix
// layout of a UMMA operand: a TMA box (8 ch, x, y, 2 planes) lands in shared
// memory ready for tcgen05.mma, and a convolution tap is a start-address offset.
struct TcLayer {
  int Cin = 0;       // input channels (multiple of 16)
  int Cout = 0;      // real output channels
  int N = 0;         // Cout padded to 16 or 64: rows of ONE term of the B operand
  int S = 1;         // terms per value
  int fp16 = 0;      // term type: 1 = IEEE half, 0 = bfloat16
  float wscale = 1;  // weights are stored multiplied by this power of two
  uint16_t* w = nullptr;        // [chunk][tap 9][K-half 2][term][N][8]
  float* bias = nullptr;        // [N]
  const float* gamma = nullptr; // InstanceNorm affine of the block (may be null)
  const float* beta = nullptr;
  size_t w_elems() const { return (size_t)(Cin / 16) * 9 * 2 * S * N * 8; }
};

enum TcEpilogue {
  TC_EPI_ACT = 0,   // bias + LeakyReLU + InstanceNorm sums -> fp32 planes
  TC_EPI_PLAIN = 1, // bias -> split AP planes (first convolution, no activation)
  TC_EPI_SIG = 2,   // bias -> (B, Cout, D, H, W) fp32 signatures (last convolution)
  TC_EPI_F32 = 3    // bias -> fp32 planes, no activation, no sums (factorised first convolution)
};

// InstanceNorm (+ residual) that follows a TC_EPI_ACT convolution, run by tc_conv3x3 either inside
// the convolution's launch (trailing normalisation warps, conv_tc.cu) or as a separate pass.
enum TcNorm {
  TC_NORM_NONE = 0,
  TC_NORM_PLAIN = 1,           // norm_out = IN(t)
  TC_NORM_RESIDUAL = 2,        // norm_out = IN(t) + res_ap (planes; norm_out may alias res_ap)
  TC_NORM_RESIDUAL_FIRST = 3   // norm_out = IN(t) + x0 rebuilt from fA / fB / fQ (matching_first.cuh)
};

struct TcConvArgs {
  const TcLayer* layer;
  int epilogue;
  int n_slices;            // output slices of THIS launch
  int n0;                  // global index (b * D + d) of the first of them
  int n_div;               // D: disparity = global slice % n_div, sample = global slice / n_div
  int H, W;
  // inputs: AP tensors.  With `in2` (first convolution) `in` holds one slice per
  // SAMPLE (left descriptors) and `in2` the right descriptors, read at x - d;
  // otherwise `in` holds the n_slices slices of this launch.
  const uint16_t* in;
  int in_slices;           // slices stored in `in` (and `in2`)
  int in_C;
  const uint16_t* in2;
  int in2_C;
  // outputs (by epilogue), indexed by the LOCAL slice except out_sig / stats
  float* out_f32;          // [n][N/4][H][W][4]
  uint16_t* out_ap;        // [n][S][N/8][H][W][8]
  float* out_sig;          // (B, Cout, D, H, W), global indexing
  double* stats;           // [global n][N][2], pre-zeroed
  // scratch: device array of CUtensorMap (>= 1 + n_div entries, 64-byte aligned) + host staging
  CUtensorMap* maps_dev;
  CUtensorMap* maps_host;
  // normalisation of this convolution's output with the layer's gamma / beta (TC_EPI_ACT only)
  int norm_mode;           // TcNorm
  uint16_t* norm_out;      // operand planes [n][S][N/8][H][W][8]
  const uint16_t* res_ap;  // TC_NORM_RESIDUAL
  const float* fA;         // TC_NORM_RESIDUAL_FIRST: per-sample fp32 planes [b][N/4][H][W][4]
  const float* fB;
  const float* fQ;
  void* sched;             // tc_sched_bytes(n_slices) bytes, ZEROED by the caller on the same stream (null: static kernel)
  int fuse;                // run the normalisation inside the convolution launch (needs sched)
};

// scratch (sum replicas + scheduler words) a fused convolution + normalisation launch needs
size_t tc_sched_bytes(int n_slices);
bool tc_fused_norm_enabled();
bool tc_dynamic_conv_enabled();

size_t tc_conv_max_maps(int n_div);

// weights (Cout, Cin, 3, 3) fp32 + bias (Cout) -> TcLayer buffers (pre-allocated)
// src_cin / ci_off: the layer reads input channels [ci_off, ci_off + l.Cin) of a source tensor
// with src_cin channels (0: the source has exactly l.Cin); bias may be null (zeros); qmode: see
// tc_compose_first.
int tc_prepare_weights(const TcLayer& l, const float* w_oihw, const float* bias, cudaStream_t st,
                       int src_cin = 0, int ci_off = 0, int qmode = 0);

// Factorised first convolution of the matching operation (conv_tc.cu, tc_compose_first_kernel):
// x0[b*D + d] = A[b] + shift_d(Bf[b]) + edge terms from Q[b], written as split AP planes.
int tc_compose_first(const float* A, const float* Bf, const float* Q, uint16_t* out_ap, int B, int C, int H,
                     int W, int D, int S, int fp16, cudaStream_t st);

// (B, C, H, W) fp32 -> AP [B][S][C/8][H][W][8]; x4_width > 0: rows of that width (H * W a multiple of
// it), written as the four x-phase sub-volumes a TCG_CONV3_S1X4 layer reads (conv_tcg.cuh)
int tc_pack_nchw(const float* in, uint16_t* ap, int B, int C, int H, int W, int S, int fp16,
                 cudaStream_t st, int x4_width = 0);

int tc_conv3x3(const TcConvArgs& a, cudaStream_t st);

// Host-side encoding of the D tensor maps the first convolution reads the right
// descriptors through (map d: width W - d, so the shifted image's right edge
// reads as zero padding).  The caller copies them to TcConvArgs::maps_dev.
int tc_encode_shift_maps(CUtensorMap* maps_host, const uint16_t* in2, int in_slices, int S, int C,
                         int H, int W, int D);

// ---- second-level factorisation (matching_factor.cu; derivation in its header) ---------------
// fp32 planes [n][C/4][H][W][4] -> split AP planes
int tc_planes_to_ap(const float* y, uint16_t* out_ap, int n, int C, int H, int W, int S, int fp16, cudaStream_t st);
// (Cout, Cin, 3, 3) fp32 -> [kx][dy][ci][co] fp32 for tc_column_ops
int tc_transpose_weights(const float* w_oihw, float* wt, int C, cudaStream_t st);
size_t tc_column_jobs(int D);   // J: columns per sample in `cols` ([b][J][H][C] fp32)
int tc_column_ops(const float* Bf, const float* Q, const float* wt, float* cols, int B, int C, int H, int W, int D,
                  cudaStream_t st);
// t[b*D + d] = LeakyReLU(conv1(x0_d) + bias) as fp32 planes + InstanceNorm sums (stats pre-zeroed);
// PA = conv1(A), PB = conv1(Bf) WITHOUT bias
int tc_compose_second(const float* PA, const float* PB, const float* cols, const float* bias, float* t,
                      double* stats, int B, int C, int H, int W, int D, cudaStream_t st);
// The same in two passes that never write the fp32 activation: sums only, then the values
// recomputed, normalised (InstanceNorm affine gamma / beta) and written as split operand planes
// out_ap [b*D + d][S][C/8][H][W][8] (bit-identical to tc_compose_second + tc_norm_split).
int tc_compose_second_stats(const float* PA, const float* PB, const float* cols, const float* bias, double* stats,
                            int B, int C, int H, int W, int D, cudaStream_t st);
int tc_compose_second_norm(const float* PA, const float* PB, const float* cols, const float* bias,
                           const double* stats, const float* gamma, const float* beta, uint16_t* out_ap, int B,
                           int C, int H, int W, int D, int S, int fp16, cudaStream_t st);
// out_ap[b*D + d] = IN(y) + x0_d with x0_d regenerated from A / Bf / Q
int tc_norm_residual_first(const float* y, const double* stats, const float* gamma, const float* beta,
                           const float* A, const float* Bf, const float* Q, uint16_t* out_ap, int B, int C, int H,
                           int W, int D, int S, int fp16, cudaStream_t st, int n_rep = 1, size_t rep_stride = 0);

// InstanceNorm apply on fp32 planes y [n][C/4][H][W][4] with the sums of
// stats[n][C][2]; the optional residual is READ FROM AP planes (sum of its
// terms) and the result is written as AP planes (out_ap may alias res_ap).
// `stats` may be n_rep private copies of the sums, rep_stride doubles apart (summed in the prologue).
int tc_norm_split(const float* y, const double* stats, const float* gamma, const float* beta,
                  const uint16_t* res_ap, uint16_t* out_ap, int n_slices, int C, int H, int W,
                  int S, int fp16, cudaStream_t st, int n_rep = 1, size_t rep_stride = 0);

// ---- last convolution (64 -> 8) with the taps on the M axis of the MMA tile (conv_last.cu) ----
size_t tc_last_weight_bytes();
bool tc_last_enabled();            // PDS_B200_LAST_TAPS=0: the generic kernel (N = 16 tile)
int tc_last_prepare(const float* w_oihw, uint16_t* packed, float wscale, cudaStream_t st);
// in: planes [n_slices][2 terms][8][H][W][8] fp16; out: (B, 8, D, H, W) fp32, slice n = b * n_div + d
int tc_conv_last(const uint16_t* packed_w, const float* bias, float wscale, const uint16_t* in, float* out,
                 int n_slices, int n_div, int H, int W, cudaStream_t st);

bool tc_available();  // driver entry point for cuTensorMapEncodeTiled resolved?

}  // namespace pds
