// Generic tcgen05 implicit-GEMM convolution engine ("tcg") for the layers of the
// 3-D hourglass (reference regularization.py:74-126, network_blocks.py:61-85,
// 106-131) and of the embedding tower (embedding.py:14-65): Conv k3 stride 1|2,
// Conv k5 stride 2, ConvTranspose k4 stride 2 padding 1, in 2-D or 3-D, any
// channel count that is a multiple of 8.
//
// Same operand scheme as conv_tc.cu (split 16-bit terms, one TMEM accumulator
// per order of magnitude, activation planes "AP" whose TMA boxes land in shared
// memory as ready-made K-major UMMA operands, taps == descriptor start-address
// offsets), but the geometry is not compiled in: the HOST plans, per layer, a
// small *program* that one persistent kernel interprets:
//
//   work item  = one CTA tile (NTX x NTZ MMA tiles of 8 x 16 grid voxels) of one
//                sample, times one output parity class for transposed layers;
//   unit       = what one pipeline stage holds: a few TMA boxes of the input
//                (all of equal shape) and, unless the layer's weights are
//                resident, the weight slab of exactly those MMAs;
//   entry      = one K=16 MMA step: A start offset inside the stage, LBO
//                (distance to the second 8-channel K half: the next channel
//                plane, or -- for 8-channel layers -- ANOTHER TAP), B offset.
//
// Stride-2 layers read a *phase-separated* input (the 2 x 2 (x 2) parity
// sub-volumes stored one after the other, written that way by the normalisation
// pass that produces the tensor), so every box is dense and a tap is again a
// plain offset.  The planner is pure host code (no CUDA calls): tests emulate
// its programs on the CPU against the oracle (tests/test_tcg_plan.py).
#pragma once
#include <cuda.h>

#include <vector>

#include "pds_common.cuh"

namespace pds {

// TCG_TCONV4_S2M: the transposed layer with its 8 output parity classes MERGED along N: one
// 3x3x3-tap convolution over the input grid with 8 * Cout output columns (column = class * Cout +
// channel; a tap a class does not use has zero weights).  An MMA with N <= 64 is operand-fetch
// bound, so the wider N is free and the 27 merged taps replace 8 x 8 class taps: chosen by the
// callers for Cout <= 8.
// TCG_CONV3_S1X4: the 8-channel 3x3(x3) layer with FOUR neighbouring output voxels of a row on the N
// axis (column = x-position * Cout + channel).  A GEMM row is a group of four voxels; the input is
// stored in four x-phase sub-volumes (x mod 4), so the six input columns a group needs (x-1 .. x+4)
// are dense boxes and a tap is again a plain offset: 3x3x6 = 54 taps (27 paired MMAs) per 4 outputs
// instead of 27 taps (14 MMAs) per output.  An N <= 32 MMA is operand-fetch bound (~47 cycles
// whatever N is), so the wider N is free: 2x fewer tensor-core cycles per voxel.  The channels-last
// output [x/4][4][Cout] is byte-identical to [x][Cout]: nothing downstream changes.
enum TcgKind { TCG_CONV3_S1 = 0, TCG_CONV3_S2 = 1, TCG_TCONV4_S2 = 2, TCG_CONV5_S2 = 3, TCG_TCONV4_S2M = 4,
               TCG_CONV3_S1X4 = 5 };
// phase argument of tcg_norm_to_ap / tc_pack_nchw for an S1X4 consumer (x mod 4 sub-volumes)
constexpr int TCG_PHASES_X4 = -4;

struct TcgShape {
  int kind = TCG_CONV3_S1;
  int nd = 3;                 // spatial dimensions: 2 (Z == 1) or 3
  int Cin = 0, Cout = 0;
  int Z = 1, Y = 1, X = 1;    // INPUT extent (of the un-separated tensor for stride 2)
  int S = 2;                  // 16-bit terms per value
};

struct TcgBox {     // one TMA box of a unit
  int dx, dy, dz;   // box origin relative to the tile origin (grid coordinates)
  int plane;        // first plane, relative to (sample, term): phase * P + p
};

struct TcgUnit {
  int ent_beg, ent_end;     // entries [beg, end) in TcgPlan::entries
  int box_beg, box_end;     // boxes
  unsigned w_off16;         // weight slab of this unit: offset from the layer's weights, 16-byte units
  unsigned w_bytes;         //   and size (streamed layers copy exactly this per stage)
};

struct TcgEntry {
  unsigned a;       // A start offset inside the stage (16-byte units) | LBO (16-byte units) << 16
  unsigned b;       // B start offset (16-byte units) relative to the unit's slab (streamed) or the layer (resident)
};

// Where the two K halves of an entry come from in the PyTorch weight tensor:
// kernel index (kz, ky, kx) and 8-channel input group; group < 0 -> zeros.
struct TcgWeightSrc { int kz[2], ky[2], kx[2], group[2]; };

struct TcgPlan {
  TcgShape shape;
  int N = 16;                 // accumulator columns per weight term (Cout, or 8 * Cout when merged, padded to 16/32/64/128)
  int merged = 0;             // TCG_TCONV4_S2M
  int xg = 1;                 // output voxels per GEMM row (4: TCG_CONV3_S1X4)
  int nacc = 1, ntx = 1, ntz = 1;   // MMA tiles per CTA tile
  int ncls = 1;               // output parity classes (8 for the 3-D transposed layer)
  int nph = 1;                // phase sub-volumes of the input (1, 4 or 8; 4 x-phases for S1X4)
  int P = 1;                  // 8-channel planes per phase
  int GZ = 1, GY = 1, GX = 1; // grid the GEMM rows enumerate (per class; S1X4: X / 4 voxel groups)
  int OZ = 1, OY = 1, OX = 1; // output extent
  int IZ = 1, IY = 1, IX = 1; // extent of ONE input (phase) sub-volume == tensor-map dims
  int BX = 0, BY = 0, BZ = 0, PB = 1;  // box shape (pixels / rows / planes / channel planes)
  int units_per_item = 0;     // units of one work item (the same for every class)
  int max_boxes = 0;          // boxes of the largest unit: one term occupies max_boxes * box_bytes of a stage
  int resident = 0;           // weights stay in shared memory for the whole launch
  int phase_arg() const { return shape.kind == TCG_CONV3_S1X4 ? TCG_PHASES_X4 : nph; }   // what the producer of in_ap is told
  int stages = 0;
  unsigned box_bytes = 0, stage_bytes = 0, wres_bytes = 0, w_total_bytes = 0;
  std::vector<TcgUnit> units;         // [cls][unit]
  std::vector<TcgBox> boxes;
  std::vector<TcgEntry> entries;
  std::vector<TcgWeightSrc> wsrc;     // per entry
  std::vector<int> tile_off16;        // per MMA tile: A offset of the tile inside a box (16-byte units)
  size_t in_planes(int n_samples) const { return (size_t)n_samples * shape.S * nph * P; }
  size_t in_ap_bytes(int n_samples) const { return in_planes(n_samples) * IZ * IY * IX * 16; }
  size_t out_elems(int n_samples) const { return (size_t)n_samples * OZ * OY * OX * shape.Cout; }
};

// Plans a layer.  Returns PDS_ERR_UNSUPPORTED (message set) for shapes the engine
// does not serve (odd extents under stride 2, channel counts not a multiple of 8...).
int tcg_plan(const TcgShape& shape, TcgPlan* plan);

// ---- device side -------------------------------------------------------------------------
struct TcgLayer {
  TcgPlan plan;
  int transposed = 0;          // PyTorch weight layout (Cin, Cout, k...) instead of (Cout, Cin, k...)
  int cin_src = 0;             // input channels of the SOURCE weight tensor when fewer than the layer's (zero-padded)
  int fp16 = 1;
  float wscale = 256.f;
  // device buffers (owned by the caller's blob)
  uint16_t* w = nullptr;       // plan.w_total_bytes
  float* bias = nullptr;       // [N]
  void* prog = nullptr;        // units | boxes | entries | tile offsets, see conv_tcg.cu
  const float* gamma = nullptr;
  const float* beta = nullptr;
  size_t prog_bytes() const;
};

// Bytes of device memory a layer needs for w + bias + prog (each 256-aligned).
size_t tcg_layer_bytes(const TcgLayer& l);
// Carves the buffers out of `blob`, uploads the program and converts the weights
// (w_src / bias_src: PyTorch layout, device pointers).  Returns the bytes consumed.
int tcg_layer_init(TcgLayer& l, char* blob, const float* w_src, const float* bias_src,
                   cudaStream_t st, size_t* consumed);

// in_ap: AP planes [n][S][phase][P][IZ][IY][IX][8]; out: fp32 channels-last
// [n][OZ][OY][OX][Cout] after bias + LeakyReLU(0.1) (if lrelu); stats[n][Cout][2] (double,
// pre-zeroed) accumulate sum / sum of squares of the stored values when non-null.
// splitk_scratch (optional, tcg_splitk_bytes(l, n_samples) bytes): layers whose few tiles would leave
// most of the GPU idle (the deep hourglass levels) are split along K over several CTAs that store raw
// partial sums there; a fix-up kernel adds them in a fixed order, then bias, activation and sums.
int tcg_conv_forward(const TcgLayer& l, int n_samples, const uint16_t* in_ap, float* out,
                     double* stats, int lrelu, cudaStream_t st, float* splitk_scratch = nullptr,
                     size_t splitk_bytes = 0);
size_t tcg_splitk_bytes(const TcgLayer& l, int n_samples);

// One affine source of a normalisation pass: y fp32 channels-last with its
// InstanceNorm sums (null stats: taken as is).
struct TcgNormSrc {
  const float* y = nullptr;
  const double* stats = nullptr;   // [n][C][2]
  const float* gamma = nullptr;
  const float* beta = nullptr;
};
// out_ap = IN(a) [+ IN(b)] [+ bcast] as split AP planes; bcast: fp32 channels-last
// [n][Y][X][C] added to every z.  phases: 1 (plain), 4 or 8 (phase-separated for a stride-2
// consumer), or TCG_PHASES_X4.
// out_f32 (optional): the same sum as fp32 channels-last (a later pass adds it as a residual).
int tcg_norm_to_ap(const TcgNormSrc& a, const TcgNormSrc* b, const float* bcast, uint16_t* out_ap,
                   int n, int C, int Z, int Y, int X, int S, int fp16, int phases, cudaStream_t st,
                   float* out_f32 = nullptr);

bool tcg_available();   // driver entry point for cuTensorMapEncodeTiled resolved?

}  // namespace pds
