// Generic tcgen05 implicit-GEMM convolution kernel: interprets the programs planned by
// conv_tcg_plan.cu (see conv_tcg.cuh for the model).  sm_100a only.
//
// Warp roles (192 threads, one CTA per SM, persistent over a contiguous range of CTA tiles):
//   warp 0    TMA producer: per unit, the unit's boxes (cp.async.bulk.tensor.5d, zero fill ==
//             the convolution's padding) for every operand term, plus the unit's weight slab
//             when the layer's weights are not resident;
//   warp 1    MMA issuer (one elected lane): per entry, NACC x S tcgen05.mma -- activation term
//             s against the weight terms 0 .. S-1-s concatenated along N, one TMEM accumulator
//             per order of magnitude (conv_tc.cu explains the scheme);
//   warps 2-5 epilogue: TMEM -> registers, sum of the orders, bias, LeakyReLU, fp32
//             channels-last store, InstanceNorm sums (double atomics once per sample).
// Accumulators are double-buffered in TMEM whenever 2 * NACC * S * N <= 512 columns.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <set>
#include <string>

#include "conv_tcg.cuh"
#include "tc_ptx.cuh"

namespace pds {
namespace {

using namespace ptx;

constexpr int kEpilogueWarps = 8;
constexpr int kThreads = 32 * (2 + kEpilogueWarps);
constexpr int kMaxStages = 6;

struct alignas(64) TcgParams {
  CUtensorMap map;
  const unsigned char* tables;   // units | boxes | entries | tile offsets (device copy of the plan)
  const uint16_t* w;
  const float* bias;
  float* out;
  double* stats;
  int n_samples, lrelu, fp16;
  int nacc, ntx, ntz, ncls, upi;
  int GZ, GY, GX, OZ, OY, OX, Cout, mul, nd3;
  int CoutS, reps;               // channels of one stored row (xg * Cout); columns per channel (1, xg or 8 classes)
  // MMA issue constants, computed on the host so that they reach the issuing lane through the uniform
  // datapath (kernel parameters) instead of ordinary registers + R2UR
  uint32_t toff[8][3];           // A start offset of (MMA tile, term), 16-byte units
  uint32_t idesc[3], idesc_rest[3], a_hi, b_hi, b_lbo;
  int tiles_x, tiles_y, tiles_z;
  int BX, planes_per_term, ntx_log2, zstride16;
  int resident, stages, reuse, merged, flat;
  int ksplit;                    // > 1: an item covers 1/ksplit of the units and stores a RAW partial sum
  size_t ksplit_stride;          //      into out + ks * ksplit_stride (tcg_splitk_finish adds them up)
  uint32_t box_bytes, box_tx_bytes, term_bytes, a_bytes, stage_bytes, wres_bytes, w_total_bytes;
  uint32_t off_boxes, off_entries, off_tiles, table_bytes;
  float inv_wscale;
  long long* trace;              // optional per-CTA cycle stamps (debug): [cta][12]
  int dbg;                       // timing experiments (PDS_B200_TCG_DEBUG, only with the trace): 1 no global stores, 2 no TMEM reads, 4 no TMA loads after the first round of stages, 16 no wait tallies
};

constexpr uint32_t pow2_cols(uint32_t c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

template <int S, int N>
__global__ void __launch_bounds__(kThreads, 1)
conv_tcg_kernel(const __grid_constant__ TcgParams p) {
  constexpr int CH = N < 32 ? N : 32;            // accumulator columns handled per epilogue step
  constexpr int NCH = N / CH;
  constexpr uint32_t ACC_COLS = S * N;
  constexpr int NACC_MAX = 128 / N > 0 ? 128 / N : 1;   // MMA tiles per CTA tile never exceed this (planner)

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  const uint32_t wres_base = smem_u32(smem);
  const uint32_t stage_base = wres_base + p.wres_bytes;
  unsigned char* tab = smem + p.wres_bytes + (size_t)p.stages * p.stage_bytes;
  const TcgUnit* units = (const TcgUnit*)tab;
  const TcgBox* boxes = (const TcgBox*)(tab + p.off_boxes);
  const TcgEntry* entries = (const TcgEntry*)(tab + p.off_entries);
  const int* tile_off = (const int*)(tab + p.off_tiles);
  unsigned char* tail = tab + p.table_bytes;
  const uint32_t bar_base = smem_u32(tail);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + 2 + a); };
  const uint32_t wfull_bar = bar_base + 8u * (2 * kMaxStages + 4);
  uint32_t* tmem_slot = (uint32_t*)(tail + 8 * (2 * kMaxStages + 5) + 8);
  float* sbias = (float*)(tail + 8 * (2 * kMaxStages + 5) + 32);
  double* sred = (double*)(tail + 8 * (2 * kMaxStages + 5) + 32 + 128 * sizeof(float));   // [Cout][2] sums of the CTA

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t buf_cols = (uint32_t)p.nacc * ACC_COLS;
  const int nbuf = 2 * buf_cols <= 512 ? 2 : 1;
  const uint32_t tmem_cols = pow2_cols(nbuf * buf_cols);

  if (threadIdx.x == 0) {
    pdl_trigger();
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 32 * kEpilogueWarps); }
    mbar_init(wfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (p.resident) {   // constant data: may be fetched while the previous kernel is still running
      mbar_expect_tx(wfull_bar, p.w_total_bytes);
      for (uint32_t o = 0; o < p.w_total_bytes; o += 32768) {
        const uint32_t n = min(32768u, p.w_total_bytes - o);
        bulk_load(wres_base + o, (const unsigned char*)p.w + o, n, wfull_bar);
      }
    }
  }
  for (uint32_t i = threadIdx.x; i < p.table_bytes / 4; i += kThreads)
    ((uint32_t*)tab)[i] = ((const uint32_t*)p.tables)[i];
  for (int i = threadIdx.x; i < N; i += kThreads) sbias[i] = p.ksplit > 1 ? 0.f : p.bias[i];   // split-K: raw partial sums
  for (int i = threadIdx.x; i < 256; i += kThreads) sred[i] = 0.0;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything below reads / writes tensors of the stream's earlier kernels
  const long long t_start = clock64();
  auto stamp = [&](int k) { if (p.trace) p.trace[(size_t)blockIdx.x * 12 + k] = clock64() - t_start; };
  auto tally = [&](int k, long long t) { if (p.trace && !(p.dbg & 16)) p.trace[(size_t)blockIdx.x * 12 + k] += t; };

  // work item = (CTA tile, output class); every CTA walks one contiguous range of items (whole
  // tiles when a stage is shared by the classes of a tile)
  const int tiles_per_sample = p.tiles_x * p.tiles_y * p.tiles_z;
  const int total_items = tiles_per_sample * p.n_samples * p.ncls * p.ksplit;
  int per_cta = (total_items + (int)gridDim.x - 1) / (int)gridDim.x;
  if (p.reuse) per_cta = (per_cta + p.ncls - 1) / p.ncls * p.ncls;
  const int item_begin = min((int)blockIdx.x * per_cta, total_items);
  const int item_end = min(item_begin + per_cta, total_items);
  struct Item { int n, tx, ty, tz, cls, u0, u1; };
  auto decode = [&](int item) {
    Item it;
    const int ks = item % p.ksplit;       // split-K: the slices of one tile are neighbouring items (run concurrently)
    item /= p.ksplit;
    it.u0 = ks * p.upi / p.ksplit; it.u1 = (ks + 1) * p.upi / p.ksplit;
    it.cls = item % p.ncls;
    const int tile = item / p.ncls;
    it.n = tile / tiles_per_sample;
    int r = tile - it.n * tiles_per_sample;
    it.tz = r / (p.tiles_x * p.tiles_y);
    r -= it.tz * p.tiles_x * p.tiles_y;
    it.ty = r / p.tiles_x;
    it.tx = r - it.ty * p.tiles_x;
    return it;
  };

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      stamp(0);
      for (int item = item_begin; item < item_end; ++item) {
        const Item it = decode(item);
        if (p.reuse && it.cls != 0) continue;       // the tile's stage is already resident
        const int x0 = it.tx * 8 * p.ntx, y0 = it.ty * 16, z0 = it.tz * p.ntz;
        for (int u = it.u0; u < it.u1; ++u) {
          const TcgUnit un = units[it.cls * p.upi + u];
          const int nb = un.box_end - un.box_beg;
          const long long tw = p.trace ? clock64() : 0;
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (p.trace) tally(10, clock64() - tw);
          if ((p.dbg & 4) && (phase || item != item_begin)) {      // timing experiment: stages keep their first contents
            mbar_arrive(full_bar(stage));
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(full_bar(stage), (uint32_t)nb * S * p.box_tx_bytes + (p.resident ? 0u : un.w_bytes));
          const uint32_t sa = stage_base + stage * p.stage_bytes;
          for (int j = 0; j < nb; ++j) {
            const TcgBox b = boxes[un.box_beg + j];
#pragma unroll
            for (int s = 0; s < S; ++s)
              if (p.flat)
                tma_load_5d(sa + s * p.term_bytes + j * p.box_bytes, &p.map, 8 * (x0 + b.dx), y0 + b.dy, z0 + b.dz,
                            (it.n * S + s) * p.planes_per_term + b.plane, 0, full_bar(stage));
              else
                tma_load_5d(sa + s * p.term_bytes + j * p.box_bytes, &p.map, 0, x0 + b.dx, y0 + b.dy,
                            z0 + b.dz, (it.n * S + s) * p.planes_per_term + b.plane, full_bar(stage));
          }
          if (!p.resident)
            bulk_load(sa + p.a_bytes, (const unsigned char*)p.w + (size_t)un.w_off16 * 16, un.w_bytes,
                      full_bar(stage));
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
      }
      stamp(1);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t ent_base = smem_u32(entries);
    // the TMEM base address comes out of shared memory, i.e. in an ordinary register; rebuilt bit by bit
    // from warp votes it is a value ptxas KNOWS to be warp-uniform (column bits only: the allocation
    // starts at lane 0), so accumulator addresses stay on the uniform datapath as well
    uint32_t tmem_u = 0;
#pragma unroll
    for (int k = 0; k < 10; ++k) tmem_u |= (__ballot_sync(0xffffffffu, (tmem_base >> k) & 1u) != 0u ? 1u : 0u) << k;
    if (p.resident) { mbar_wait(wfull_bar, 0); tc_fence_after(); }
    if (lane == 0) stamp(2);
    // The whole warp walks the (warp-uniform) loops and waits on the barriers; one elected lane issues.
    // Stamps of PDS_B200_TCG_TRACE show the bursts of a unit running at the tensor pipe's own rate with
    // the pipe idle in between (~540 cycles per unit, ~1.6 K per item: commit, barrier wait, unit table,
    // first entry).  Tried against those gaps: awaiting the next unit's operands before
    // the last entry of the current one (gaps 370 / 1.1 K cycles, layer times unchanged within noise) and
    // ONE elected lane running the whole role (shorter gaps, but ptxas then emits a slower issue
    // sequence: 84 instead of 76 us for the one-voxel 8 -> 8 layer).  Kept: issue constants from kernel
    // parameters, a warp-uniform accumulate flag, no integer divisions at an item boundary unless the
    // layer is split along K (12 instead of 26 SASS instructions per tile of an entry).
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    bool first_full = true;
    int n_regions = 0;      // debug trace: (start, end) stamps of the first 64 issue regions of CTA 0
    auto issue_entries = [&](const TcgUnit& un, int k0, int k1, uint32_t sa16, uint32_t sw16, uint32_t d_base, bool first_unit) {
      // entries come from shared memory one ahead of their use
      uint2 en = lds64(ent_base + 8u * (un.ent_beg + k0));
      const int n_ent = un.ent_end - un.ent_beg;
      for (int k = k0; k < k1; ++k) {     // k and first_unit are warp-uniform: so is the accumulate flag (no R2UR / vote per MMA)
        const uint2 nxt = lds64(ent_base + 8u * (un.ent_beg + min(k + 1, n_ent - 1)));
        const uint32_t a_lo = en.x + sa16;
        const uint64_t bd = umma_desc((en.y + sw16) | p.b_lbo, p.b_hi);
        const uint32_t accumulate = (first_unit && k == 0) ? 0u : 1u;
#pragma unroll
        for (int i = 0; i < NACC_MAX; ++i) {
          if (i < p.nacc) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const uint64_t ad = umma_desc(a_lo + p.toff[i][s], p.a_hi);
              tc_mma(d_base + i * ACC_COLS + s * N, ad, bd, p.idesc[s], s == 0 ? accumulate : 1u);
              if (N * (S - s) > 256)   // B rows 256.. (16 bytes each), accumulator columns 256..
                tc_mma(d_base + i * ACC_COLS + s * N + 256, ad, bd + 256, p.idesc_rest[s],
                       s == 0 ? accumulate : 1u);
            }
          }
        }
        en = nxt;
      }
    };
    for (int item = item_begin; item < item_end; ++item) {
      int cls = 0, u0 = 0, u1 = p.upi;
      if (p.ksplit > 1) {
        cls = (item / p.ksplit) % p.ncls;
        u0 = (item % p.ksplit) * p.upi / p.ksplit; u1 = (item % p.ksplit + 1) * p.upi / p.ksplit;
      } else if (p.ncls > 1) {
        cls = item % p.ncls;
      }
      const long long tw0 = p.trace ? clock64() : 0;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      if (p.trace && lane == 0) tally(9, clock64() - tw0);
      tc_fence_after();
      const uint32_t d_base = tmem_u + acc * buf_cols;
      const bool hold = p.reuse && cls + 1 < p.ncls;      // the stage serves the next class too
      for (int u = u0; u < u1; ++u) {
        const TcgUnit un = units[cls * p.upi + u];
        if (!(p.reuse && cls > 0)) {
          const long long tw1 = p.trace ? clock64() : 0;
          mbar_wait(full_bar(stage), phase);
          if (p.trace && lane == 0) tally(8, clock64() - tw1);
          tc_fence_after();
          if (first_full && lane == 0) { stamp(3); first_full = false; }
        }
        const uint32_t sa16 = (stage_base + stage * p.stage_bytes) >> 4;
        const uint32_t sw16 = p.resident ? (wres_base >> 4) : sa16;
        const long long ti0 = p.trace ? clock64() : 0;
        if (p.trace && blockIdx.x == 0 && lane == 0 && n_regions < 64) p.trace[148 * 12 + 2 * n_regions] = ti0 - t_start;
        if (elect_one()) {
          issue_entries(un, 0, un.ent_end - un.ent_beg, sa16, sw16, d_base, u == u0);
          if (!hold) tc_commit(empty_bar(stage));
          if (u == u1 - 1) tc_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (p.trace && lane == 0) tally(11, clock64() - ti0);
        if (p.trace && blockIdx.x == 0 && lane == 0 && n_regions < 64) { p.trace[148 * 12 + 2 * n_regions + 1] = clock64() - t_start; ++n_regions; }
        if (!hold && ++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
      }
      if (nbuf == 2) { acc ^= 1; if (acc == 0) acc_phase ^= 1; } else { acc_phase ^= 1; }
    }
    if (lane == 0) stamp(4);
  } else {
    // ===== epilogue: 8 warps, two per TMEM lane quarter (32 * (warp % 4) .. + 31) =====
    // The two warps of a quarter split the MMA tiles of an item (or, with a single tile, its
    // column chunks).
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = 32 * q + lane;
    const int px = row & 7, py = row >> 3;
    uint32_t acc = 0, acc_phase = 0;
    double s1[NCH], s2[NCH];        // InstanceNorm sums of the current sample: channel c*CH + lane % CH
#pragma unroll
    for (int c = 0; c < NCH; ++c) { s1[c] = 0.0; s2[c] = 0.0; }
    int stat_n = -1;
    // The sums of a sample leave the CTA as ONE double atomic per (channel, moment): the eight
    // epilogue warps first combine in shared memory.  (Every CTA flushes at the same moment, at
    // the end of the launch; eight warps x 148 CTAs x 8 parity classes on the same few addresses
    // serialised in L2 for ~35 us in the merged transposed layer.)
    auto flush_stats = [&]() {
      if (stat_n >= 0 && p.stats) {
        if (lane < CH) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            const int col = c * CH + lane;
            const int ch = p.reps > 1 ? col % p.Cout : col;     // merged classes / voxel groups: several columns per channel
            if (col < p.reps * p.Cout && (s1[c] != 0.0 || s2[c] != 0.0)) {
              atomicAdd(&sred[2 * ch], s1[c]); atomicAdd(&sred[2 * ch + 1], s2[c]);
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpilogueWarps) : "memory");
        if (warp == 2) {
          for (int i = lane; i < 2 * p.Cout; i += 32) {
            const double v = sred[i];
            if (v != 0.0) atomicAdd(p.stats + (size_t)stat_n * p.Cout * 2 + i, v);
            sred[i] = 0.0;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpilogueWarps) : "memory");
      }
#pragma unroll
      for (int c = 0; c < NCH; ++c) { s1[c] = 0.0; s2[c] = 0.0; }
    };
    const bool split_tiles = p.nacc > 1;
    for (int item = item_begin; item < item_end; ++item) {
      const Item it = decode(item);
      if (it.n != stat_n) { flush_stats(); stat_n = it.n; }
      const int cz = (it.cls >> 2) & 1, cy = (it.cls >> 1) & 1, cx = it.cls & 1;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(32 * q) << 16) + acc * buf_cols;
      // per-thread partial sums of the item (NCH == 1: reduced across the warp once per item)
      float a1[CH], a2[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) { a1[j] = 0.f; a2[j] = 0.f; }
#pragma unroll 1
      for (int i = split_tiles ? half : 0; i < ((p.dbg & 2) ? 0 : p.nacc); i += split_tiles ? 2 : 1) {
        const int ix = i % p.ntx, iz = i / p.ntx;
        const int gx = it.tx * 8 * p.ntx + 8 * ix + px, gy = it.ty * 16 + py, gz = it.tz * p.ntz + iz;
        const bool valid = gx < p.GX && gy < p.GY && gz < p.GZ && !(p.dbg & 1);
        const int oz = p.nd3 ? p.mul * gz + cz : 0, oy = p.mul * gy + cy, ox = p.mul * gx + cx;
        float* o = p.out + (size_t)(item % p.ksplit) * p.ksplit_stride +
                   ((((size_t)it.n * p.OZ + oz) * p.OY + oy) * p.OX + ox) * p.CoutS;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (!split_tiles && NCH > 1 && (c >= NCH / 2) != (half != 0)) continue;
          float v[S][CH];
#pragma unroll
          for (int s = 0; s < S; ++s) tmem_ld_issue<CH>(t_base + i * ACC_COLS + s * N + c * CH, v[s]);
          tmem_ld_wait();
#pragma unroll
          for (int s = 0; s < S; ++s) tmem_ld_fence<CH>(v[s]);
#pragma unroll
          for (int s = S - 2; s >= 0; --s)       // smallest terms first
#pragma unroll
            for (int j = 0; j < CH; ++j) v[S - 1][j] += v[s][j];
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            float t = fmaf(v[S - 1][j], p.inv_wscale, sbias[c * CH + j]);
            if (p.lrelu) t = t > 0.f ? t : 0.1f * t;
            v[0][j] = valid ? t : 0.f;
          }
          if (valid) {
            if (p.merged && p.Cout == 4) {
              // two x-parity classes = 8 consecutive columns = 32 contiguous bytes of the output row
#pragma unroll
              for (int j = 0; j < CH; j += 8) {
                const int mc = (c * CH + j) >> 2;
                if (mc < 8)
                  stg256(p.out + ((((size_t)it.n * p.OZ + 2 * gz + ((mc >> 2) & 1)) * p.OY + 2 * gy + ((mc >> 1) & 1)) * p.OX +
                                  2 * gx) * 4, &v[0][j]);
              }
            } else if (p.merged) {
              // column = class * Cout + channel: scatter the 8 parity classes of this input voxel
#pragma unroll
              for (int j = 0; j < CH; j += 4) {
                const int col = c * CH + j;
                if (col < 8 * p.Cout) {
                  const int mc = col / p.Cout, co = col - mc * p.Cout;
                  float* om = p.out + ((((size_t)it.n * p.OZ + 2 * gz + ((mc >> 2) & 1)) * p.OY + 2 * gy + ((mc >> 1) & 1)) * p.OX +
                                       2 * gx + (mc & 1)) * p.Cout + co;
                  *reinterpret_cast<float4*>(om) = make_float4(v[0][j], v[0][j + 1], v[0][j + 2], v[0][j + 3]);
                }
              }
            } else if ((p.CoutS & 7) == 0) {
#pragma unroll
              for (int j = 0; j < CH; j += 8)
                if (c * CH + j < p.CoutS) stg256(o + c * CH + j, &v[0][j]);
            } else {
#pragma unroll
              for (int j = 0; j < CH; j += 4)
                if (c * CH + j < p.CoutS)
                  *reinterpret_cast<float4*>(o + c * CH + j) = make_float4(v[0][j], v[0][j + 1], v[0][j + 2], v[0][j + 3]);
            }
          }
          if (p.stats) {
            if (NCH == 1) {
#pragma unroll
              for (int j = 0; j < CH; ++j) { a1[j] += v[0][j]; a2[j] = fmaf(v[0][j], v[0][j], a2[j]); }
            } else {
              float sq[CH];
#pragma unroll
              for (int j = 0; j < CH; ++j) sq[j] = v[0][j] * v[0][j];
              s1[c] += (double)warp_transpose_reduce<CH>(v[0], lane);
              s2[c] += (double)warp_transpose_reduce<CH>(sq, lane);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (nbuf == 2) { acc ^= 1; if (acc == 0) acc_phase ^= 1; } else { acc_phase ^= 1; }
      if (NCH == 1 && p.stats) {
        s1[0] += (double)warp_transpose_reduce<CH>(a1, lane);
        s2[0] += (double)warp_transpose_reduce<CH>(a2, lane);
      }
    }
    if (warp == 2 && lane == 0) stamp(5);
    flush_stats();
    if (warp == 2 && lane == 0) stamp(6);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(7);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---- weights: PyTorch layout fp32 -> per entry [K half][term][N][8] 16-bit ------------------------
template <bool FP16>
__global__ void tcg_prepare_weights_kernel(const float* __restrict__ w, const TcgWeightSrc* __restrict__ src,
                                           uint16_t* __restrict__ out, int n_entries, int Cin, int Cout,
                                           int N, int S, int KZ, int KY, int KX, int transposed,
                                           float wscale, int merged, int xg) {   // Cin: channels of the SOURCE tensor (<= the layer's)
  const size_t per_entry = (size_t)2 * N * 8;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_entry * n_entries) return;
  const int e = (int)(i / per_entry);
  const int r = (int)(i % per_entry);
  const int c = r % 8, co = (r / 8) % N, h = r / (8 * N);
  const TcgWeightSrc ws = src[e];
  float x = 0.f;
  if (merged) {
    // column = class * Cout + channel; tap offset o = k - 1 per dimension; class bit c uses input
    // offsets c (kernel index 1 - c) and c - 1 (kernel index 3 - c) -- ConvTranspose k4 s2 p1
    const int mc = co / Cout, ch = co - mc * Cout;
    const int cb[3] = {(mc >> 2) & 1, (mc >> 1) & 1, mc & 1};
    const int off[3] = {ws.kz[h] - 1, ws.ky[h] - 1, ws.kx[h] - 1};
    int k[3];
    bool live = ws.group[h] >= 0 && mc < 8 && 8 * ws.group[h] + c < Cin;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int t = cb[d] - off[d];
      live = live && (t == 0 || t == 1);
      k[d] = 1 - cb[d] + 2 * t;
    }
    if (live) {
      const int ci = 8 * ws.group[h] + c;
      x = w[((((size_t)ci * Cout + ch) * 4 + k[0]) * 4 + k[1]) * 4 + k[2]] * wscale;
    }
  } else if (xg > 1) {
    // column = output position j of the voxel group * Cout + channel; the extended x tap k' reaches
    // output j through kernel index k' - j
    const int j = co / Cout, ch = co - j * Cout, kx = ws.kx[h] - j;
    if (ws.group[h] >= 0 && j < xg && kx >= 0 && kx < KX && 8 * ws.group[h] + c < Cin) {
      const int ci = 8 * ws.group[h] + c;
      x = w[((((size_t)ch * Cin + ci) * KZ + ws.kz[h]) * KY + ws.ky[h]) * KX + kx] * wscale;
    }
  } else if (ws.group[h] >= 0 && co < Cout && 8 * ws.group[h] + c < Cin) {
    const int ci = 8 * ws.group[h] + c;
    const size_t a = transposed ? ((size_t)ci * Cout + co) : ((size_t)co * Cin + ci);
    x = w[((a * KZ + ws.kz[h]) * KY + ws.ky[h]) * KX + ws.kx[h]] * wscale;
  }
  uint16_t t[3];
  split_terms<FP16>(x, t);
  for (int s = 0; s < S; ++s)
    out[(size_t)e * (2 * S * N * 8) + ((size_t)(h * S + s) * N + co) * 8 + c] = t[s];
}

__global__ void tcg_pad_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int Cout, int N,
                                    int reps) {   // reps columns per channel (merged classes / voxel groups)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = i < reps * Cout ? b[i % Cout] : 0.f;
}

// ---- normalisation pass: fp32 channels-last -> split AP planes ------------------------------------
struct NormParams {
  const float* ya; const double* sa; const float* ga; const float* ba;
  const float* yb; const double* sb; const float* gb; const float* bb;
  const float* bcast;
  uint16_t* out;
  float* out_f32;
  int C, Z, Y, X, S, phases;
};

template <bool FP16, int S, bool TWO>     // TWO: a second normalised source is added
__global__ void __launch_bounds__(256, TWO ? 2 : 3)
tcg_norm_to_ap_kernel(const NormParams p) {
  extern __shared__ float sm[];          // scale_a[C], shift_a[C], scale_b[C], shift_b[C]
  const int n = blockIdx.y;
  const size_t V = (size_t)p.Z * p.Y * p.X;
  if (threadIdx.x == 0) pdl_trigger();
  pdl_wait();
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    float sc = 1.f, sh = 0.f;
    if (p.sa) {
      const double s = p.sa[((size_t)n * p.C + c) * 2], q = p.sa[((size_t)n * p.C + c) * 2 + 1];
      const double mean = s / (double)V;
      double var = q / (double)V - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + 1e-5));
      sc = rstd * (p.ga ? p.ga[c] : 1.f);
      sh = (p.ba ? p.ba[c] : 0.f) - (float)mean * sc;
    }
    sm[c] = sc; sm[p.C + c] = sh;
    sc = 1.f; sh = 0.f;
    if (TWO && p.sb) {
      const double s = p.sb[((size_t)n * p.C + c) * 2], q = p.sb[((size_t)n * p.C + c) * 2 + 1];
      const double mean = s / (double)V;
      double var = q / (double)V - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + 1e-5));
      sc = rstd * (p.gb ? p.gb[c] : 1.f);
      sh = (p.bb ? p.bb[c] : 0.f) - (float)mean * sc;
    }
    sm[2 * p.C + c] = sc; sm[3 * p.C + c] = sh;
  }
  __syncthreads();
  // work unit = 256 consecutive (x, channel group) items of one image row (z, y); the CTAs walk
  // the units grid-stride, two at a time with both units' loads issued before the first store (the
  // per-CTA prologue above -- double-precision statistics -- is amortised over many units, and
  // enough bytes are in flight per SM).  No per-item division: G is a power of two.
  const int G = p.C / 8, gshift = 31 - __clz(G);
  const bool x4 = p.phases == TCG_PHASES_X4;      // four x-phase sub-volumes (S1X4 consumer)
  const int nph = x4 ? 4 : p.phases;
  const int sep = !x4 && p.phases > 1;
  const int sepz = p.phases == 8;
  const int IZ = sepz ? p.Z / 2 : p.Z, IY = sep ? p.Y / 2 : p.Y, IX = x4 ? p.X / 4 : (sep ? p.X / 2 : p.X);
  const size_t plane = (size_t)IZ * IY * IX;
  const int rows = p.Z * p.Y, items = p.X * G;
  const int chunks = (items + 255) >> 8, units = rows * chunks;
  struct Unit { int row, y, x, g; bool live; size_t off; float a[8], b[8]; float4 c0, c1; };
  auto load = [&](int u, Unit& w) {
    w.live = u < units;
    if (!w.live) return;
    w.row = u / chunks;
    const int i = ((u - w.row * chunks) << 8) + (int)threadIdx.x;
    w.live = i < items;
    if (!w.live) return;
    w.g = i & (G - 1); w.x = i >> gshift;
    w.y = w.row % p.Y;
    w.off = ((size_t)n * V + (size_t)w.row * p.X + w.x) * p.C + 8 * w.g;
    ldg256_stream(p.ya + w.off, w.a);              // 8 channels = one 32-byte sector per lane
    if (TWO) ldg256_stream(p.yb + w.off, w.b);
    if (p.bcast) {
      const float4* pc = reinterpret_cast<const float4*>(p.bcast + (((size_t)n * p.Y + w.y) * p.X + w.x) * p.C + 8 * w.g);
      w.c0 = __ldg(pc); w.c1 = __ldg(pc + 1);
    }
  };
  auto finish = [&](const Unit& w) {
    if (!w.live) return;
    const int z = w.row / p.Y, y = w.y, x = w.x, g = w.g;
    float v[8];
    const float* sa = sm + 8 * g;
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = fmaf(w.a[e], sa[e], sa[p.C + e]);
    if (TWO) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += fmaf(w.b[e], sa[2 * p.C + e], sa[3 * p.C + e]);
    }
    if (p.bcast) {
      v[0] += w.c0.x; v[1] += w.c0.y; v[2] += w.c0.z; v[3] += w.c0.w;
      v[4] += w.c1.x; v[5] += w.c1.y; v[6] += w.c1.z; v[7] += w.c1.w;
    }
    if (p.out_f32) stg256(p.out_f32 + w.off, v);
    uint16_t t[8][3];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_terms<FP16>(v[e], t[e]);
    const int zz = sepz ? z >> 1 : z, yy = sep ? y >> 1 : y;
    const int ph = x4 ? (x & 3) : (sep ? ((sepz ? (z & 1) * 4 : 0) + (y & 1) * 2) : 0) + (sep ? (x & 1) : 0);
    const size_t pos = ((size_t)zz * IY + yy) * IX + (x4 ? x >> 2 : (sep ? x >> 1 : x));
#pragma unroll
    for (int s = 0; s < S; ++s) {
      float4 pk;
      pk.x = __uint_as_float(t[0][s] | ((uint32_t)t[1][s] << 16)); pk.y = __uint_as_float(t[2][s] | ((uint32_t)t[3][s] << 16));
      pk.z = __uint_as_float(t[4][s] | ((uint32_t)t[5][s] << 16)); pk.w = __uint_as_float(t[6][s] | ((uint32_t)t[7][s] << 16));
      stg_stream(reinterpret_cast<float4*>(p.out) + ((((size_t)n * S + s) * nph + ph) * G + g) * plane + pos, pk);
    }
  };
  for (int u = blockIdx.x; u < units; u += 2 * gridDim.x) {
    Unit w0, w1;
    load(u, w0);
    load(u + gridDim.x, w1);
    finish(w0);
    finish(w1);
  }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult res;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &res) == cudaSuccess &&
        res == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

constexpr size_t kTailBytes = 8 * (2 * kMaxStages + 5) + 32 + 128 * sizeof(float) + 256 * sizeof(double) + 256;

struct TableLayout { uint32_t off_boxes, off_entries, off_tiles, table_bytes, off_wsrc, total; };

TableLayout table_layout(const TcgPlan& pl) {
  TableLayout t;
  uint32_t o = (uint32_t)(pl.units.size() * sizeof(TcgUnit));
  o = (uint32_t)align_up(o, 16); t.off_boxes = o; o += (uint32_t)(pl.boxes.size() * sizeof(TcgBox));
  o = (uint32_t)align_up(o, 16); t.off_entries = o; o += (uint32_t)(pl.entries.size() * sizeof(TcgEntry));
  o = (uint32_t)align_up(o, 16); t.off_tiles = o; o += (uint32_t)(pl.tile_off16.size() * 4);
  t.table_bytes = (uint32_t)align_up(o, 128);
  t.off_wsrc = t.table_bytes;
  t.total = t.off_wsrc + (uint32_t)(pl.wsrc.size() * sizeof(TcgWeightSrc));
  return t;
}

// Kernel-class name for the profiler; PDS_B200_PROFILE_DETAIL=1 appends the layer geometry so that
// bench.py / tools can attribute time to individual layers.
const char* tcg_name(int S, int N, const TcgPlan& pl) {
  static const bool detail = getenv("PDS_B200_PROFILE_DETAIL") && atoi(getenv("PDS_B200_PROFILE_DETAIL"));
  static std::mutex mu;
  static std::set<std::string> names;
  std::string n = "conv_tcg<S=" + std::to_string(S) + ",N=" + std::to_string(N) + ">";
  if (detail) {
    static const char* kinds[] = {"c3s1", "c3s2", "t4s2", "c5s2", "t4s2m", "c3s1x4"};
    n += std::string("[") + kinds[pl.shape.kind] + " " + std::to_string(pl.shape.Cin) + "->" +
         std::to_string(pl.shape.Cout) + " " + std::to_string(pl.shape.Z) + "x" + std::to_string(pl.shape.Y) +
         "x" + std::to_string(pl.shape.X) + "]";
  }
  std::lock_guard<std::mutex> lock(mu);
  return names.insert(n).first->c_str();
}

template <int S, int N>
int launch_tcg(const TcgParams& p, const TcgPlan& pl, size_t smem, int grid, cudaStream_t st, double flops,
               double bytes) {
  PDS_CUDA(allow_dynamic_smem(conv_tcg_kernel<S, N>, 227 * 1024));
  PDS_KERNEL(tcg_name(S, N, pl), st);
  PDS_KERNEL_WORK(flops, bytes);
  PDS_CUDA(launch_pdl(conv_tcg_kernel<S, N>, dim3(grid), dim3(kThreads), smem, st, p));
  return PDS_OK;
}

}  // namespace

bool tcg_available() { return encode_fn() != nullptr; }

size_t TcgLayer::prog_bytes() const { return table_layout(plan).total; }

size_t tcg_layer_bytes(const TcgLayer& l) {
  return align_up(l.plan.w_total_bytes, 256) + align_up((size_t)l.plan.N * 4, 256) + align_up(l.prog_bytes(), 256);
}

int tcg_layer_init(TcgLayer& l, char* blob, const float* w_src, const float* bias_src, cudaStream_t st,
                   size_t* consumed) {
  const TcgPlan& pl = l.plan;
  const TableLayout tl = table_layout(pl);
  char* cur = blob;
  l.w = (uint16_t*)cur; cur += align_up(pl.w_total_bytes, 256);
  l.bias = (float*)cur; cur += align_up((size_t)pl.N * 4, 256);
  l.prog = cur; cur += align_up(tl.total, 256);
  *consumed = (size_t)(cur - blob);
  std::vector<unsigned char> host(tl.total, 0);
  memcpy(host.data(), pl.units.data(), pl.units.size() * sizeof(TcgUnit));
  memcpy(host.data() + tl.off_boxes, pl.boxes.data(), pl.boxes.size() * sizeof(TcgBox));
  memcpy(host.data() + tl.off_entries, pl.entries.data(), pl.entries.size() * sizeof(TcgEntry));
  memcpy(host.data() + tl.off_tiles, pl.tile_off16.data(), pl.tile_off16.size() * 4);
  memcpy(host.data() + tl.off_wsrc, pl.wsrc.data(), pl.wsrc.size() * sizeof(TcgWeightSrc));
  PDS_CUDA(cudaMemcpyAsync(l.prog, host.data(), tl.total, cudaMemcpyHostToDevice, st));
  PDS_CUDA(cudaStreamSynchronize(st));   // `host` goes out of scope
  const int k = pl.shape.kind == TCG_TCONV4_S2 ? 4 : (pl.shape.kind == TCG_CONV5_S2 ? 5 : 3);
  const int KZ = pl.shape.nd == 3 ? k : 1;
  const size_t total = (size_t)2 * pl.N * 8 * pl.entries.size();
  const int cin_src = l.cin_src > 0 ? l.cin_src : pl.shape.Cin;
  {
    PDS_KERNEL("tcg_prepare_weights", st);
    const unsigned g = (unsigned)((total + 255) / 256);
    const TcgWeightSrc* src = (const TcgWeightSrc*)((const char*)l.prog + tl.off_wsrc);
    if (l.fp16)
      tcg_prepare_weights_kernel<true><<<g, 256, 0, st>>>(w_src, src, l.w, (int)pl.entries.size(), cin_src,
                                                          pl.shape.Cout, pl.N, pl.shape.S, KZ, k, k,
                                                          l.transposed, l.wscale, pl.merged, pl.xg);
    else
      tcg_prepare_weights_kernel<false><<<g, 256, 0, st>>>(w_src, src, l.w, (int)pl.entries.size(), cin_src,
                                                           pl.shape.Cout, pl.N, pl.shape.S, KZ, k, k,
                                                           l.transposed, l.wscale, pl.merged, pl.xg);
    PDS_LAUNCH_CHECK("tcg_prepare_weights_kernel");
  }
  PDS_KERNEL("tcg_pad_bias", st);
  tcg_pad_bias_kernel<<<1, 128, 0, st>>>(bias_src, l.bias, pl.shape.Cout, pl.N, pl.merged ? 8 : pl.xg);
  PDS_LAUNCH_CHECK("tcg_pad_bias_kernel");
  return PDS_OK;
}

long long* g_tcg_trace = nullptr;   // set by the debug hook (PDS_B200_TCG_TRACE)

namespace {

// Split-K fix-up: out = act(sum over the k slices of the raw partial sums + bias), plus the InstanceNorm
// sums of the stored values.  The slices are added in a fixed order (deterministic results).  Tensors are
// channels-last [rows][Cout]; a thread owns four channels and walks rows, the sums of a block leave it
// as one double atomic per (channel, moment).
__global__ void __launch_bounds__(256)
tcg_splitk_finish_kernel(const float* __restrict__ partials, int ksplit, size_t stride, const float* __restrict__ bias,
                         int lrelu, float* __restrict__ out, double* __restrict__ stats, size_t rows_per_sample,
                         int Cout) {
  extern __shared__ float fred[];      // [rows per iteration][2 * Cout]: per-thread sums, combined without atomics
  const int n = blockIdx.y, groups = Cout / 4;
  const int cg = threadIdx.x % groups, lane_row = threadIdx.x / groups, rows_per_iter = blockDim.x / groups;
  const float4 b4 = *reinterpret_cast<const float4*>(bias + 4 * cg);
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (lane_row < rows_per_iter)
    for (size_t r = (size_t)blockIdx.x * rows_per_iter + lane_row; r < rows_per_sample; r += (size_t)gridDim.x * rows_per_iter) {
      const size_t o = ((size_t)n * rows_per_sample + r) * Cout + 4 * cg;
      float4 a = __ldcg(reinterpret_cast<const float4*>(partials + o));
      for (int k = 1; k < ksplit; ++k) {
        const float4 q = __ldcg(reinterpret_cast<const float4*>(partials + (size_t)k * stride + o));
        a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
      }
      float v[4] = {a.x + b4.x, a.y + b4.y, a.z + b4.z, a.w + b4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (lrelu) v[j] = v[j] > 0.f ? v[j] : 0.1f * v[j];
        s1[j] += v[j]; s2[j] = fmaf(v[j], v[j], s2[j]);
      }
      *reinterpret_cast<float4*>(out + o) = make_float4(v[0], v[1], v[2], v[3]);
    }
  if (!stats) return;
  if (lane_row < rows_per_iter) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      fred[lane_row * 2 * Cout + 2 * (4 * cg + j)] = s1[j];
      fred[lane_row * 2 * Cout + 2 * (4 * cg + j) + 1] = s2[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * Cout; i += blockDim.x) {
    double t = 0.0;
    for (int r = 0; r < rows_per_iter; ++r) t += (double)fred[r * 2 * Cout + i];
    if (t != 0.0) atomicAdd(stats + (size_t)n * Cout * 2 + i, t);
  }
}

// k slices for a layer: only when its items leave most of the GPU idle (the deep hourglass levels:
// a handful of tiles, each streaming the whole weight tensor through ONE CTA), never for merged /
// stage-sharing transposed layers.  PDS_B200_TCG_SPLITK=0 disables.
int splitk_for(const TcgPlan& pl, int n_samples) {
  // PDS_B200_TCG_SPLITK: smallest k-slice count worth the fix-up launch (0: never split)
  static const int least = getenv("PDS_B200_TCG_SPLITK") ? atoi(getenv("PDS_B200_TCG_SPLITK")) : 4;
  if (least <= 0 || pl.merged || pl.xg > 1 || pl.units_per_item < 2 || pl.shape.Cout % 4) return 1;
  if (pl.ncls > 1 && pl.units_per_item == 1) return 1;
  const int tiles = ((pl.GX + 8 * pl.ntx - 1) / (8 * pl.ntx)) * ((pl.GY + 15) / 16) * ((pl.GZ + pl.ntz - 1) / pl.ntz);
  const int items = tiles * n_samples * pl.ncls;
  int k = num_sms() / (items > 0 ? items : 1);
  if (k > pl.units_per_item) k = pl.units_per_item;
  if (k > 8) k = 8;
  return k >= (least > 2 ? least : 2) ? k : 1;
}

}  // namespace

size_t tcg_splitk_bytes(const TcgLayer& l, int n_samples) {
  const int k = splitk_for(l.plan, n_samples);
  return k > 1 ? align_up((size_t)k * l.plan.out_elems(n_samples) * sizeof(float), 256) : 0;
}

int tcg_conv_forward(const TcgLayer& l, int n_samples, const uint16_t* in_ap, float* out, double* stats,
                     int lrelu, cudaStream_t st, float* splitk_scratch, size_t splitk_bytes) {
  const TcgPlan& pl = l.plan;
  if (n_samples == 0) return PDS_OK;
  if (!encode_fn()) {
    set_error("conv_tcg: cuTensorMapEncodeTiled is not available from this driver");
    return PDS_ERR_UNSUPPORTED;
  }
  const TableLayout tl = table_layout(pl);
  TcgParams p = {};
  {
    // pixel and channel group as ONE dimension when the box row fits 256 elements: a box row is a
    // single request of 16 * BX bytes instead of BX requests of 16 (same bytes in shared memory)
    const bool flat = 8 * pl.BX <= 256;
    const cuuint64_t planes = (cuuint64_t)pl.in_planes(n_samples), vol = (cuuint64_t)pl.IZ * pl.IY * pl.IX * 16;
    const cuuint64_t dims5[5] = {8, (cuuint64_t)pl.IX, (cuuint64_t)pl.IY, (cuuint64_t)pl.IZ, planes};
    const cuuint64_t strides5[4] = {16, (cuuint64_t)pl.IX * 16, (cuuint64_t)pl.IY * pl.IX * 16, vol};
    const cuuint32_t box5[5] = {8, (cuuint32_t)pl.BX, (cuuint32_t)pl.BY, (cuuint32_t)pl.BZ, (cuuint32_t)pl.PB};
    const cuuint64_t dimsf[5] = {(cuuint64_t)pl.IX * 8, (cuuint64_t)pl.IY, (cuuint64_t)pl.IZ, planes, 1};
    const cuuint64_t stridesf[4] = {(cuuint64_t)pl.IX * 16, (cuuint64_t)pl.IY * pl.IX * 16, vol, planes * vol};
    const cuuint32_t boxf[5] = {(cuuint32_t)pl.BX * 8, (cuuint32_t)pl.BY, (cuuint32_t)pl.BZ, (cuuint32_t)pl.PB, 1};
    const cuuint64_t* dims = flat ? dimsf : dims5;
    const cuuint64_t* strides = flat ? stridesf : strides5;
    const cuuint32_t* box = flat ? boxf : box5;
    p.flat = flat ? 1 : 0;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode_fn()(&p.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<uint16_t*>(in_ap), dims,
                             strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv_tcg: cuTensorMapEncodeTiled failed with CUresult %d (dims %d x %d x %d x %zu, box %d x %d x %d x %d, flat %d)",
                (int)r, pl.IX, pl.IY, pl.IZ, pl.in_planes(n_samples), pl.BX, pl.BY, pl.BZ, pl.PB, (int)flat);
      return PDS_ERR_CUDA;
    }
  }
  p.tables = (const unsigned char*)l.prog; p.w = l.w; p.bias = l.bias; p.out = out; p.stats = stats;
  p.n_samples = n_samples; p.lrelu = lrelu; p.fp16 = l.fp16;
  p.ksplit = 1; p.ksplit_stride = 0;
  int ksplit = splitk_for(pl, n_samples);
  if (ksplit > 1 && (!splitk_scratch || splitk_bytes < tcg_splitk_bytes(l, n_samples))) ksplit = 1;
  if (ksplit > 1) {     // raw partial sums per k slice; tcg_splitk_finish_kernel adds bias, activation, sums
    p.ksplit = ksplit; p.ksplit_stride = pl.out_elems(n_samples);
    p.out = splitk_scratch; p.stats = nullptr; p.lrelu = 0;
  }
  p.nacc = pl.nacc; p.ntx = pl.ntx; p.ntz = pl.ntz; p.ncls = pl.ncls; p.upi = pl.units_per_item;
  p.GZ = pl.GZ; p.GY = pl.GY; p.GX = pl.GX; p.OZ = pl.OZ; p.OY = pl.OY; p.OX = pl.OX / pl.xg;   // a stored row = xg voxels
  p.CoutS = pl.xg * pl.shape.Cout; p.reps = pl.merged ? 8 : pl.xg;
  p.Cout = pl.shape.Cout; p.mul = (pl.ncls > 1 || pl.merged) ? 2 : 1; p.nd3 = pl.shape.nd == 3;
  p.tiles_x = (pl.GX + 8 * pl.ntx - 1) / (8 * pl.ntx);
  p.tiles_y = (pl.GY + 15) / 16;
  p.tiles_z = (pl.GZ + pl.ntz - 1) / pl.ntz;
  p.BX = pl.BX; p.planes_per_term = pl.nph * pl.P;
  p.ntx_log2 = pl.ntx == 1 ? 0 : (pl.ntx == 2 ? 1 : (pl.ntx == 4 ? 2 : 3));
  p.zstride16 = pl.BY * pl.BX;
  if ((1 << p.ntx_log2) != pl.ntx || pl.nacc > (128 / pl.N > 0 ? 128 / pl.N : 1)) {
    set_error("conv_tcg: unsupported tile arrangement %d x %d", pl.ntx, pl.ntz);
    return PDS_ERR_UNSUPPORTED;
  }
  p.resident = pl.resident; p.stages = pl.stages; p.merged = pl.merged;
  p.reuse = pl.ncls > 1 && pl.units_per_item == 1 && pl.resident;
  p.box_tx_bytes = (uint32_t)pl.PB * pl.BZ * pl.BY * pl.BX * 16;   // what TMA delivers (box_bytes is its 128-aligned slot)
  p.box_bytes = pl.box_bytes; p.term_bytes = (uint32_t)pl.max_boxes * pl.box_bytes;
  p.a_bytes = p.term_bytes * pl.shape.S; p.stage_bytes = pl.stage_bytes; p.wres_bytes = pl.wres_bytes;
  p.w_total_bytes = pl.w_total_bytes;
  p.off_boxes = tl.off_boxes; p.off_entries = tl.off_entries; p.off_tiles = tl.off_tiles;
  p.table_bytes = tl.table_bytes;
  p.inv_wscale = 1.0f / l.wscale;
  {
    const uint32_t fmt = (1u << 4) | (l.fp16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(128 >> 4) << 24);
    const int S = pl.shape.S;
    for (int s = 0; s < 3; ++s) {
      // activation term s multiplies the N * (S - s) concatenated weight rows; an MMA is at most 256
      // columns wide, so the widest case (S = 3, N = 128) takes a second instruction for the rest
      const int cols = s < S ? pl.N * (S - s) : 0, first = cols > 256 ? 256 : cols;
      p.idesc[s] = fmt | ((uint32_t)(first >> 3) << 17);
      p.idesc_rest[s] = fmt | ((uint32_t)((cols - first) >> 3) << 17);
    }
    p.a_hi = (((uint32_t)pl.BX * 16 >> 4) & 0x3fff) | (1u << 14);      // umma_desc_hi(SBO = box row pitch)
    p.b_hi = ((128u >> 4) & 0x3fff) | (1u << 14);
    p.b_lbo = ((uint32_t)(S * pl.N) & 0x3fff) << 16;                     // K-half stride of the B operand, 16-byte units
    for (int i = 0; i < 8; ++i)
      for (int s = 0; s < 3; ++s)
        p.toff[i][s] = (uint32_t)((i >> p.ntx_log2) * p.zstride16 + ((i & (pl.ntx - 1)) << 3)) + s * (p.term_bytes >> 4);
  }
  p.trace = g_tcg_trace;
  p.dbg = (g_tcg_trace && getenv("PDS_B200_TCG_DEBUG")) ? atoi(getenv("PDS_B200_TCG_DEBUG")) : 0;
  if (pl.shape.Cout % 4) { set_error("conv_tcg: Cout must be a multiple of 4"); return PDS_ERR_UNSUPPORTED; }
  const size_t smem = (size_t)pl.wres_bytes + (size_t)pl.stages * pl.stage_bytes + tl.table_bytes + kTailBytes + 128;
  if (smem > 227 * 1024) {
    set_error("conv_tcg: plan needs %zu bytes of shared memory", smem);
    return PDS_ERR_UNSUPPORTED;
  }
  const int total = p.tiles_x * p.tiles_y * p.tiles_z * n_samples * (p.reuse ? 1 : pl.ncls) * p.ksplit;
  const int grid = total < num_sms() ? total : num_sms();
  const int k = (pl.shape.kind == TCG_TCONV4_S2 || pl.merged) ? 2 : (pl.shape.kind == TCG_CONV5_S2 ? 5 : 3);
  const double taps = (double)k * k * (pl.shape.nd == 3 ? k : 1);
  const double rows = (double)n_samples * pl.GZ * pl.GY * pl.GX * (pl.merged ? 8 : pl.ncls) * pl.xg;
  const double flops = 2.0 * taps * pl.shape.Cin * pl.shape.Cout * rows;
  const double bytes = (double)pl.in_ap_bytes(n_samples) + 4.0 * pl.out_elems(n_samples);
  int rc = PDS_ERR_UNSUPPORTED;
#define PDS_TCG_CASE(SS, NN) \
  if (pl.shape.S == SS && pl.N == NN) rc = launch_tcg<SS, NN>(p, pl, smem, grid, st, flops, bytes);
  PDS_TCG_CASE(2, 16) PDS_TCG_CASE(2, 32) PDS_TCG_CASE(2, 64) PDS_TCG_CASE(2, 128)
  PDS_TCG_CASE(3, 16) PDS_TCG_CASE(3, 32) PDS_TCG_CASE(3, 64) PDS_TCG_CASE(3, 128)
  PDS_TCG_CASE(1, 16) PDS_TCG_CASE(1, 32) PDS_TCG_CASE(1, 64) PDS_TCG_CASE(1, 128)
#undef PDS_TCG_CASE
  if (rc == PDS_ERR_UNSUPPORTED) set_error("conv_tcg: no kernel for S=%d N=%d", pl.shape.S, pl.N);
  if (rc != PDS_OK || ksplit == 1) return rc;
  {
    const size_t rows = (size_t)pl.OZ * pl.OY * pl.OX;
    const int Cout = pl.shape.Cout, rows_per_iter = 256 / (Cout / 4);
    unsigned gx = (unsigned)((rows + 4 * rows_per_iter - 1) / (4 * rows_per_iter));     // ~4 rows per thread
    if (gx > 296) gx = 296;
    if (gx < 1) gx = 1;
    PDS_KERNEL("tcg_splitk_finish", st);
    PDS_KERNEL_WORK(0, (double)(ksplit + 1) * pl.out_elems(n_samples) * 4.0);
    tcg_splitk_finish_kernel<<<dim3(gx, (unsigned)n_samples), 256, (size_t)rows_per_iter * 2 * Cout * sizeof(float), st>>>(
        splitk_scratch, ksplit, pl.out_elems(n_samples), l.bias, lrelu, out, stats, rows, Cout);
    PDS_LAUNCH_CHECK("tcg_splitk_finish_kernel");
  }
  return PDS_OK;
}

int tcg_norm_to_ap(const TcgNormSrc& a, const TcgNormSrc* b, const float* bcast, uint16_t* out_ap,
                   int n, int C, int Z, int Y, int X, int S, int fp16, int phases, cudaStream_t st,
                   float* out_f32) {
  if (n == 0) return PDS_OK;
  if (C % 8 || C > 128 || (phases != 1 && phases != 4 && phases != 8 && phases != TCG_PHASES_X4) ||
      (phases > 1 && ((X | Y) & 1)) || (phases == 8 && (Z & 1)) || (phases == TCG_PHASES_X4 && (X & 3))) {
    set_error("tcg_norm_to_ap: unsupported shape (C=%d, %d x %d x %d, phases %d)", C, Z, Y, X, phases);
    return PDS_ERR_UNSUPPORTED;
  }
  NormParams p;
  p.ya = a.y; p.sa = a.stats; p.ga = a.gamma; p.ba = a.beta;
  p.yb = b ? b->y : nullptr; p.sb = b ? b->stats : nullptr; p.gb = b ? b->gamma : nullptr; p.bb = b ? b->beta : nullptr;
  p.bcast = bcast; p.out = out_ap; p.out_f32 = out_f32;
  p.C = C; p.Z = Z; p.Y = Y; p.X = X; p.S = S; p.phases = phases;
  if (C & (C - 1)) { set_error("tcg_norm_to_ap: channel count must be a power of two"); return PDS_ERR_UNSUPPORTED; }
  // units of 256 items, two per CTA iteration; at most two waves of the resident CTAs
  const unsigned units = (unsigned)(Z * Y) * (unsigned)((X * (C / 8) + 255) / 256);
  unsigned gx = (units + 1) / 2;
  const unsigned cap = (unsigned)(num_sms() * (b ? 4 : 6));
  if (gx > cap) gx = cap;
  dim3 grid(gx, (unsigned)n);
  static const bool detail = getenv("PDS_B200_PROFILE_DETAIL") && atoi(getenv("PDS_B200_PROFILE_DETAIL"));
  const char* name = b ? "tcg_norm2_to_ap" : "tcg_norm_to_ap";
  if (detail) {
    static std::mutex mu;
    static std::set<std::string> names;
    std::string nm = std::string(name) + "[" + std::to_string(C) + "ch " + std::to_string(Z) + "x" + std::to_string(Y) +
                     "x" + std::to_string(X) + " ph" + std::to_string(phases) + (bcast ? " +bcast" : "") +
                     (out_f32 ? " +f32" : "") + "]";
    std::lock_guard<std::mutex> lock(mu);
    name = names.insert(nm).first->c_str();
  }
  PDS_KERNEL(name, st);
  PDS_KERNEL_WORK(0, (double)n * Z * Y * X * C * (4.0 + 2.0 * S + (b ? 4.0 : 0.0) + (out_f32 ? 4.0 : 0.0)));
  const size_t smem = (size_t)4 * C * sizeof(float);
#define PDS_TCG_NORM_CASE(FF, SS) \
  if ((fp16 != 0) == FF && S == SS) {  \
    if (b) PDS_CUDA(launch_pdl(tcg_norm_to_ap_kernel<FF, SS, true>, grid, dim3(256), smem, st, p));  \
    else PDS_CUDA(launch_pdl(tcg_norm_to_ap_kernel<FF, SS, false>, grid, dim3(256), smem, st, p));  \
  }
  PDS_TCG_NORM_CASE(true, 1) PDS_TCG_NORM_CASE(true, 2) PDS_TCG_NORM_CASE(true, 3)
  PDS_TCG_NORM_CASE(false, 1) PDS_TCG_NORM_CASE(false, 2) PDS_TCG_NORM_CASE(false, 3)
#undef PDS_TCG_NORM_CASE
  PDS_LAUNCH_CHECK("tcg_norm_to_ap_kernel");
  return PDS_OK;
}

}  // namespace pds

// ---- test hook (tests/test_gpu_tcg.py): one layer, PyTorch layouts in and out ---------------------
// x (n, Cin, Z, Y, X), w / bias in PyTorch layout, out (n, Cout, OZ, OY, OX), stats (n, Cout, 2) double
// or null; all device pointers.  Allocates its own scratch (test use only).
#include "conv_layers.cuh"
extern "C" int pds_tcg_conv_debug(int kind, int nd, int Cin, int Cout, int Z, int Y, int X, int S, int fp16,
                                  int n, const float* x, const float* w, const float* bias, float* out,
                                  double* stats, int lrelu, void* stream) {
  using namespace pds;
  cudaStream_t st = (cudaStream_t)stream;
  TcgLayer l;
  TcgShape sh;
  sh.kind = kind; sh.nd = nd; sh.Cin = Cin; sh.Cout = Cout; sh.Z = Z; sh.Y = Y; sh.X = X; sh.S = S;
  int rc = tcg_plan(sh, &l.plan);
  if (rc != PDS_OK) return rc;
  l.transposed = kind == TCG_TCONV4_S2; l.fp16 = fp16; l.wscale = fp16 ? 256.f : 1.f;
  const TcgPlan& pl = l.plan;
  const size_t vin = (size_t)Z * Y * X, vout = (size_t)pl.OZ * pl.OY * pl.OX;   // (merged: OZ = 2Z ...)
  char *blob = nullptr, *scratch = nullptr;
  const size_t b_xcl = align_up(n * vin * Cin * 4, 256), b_ap = align_up(pl.in_ap_bytes(n), 256),
               b_ycl = align_up(n * vout * Cout * 4, 256), b_sk = tcg_splitk_bytes(l, n);   // 0: layer is not split
  PDS_CUDA(cudaMalloc(&blob, tcg_layer_bytes(l)));
  if (cudaMalloc(&scratch, b_xcl + b_ap + b_ycl + b_sk) != cudaSuccess) { cudaFree(blob); set_error("out of memory"); return PDS_ERR_CUDA; }
  float* x_cl = (float*)scratch;
  uint16_t* ap = (uint16_t*)(scratch + b_xcl);
  float* y_cl = (float*)(scratch + b_xcl + b_ap);
  float* sk = b_sk ? (float*)(scratch + b_xcl + b_ap + b_ycl) : nullptr;
  size_t used = 0;
  rc = tcg_layer_init(l, blob, w, bias, st, &used);
  if (rc == PDS_OK) rc = nchw_to_nhwc(x, x_cl, n, Cin, vin, st);
  TcgNormSrc src; src.y = x_cl;
  if (rc == PDS_OK) rc = tcg_norm_to_ap(src, nullptr, nullptr, ap, n, Cin, Z, Y, X, S, fp16, pl.phase_arg(), st);
  if (rc == PDS_OK && stats) {
    cudaError_t e = cudaMemsetAsync(stats, 0, (size_t)n * Cout * 2 * sizeof(double), st);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemsetAsync");
  }
  const bool trace = getenv("PDS_B200_TCG_TRACE") != nullptr;
  if (trace) {
    cudaMalloc(&g_tcg_trace, (148 * 12 + 128) * sizeof(long long));
    cudaMemset(g_tcg_trace, 0, (148 * 12 + 128) * sizeof(long long));
    if (rc == PDS_OK) rc = tcg_conv_forward(l, n, ap, y_cl, nullptr, lrelu, st, sk, b_sk);   // warm (instruction cache, L2)
    cudaStreamSynchronize(st);
    cudaMemset(g_tcg_trace, 0, (148 * 12 + 128) * sizeof(long long));   // the wait tallies accumulate
  }
  if (rc == PDS_OK) rc = tcg_conv_forward(l, n, ap, y_cl, stats, lrelu, st, sk, b_sk);
  if (trace) {
    cudaStreamSynchronize(st);
    long long h[148 * 12 + 128];
    cudaMemcpy(h, g_tcg_trace, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[12] = {"producer start", "producer done", "weights ready", "first stage full", "MMA issue done",
                             "epilogue done", "stats flushed", "CTA end", "MMA waits: data", "MMA waits: accumulators",
                             "producer waits: stage", "MMA issue regions"};
    for (int k = 0; k < 12; ++k) {
      long long mx = 0, sum = 0; int cnt = 0;
      for (int c = 0; c < 148; ++c) if (h[c * 12 + 7]) { mx = mx > h[c * 12 + k] ? mx : h[c * 12 + k]; sum += h[c * 12 + k]; ++cnt; }
      printf("trace %-18s avg %8lld max %8lld cycles (%d CTAs)\n", names[k], cnt ? sum / cnt : 0, mx, cnt);
    }
    printf("issue regions of CTA 0 (start, length, gap to the next):");
    for (int r = 0; r < 64 && h[148 * 12 + 2 * r + 1]; ++r)
      printf(" [%lld %lld %lld]", h[148 * 12 + 2 * r], h[148 * 12 + 2 * r + 1] - h[148 * 12 + 2 * r],
             r + 1 < 64 && h[148 * 12 + 2 * r + 3] ? h[148 * 12 + 2 * r + 2] - h[148 * 12 + 2 * r + 1] : 0ll);
    printf("\n");
    cudaFree(g_tcg_trace); g_tcg_trace = nullptr;
  }
  if (trace && rc == PDS_OK) {
    // kernel duration back to back vs alternating with a small-shared-memory kernel (the
    // normalisation pass that precedes every convolution in the pipelines)
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_a = 0.f, ms_b = 0.f, ms_n = 0.f;
    cudaEventRecord(e0, st);
    for (int i = 0; i < 10; ++i) tcg_conv_forward(l, n, ap, y_cl, nullptr, lrelu, st);
    cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_a, e0, e1);
    float ms_s = 0.f;
    if (stats) {
      cudaEventRecord(e0, st);
      for (int i = 0; i < 10; ++i) tcg_conv_forward(l, n, ap, y_cl, stats, lrelu, st);
      cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_s, e0, e1);
      printf("timing: conv back-to-back WITH InstanceNorm sums %.1f us\n", ms_s * 100);
    }
    cudaEventRecord(e0, st);
    for (int i = 0; i < 10; ++i) tcg_norm_to_ap(src, nullptr, nullptr, ap, n, Cin, Z, Y, X, S, fp16, pl.phase_arg(), st);
    cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_n, e0, e1);
    cudaEventRecord(e0, st);
    for (int i = 0; i < 10; ++i) {
      tcg_norm_to_ap(src, nullptr, nullptr, ap, n, Cin, Z, Y, X, S, fp16, pl.phase_arg(), st);
      tcg_conv_forward(l, n, ap, y_cl, nullptr, lrelu, st);
    }
    cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_b, e0, e1);
    printf("timing: conv back-to-back %.1f us, norm alone %.1f us, norm+conv alternating %.1f us per pair\n",
           ms_a * 100, ms_n * 100, ms_b * 100);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  if (rc == PDS_OK) rc = nhwc_to_nchw(y_cl, out, n, Cout, vout, st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (rc == PDS_OK && e != cudaSuccess) rc = cuda_fail(e, "pds_tcg_conv_debug");
  cudaFree(blob); cudaFree(scratch);
  return rc;
}
