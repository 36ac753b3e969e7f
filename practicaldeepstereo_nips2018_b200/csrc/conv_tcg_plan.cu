// Host-side planner of the generic tcgen05 convolution engine (conv_tcg.cuh).
// Pure host code: no CUDA call, so tests can run it -- and emulate its programs
// against the oracle -- on a machine without a GPU (pds_tcg_plan_describe).
#include <algorithm>

#include "conv_tcg.cuh"

namespace pds {
namespace {

struct DimTap { int k, phase, off; };

// Taps of ONE spatial dimension for output parity class `cls`:
//   CONV3_S1  in = o + k - 1
//   CONV3_S2  in = 2o + k - 1 -> phase (k-1)&1, index o + floor((k-1)/2)
//   CONV5_S2  in = 2o + k - 2 -> phase k&1,     index o + floor((k-2)/2)
//   TCONV4_S2 out = 2z + cls; in = z + cls - t, kernel index 1 - cls + 2t  (t = 0, 1)
//   CONV3_S1X4 (x only): group X covers outputs 4X .. 4X+3; extended tap k' = 0..5 reads input
//              4X + k' - 1 -> phase (k'-1) mod 4, group offset floor((k'-1)/4); output j of the group
//              sees it through kernel index k' - j (zero weight outside 0..2)
std::vector<DimTap> dim_taps(int kind, int cls, int dim) {
  if (kind == TCG_CONV3_S1X4)
    return dim == 2 ? std::vector<DimTap>{{0, 3, -1}, {1, 0, 0}, {2, 1, 0}, {3, 2, 0}, {4, 3, 0}, {5, 0, 1}}
                    : std::vector<DimTap>{{0, 0, -1}, {1, 0, 0}, {2, 0, 1}};
  switch (kind) {
    case TCG_CONV3_S1: case TCG_TCONV4_S2M: return {{0, 0, -1}, {1, 0, 0}, {2, 0, 1}};   // merged: k = offset + 1
    case TCG_CONV3_S2: return {{0, 1, -1}, {1, 0, 0}, {2, 1, 0}};
    case TCG_CONV5_S2: return {{0, 0, -1}, {1, 1, -1}, {2, 0, 0}, {3, 1, 0}, {4, 0, 1}};
    default: return {{1 - cls, 0, cls}, {3 - cls, 0, cls - 1}};
  }
}

struct Tap { int k[3], phase[3], off[3]; };   // z, y, x

int pad_n(int cout) { return cout <= 16 ? 16 : cout <= 32 ? 32 : cout <= 64 ? 64 : 128; }

constexpr size_t kSmemBudget = 227 * 1024 - 4096;   // barriers, bias, per-CTA InstanceNorm sums, alignment slack

}  // namespace

int tcg_plan(const TcgShape& sh, TcgPlan* out) {
  TcgPlan p;
  p.shape = sh;
  const bool strided = sh.kind == TCG_CONV3_S2 || sh.kind == TCG_CONV5_S2;
  const bool merged = sh.kind == TCG_TCONV4_S2M;
  const bool transposed = sh.kind == TCG_TCONV4_S2;
  const bool x4 = sh.kind == TCG_CONV3_S1X4;
  if (sh.nd != 2 && sh.nd != 3) { set_error("tcg_plan: nd must be 2 or 3"); return PDS_ERR_UNSUPPORTED; }
  if (sh.nd == 2 && sh.Z != 1) { set_error("tcg_plan: 2-D layers need Z == 1"); return PDS_ERR_UNSUPPORTED; }
  if (sh.Cin < 8 || sh.Cin % 8 || (sh.Cin > 8 && sh.Cin % 16) || sh.Cout < 1 || sh.Cout > 128 ||
      sh.S < 1 || sh.S > 3 || sh.X < 1 || sh.Y < 1 || sh.Z < 1) {
    set_error("tcg_plan: unsupported channels / terms (Cin=%d Cout=%d S=%d)", sh.Cin, sh.Cout, sh.S);
    return PDS_ERR_UNSUPPORTED;
  }
  if (strided && ((sh.X & 1) || (sh.Y & 1) || (sh.nd == 3 && (sh.Z & 1)))) {
    set_error("tcg_plan: stride-2 layers need even extents");
    return PDS_ERR_UNSUPPORTED;
  }
  if (merged && (sh.nd != 3 || 8 * sh.Cout > 128)) {
    set_error("tcg_plan: merged transposed layers are 3-D with Cout <= 16");
    return PDS_ERR_UNSUPPORTED;
  }
  if (x4 && (sh.Cin != 8 || 4 * sh.Cout > 128 || (sh.X & 3))) {
    set_error("tcg_plan: the four-voxels-per-row layer needs Cin == 8, Cout <= 32 and X %% 4 == 0");
    return PDS_ERR_UNSUPPORTED;
  }
  if (transposed && sh.nd != 3) { set_error("tcg_plan: transposed layers are 3-D only"); return PDS_ERR_UNSUPPORTED; }
  const int nd = sh.nd;
  const int div = strided ? 2 : 1;
  p.xg = x4 ? 4 : 1;
  p.IZ = nd == 3 ? sh.Z / div : 1; p.IY = sh.Y / div; p.IX = sh.X / (x4 ? 4 : div);
  p.GZ = p.IZ; p.GY = p.IY; p.GX = p.IX;
  const int mul = (transposed || merged) ? 2 : 1;
  p.OZ = p.GZ * (nd == 3 ? mul : 1); p.OY = p.GY * mul; p.OX = p.GX * mul * p.xg;
  p.ncls = transposed ? 8 : 1;
  p.nph = strided ? (nd == 3 ? 8 : 4) : (x4 ? 4 : 1);
  p.P = sh.Cin / 8;
  p.PB = sh.Cin >= 16 ? 2 : 1;
  p.N = pad_n(merged ? 8 * sh.Cout : p.xg * sh.Cout);
  p.merged = merged ? 1 : 0;
  const int S = sh.S;
  const int nchunks = sh.Cin >= 16 ? sh.Cin / 16 : 1;
  const unsigned ent16 = 2u * S * p.N;          // 16-byte units of one entry's B operand

  // per-dimension offset range over all classes
  int min_off[3] = {0, 0, 0}, max_off[3] = {0, 0, 0};
  for (int d = 3 - nd; d < 3; ++d)
    for (int c = 0; c < (transposed ? 2 : 1); ++c)
      for (const DimTap& t : dim_taps(sh.kind, c, d)) {
        min_off[d] = std::min(min_off[d], t.off);
        max_off[d] = std::max(max_off[d], t.off);
      }

  int nacc = std::max(1, 128 / p.N);
  while (nacc > 1 && nacc * S * p.N > 512) nacc /= 2;
  int ntx, ntz;
  auto arrange = [&](int n) {
    if (nd == 2) { ntx = n; ntz = 1; return; }
    ntx = x4 ? 1 : (n >= 4 ? 2 : 1);      // voxel groups: a row of the grid is short, stack the tiles in z (2 x 2 measured the same)
    ntz = n / ntx;
    while (ntz > 1 && ntz / 2 >= p.GZ) { ntz /= 2; }   // no point in stacking beyond the grid
  };

  for (;; nacc /= 2) {
    arrange(nacc);
    const int T[3] = {ntz, 16, 8 * ntx};
    // split candidates: (split_z, split_y); the x-phase layer tries one unit per PHASE first (one box
    // with the full z / y halo per stage: splitting its taps by kz would fetch every plane three times)
    const int cand[4][2] = {{0, 0}, {0, 0}, {nd == 3 ? 1 : 0, nd == 3 ? 0 : 1}, {1, 1}};
    for (int ci = x4 ? 0 : 1; ci < (nd == 3 ? 4 : 3); ++ci) {
      const bool split[3] = {cand[ci][0] != 0, cand[ci][1] != 0, false};
      const bool split_ph = x4 && ci == 0;
      int B[3];
      for (int d = 0; d < 3; ++d) B[d] = d < 3 - nd ? 1 : T[d] + (split[d] ? 0 : max_off[d] - min_off[d]);
      p.BZ = B[0]; p.BY = B[1]; p.BX = B[2];
      const unsigned plane16 = (unsigned)(p.BZ * p.BY * p.BX);
      p.box_bytes = (unsigned)align_up((size_t)p.PB * plane16 * 16, 128);
      const unsigned box16 = p.box_bytes / 16;
      p.units.clear(); p.boxes.clear(); p.entries.clear(); p.wsrc.clear();
      unsigned w16 = 0, max_unit_w = 0;
      int max_boxes = 0;
      bool ok = true;
      for (int cls = 0; cls < p.ncls; ++cls) {
        const int cc[3] = {(cls >> 2) & 1, (cls >> 1) & 1, cls & 1};
        std::vector<DimTap> dt[3];
        for (int d = 0; d < 3; ++d)
          dt[d] = d < 3 - nd ? std::vector<DimTap>{{0, 0, 0}} : dim_taps(sh.kind, cc[d], d);
        // tap groups
        const int gz = split[0] ? (int)dt[0].size() : 1, gy = split[1] ? (int)dt[1].size() : 1;
        const int gp = split_ph ? p.nph : 1;
        int units_this_class = 0;
        for (int c = 0; c < nchunks; ++c)
          for (int iz = 0; iz < gz; ++iz)
            for (int iy = 0; iy < gy; ++iy)
            for (int ip = 0; ip < gp; ++ip) {
              TcgUnit u;
              u.ent_beg = (int)p.entries.size(); u.box_beg = (int)p.boxes.size();
              struct Ref { unsigned a; Tap t; };
              std::vector<Ref> refs;
              std::vector<int> phases;   // phase id of each box of this unit
              for (size_t tz = 0; tz < dt[0].size(); ++tz) {
                if (split[0] && (int)tz != iz) continue;
                for (size_t ty = 0; ty < dt[1].size(); ++ty) {
                  if (split[1] && (int)ty != iy) continue;
                  for (size_t tx = 0; tx < dt[2].size(); ++tx) {
                    Tap t;
                    const DimTap* s[3] = {&dt[0][tz], &dt[1][ty], &dt[2][tx]};
                    for (int d = 0; d < 3; ++d) { t.k[d] = s[d]->k; t.phase[d] = s[d]->phase; t.off[d] = s[d]->off; }
                    const int ph = x4 ? t.phase[2]
                                      : nd == 3 ? (t.phase[0] * 2 + t.phase[1]) * 2 + t.phase[2]
                                                : t.phase[1] * 2 + t.phase[2];
                    if (split_ph && ph != ip) continue;
                    int bi = -1;
                    for (size_t j = 0; j < phases.size(); ++j) if (phases[j] == ph) bi = (int)j;
                    if (bi < 0) {
                      bi = (int)phases.size();
                      phases.push_back(ph);
                      TcgBox b;
                      b.dz = split[0] ? t.off[0] : min_off[0];
                      b.dy = split[1] ? t.off[1] : min_off[1];
                      b.dx = min_off[2];
                      b.plane = ph * p.P + (sh.Cin >= 16 ? 2 * c : 0);
                      p.boxes.push_back(b);
                    }
                    const int l[3] = {split[0] ? 0 : t.off[0] - min_off[0],
                                      split[1] ? 0 : t.off[1] - min_off[1], t.off[2] - min_off[2]};
                    Ref r;
                    r.a = (unsigned)bi * box16 + (unsigned)((l[0] * p.BY + l[1]) * p.BX + l[2]);
                    r.t = t;
                    refs.push_back(r);
                  }
                }
              }
              if (sh.Cin >= 16) {
                for (const Ref& r : refs) {
                  TcgEntry e;
                  e.a = r.a | (plane16 << 16);
                  e.b = 0;
                  TcgWeightSrc w;
                  for (int h = 0; h < 2; ++h) { w.kz[h] = r.t.k[0]; w.ky[h] = r.t.k[1]; w.kx[h] = r.t.k[2]; w.group[h] = 2 * c + h; }
                  p.entries.push_back(e); p.wsrc.push_back(w);
                }
                if (plane16 >= (1u << 14)) ok = false;
              } else {
                std::sort(refs.begin(), refs.end(), [](const Ref& x, const Ref& y) { return x.a < y.a; });
                for (size_t i = 0; i < refs.size(); i += 2) {
                  TcgEntry e;
                  TcgWeightSrc w;
                  w.kz[0] = refs[i].t.k[0]; w.ky[0] = refs[i].t.k[1]; w.kx[0] = refs[i].t.k[2]; w.group[0] = 0;
                  unsigned lbo = 0;
                  if (i + 1 < refs.size()) {
                    lbo = refs[i + 1].a - refs[i].a;
                    w.kz[1] = refs[i + 1].t.k[0]; w.ky[1] = refs[i + 1].t.k[1]; w.kx[1] = refs[i + 1].t.k[2]; w.group[1] = 0;
                  } else {
                    w.kz[1] = w.ky[1] = w.kx[1] = 0; w.group[1] = -1;   // zero weights; reads the same tap again
                  }
                  if (lbo >= (1u << 14)) ok = false;
                  e.a = refs[i].a | (lbo << 16);
                  e.b = 0;
                  p.entries.push_back(e); p.wsrc.push_back(w);
                }
              }
              u.ent_end = (int)p.entries.size(); u.box_end = (int)p.boxes.size();
              u.w_off16 = w16;
              u.w_bytes = (unsigned)(u.ent_end - u.ent_beg) * ent16 * 16;
              w16 += (unsigned)(u.ent_end - u.ent_beg) * ent16;
              max_unit_w = std::max(max_unit_w, u.w_bytes);
              max_boxes = std::max(max_boxes, u.box_end - u.box_beg);
              p.units.push_back(u);
              ++units_this_class;
            }
        p.units_per_item = units_this_class;
      }
      if (!ok) continue;
      p.w_total_bytes = w16 * 16;
      p.nacc = ntx * ntz; p.ntx = ntx; p.ntz = ntz; p.max_boxes = max_boxes;
      p.tile_off16.clear();
      for (int iz = 0; iz < ntz; ++iz)
        for (int ix = 0; ix < ntx; ++ix) p.tile_off16.push_back(iz * p.BY * p.BX + ix * 8);
      const size_t tables = align_up(p.units.size() * sizeof(TcgUnit) + p.boxes.size() * sizeof(TcgBox) +
                                     p.entries.size() * sizeof(TcgEntry) + p.tile_off16.size() * 4, 128) + 256;
      const size_t a_bytes = (size_t)max_boxes * S * p.box_bytes;
      for (int resident = 1; resident >= 0; --resident) {
        const size_t fixed = tables + (resident ? align_up(p.w_total_bytes, 128) : 0);
        const size_t stage = align_up(a_bytes + (resident ? 0 : max_unit_w), 128);
        if (fixed + 2 * stage > kSmemBudget) continue;
        int stages = (int)((kSmemBudget - fixed) / stage);
        // resident weights are worth it only if they leave room for 3 stages (or all units of an item)
        if (resident && stages < 3 && stages < p.units_per_item) continue;
        p.resident = resident;
        p.stages = std::min(stages, 6);
        p.stage_bytes = (unsigned)stage;
        p.wres_bytes = resident ? (unsigned)align_up(p.w_total_bytes, 128) : 0;
        // entry B offsets: relative to the layer (resident) or to the unit's slab (streamed)
        for (const TcgUnit& u : p.units)
          for (int e = u.ent_beg; e < u.ent_end; ++e)
            p.entries[e].b = (resident ? u.w_off16 : (unsigned)(a_bytes / 16)) + (unsigned)(e - u.ent_beg) * ent16;
        *out = p;
        return PDS_OK;
      }
    }
    if (nacc == 1) break;
  }
  set_error("tcg_plan: no tiling of this layer fits shared memory (Cin=%d Cout=%d kind=%d)", sh.Cin,
            sh.Cout, sh.kind);
  return PDS_ERR_UNSUPPORTED;
}

}  // namespace pds

// ---- test hook (not part of include/pds_b200.h: host logic only, used by tests/test_tcg_plan.py) ----
// Serialises the plan of a layer into `buf` (int32 words); returns the number of words needed.
// Layout: header[40] | units[nu][6] | boxes[nb][4] | entries[ne][2] | wsrc[ne][8] | tile_off[nacc]
extern "C" int pds_tcg_plan_describe(int kind, int nd, int Cin, int Cout, int Z, int Y, int X, int S,
                                     int* buf, int buf_words) {
  using namespace pds;
  TcgShape sh;
  sh.kind = kind; sh.nd = nd; sh.Cin = Cin; sh.Cout = Cout; sh.Z = Z; sh.Y = Y; sh.X = X; sh.S = S;
  TcgPlan p;
  const int rc = tcg_plan(sh, &p);
  if (rc != PDS_OK) return -rc;
  const int nu = (int)p.units.size(), nb = (int)p.boxes.size(), ne = (int)p.entries.size();
  const int need = 40 + nu * 6 + nb * 4 + ne * 2 + ne * 8 + p.nacc;
  if (!buf || buf_words < need) return need;
  const int hdr[40] = {p.N, p.nacc, p.ntx, p.ntz, p.ncls, p.nph, p.P, p.GZ, p.GY, p.GX, p.OZ, p.OY, p.OX,
                       p.IZ, p.IY, p.IX, p.BX, p.BY, p.BZ, p.PB, p.units_per_item, p.resident, p.stages,
                       (int)p.box_bytes, (int)p.stage_bytes, (int)p.wres_bytes, (int)p.w_total_bytes,
                       nu, nb, ne, p.max_boxes, p.merged, p.xg};
  int* w = buf;
  for (int i = 0; i < 40; ++i) *w++ = hdr[i];
  for (const TcgUnit& u : p.units) {
    *w++ = u.ent_beg; *w++ = u.ent_end; *w++ = u.box_beg; *w++ = u.box_end; *w++ = (int)u.w_off16; *w++ = (int)u.w_bytes;
  }
  for (const TcgBox& b : p.boxes) { *w++ = b.dx; *w++ = b.dy; *w++ = b.dz; *w++ = b.plane; }
  for (const TcgEntry& e : p.entries) { *w++ = (int)e.a; *w++ = (int)e.b; }
  for (const TcgWeightSrc& s : p.wsrc) {
    *w++ = s.kz[0]; *w++ = s.ky[0]; *w++ = s.kx[0]; *w++ = s.group[0];
    *w++ = s.kz[1]; *w++ = s.ky[1]; *w++ = s.kx[1]; *w++ = s.group[1];
  }
  for (int t : p.tile_off16) *w++ = t;
  return need;
}
