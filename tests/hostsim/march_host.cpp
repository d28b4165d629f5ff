// Host build of the production march code (uintah_b200/csrc/march.cuh, rng.cuh) for CPU-side logic tests.
#include "cuda_shim.h"
#include <vector>
#include "../../uintah_b200/csrc/march.cuh"
#include "../../uintah_b200/csrc/rng.cuh"
#include "../../uintah_b200/csrc/indep.cuh"
using namespace rmcrt;

extern "C" {

// pack like pack_kernel: sign bit <=> cellType != flowCell
void hs_pack(const double* abskg, const double* sig, const int* ct, long long n, int flowCell, int kind, void* out) {
  for (long long i = 0; i < n; ++i) {
    double k = std::fabs(abskg[i]);
    if (ct[i] != flowCell) k = -k;
    if (kind == 0) static_cast<Rec8*>(out)[i] = Rec8{(float)k, (float)sig[i]};
    else static_cast<Rec16*>(out)[i] = Rec16{k, sig[i]};
  }
}

struct HsLevel { const void* rec; int lo[3], hi[3], intLo[3], intHi[3], rr[3]; double dx[3], anchor[3]; };

static void rayDxOf(const HsLevel& g, const int c[3], const double* o, double r[3]) {
  for (int d = 0; d < 3; ++d) { const double cc = g.anchor[d] + g.dx[d] * c[d] + g.dx[d] * 0.5; r[d] = o[d] - (cc - 0.5 * g.dx[d]); }
}

namespace {
struct Do2Setup {
  OnionGeom<4> G;
  RoiBox roi;
  std::vector<Rec8> buf;
};
// the shell-wrapped ROI copy of roi_copy_kernel + the two-level geometry
void do2_setup(const HsLevel* lv, const int* roiLo, const int* roiHi, const int* regLo, const int* regHi, Do2Setup& S) {
  OnionGeom<4>& G = S.G;
  RoiBox& roi = S.roi;
  std::memset(&G, 0, sizeof G);
  G.maxLevels = 2;
  for (int L = 0; L < 2; ++L) {
    LevelGeom& g = G.lev[L];
    g.rec = lv[L].rec;
    for (int d = 0; d < 3; ++d) {
      g.lo[d] = lv[L].lo[d]; g.n[d] = lv[L].hi[d] - lv[L].lo[d]; g.intLo[d] = lv[L].intLo[d]; g.intHi[d] = lv[L].intHi[d];
      g.rr[d] = lv[L].rr[d]; g.regLo[d] = regLo[3 * L + d]; g.regHi[d] = regHi[3 * L + d]; g.dx[d] = lv[L].dx[d];
      g.anchor[d] = lv[L].anchor[d]; g.off[d] = lv[L].anchor[d] - lv[1].anchor[d];
    }
  }
  long long cellsB = 1;
  for (int d = 0; d < 3; ++d) { roi.lo[d] = roiLo[d] - 1; roi.n[d] = roiHi[d] - roiLo[d] + 2; cellsB *= roi.n[d]; }
  const long long guard = (long long)roi.n[0] * roi.n[1] + roi.n[0] + 1;
  S.buf.assign((size_t)(cellsB + 2 * guard), Rec8{-INFINITY, 0.0f});
  Rec8* base = S.buf.data() + guard;
  const LevelGeom& gf = G.lev[1];
  const Rec8* fine = static_cast<const Rec8*>(gf.rec);
  for (int k = roiLo[2]; k < roiHi[2]; ++k)
    for (int j = roiLo[1]; j < roiHi[1]; ++j)
      for (int i = roiLo[0]; i < roiHi[0]; ++i) {
        const long long src = (long long)(i - gf.lo[0]) + (long long)gf.n[0] * ((j - gf.lo[1]) + (long long)gf.n[1] * (k - gf.lo[2]));
        const long long dst = (long long)(i - roi.lo[0]) + (long long)roi.n[0] * ((j - roi.lo[1]) + (long long)roi.n[1] * (k - roi.lo[2]));
        base[dst] = fine[src];
      }
  roi.rec = base;
}
}  // namespace

// single level: kind 0 = fp32/Rec8, 1 = fp64/Rec16
void hs_trace_1L(const HsLevel* g, int kind, int n, const int* cells, const double* dirs, const double* orgs, double threshold,
                 int allowReflect, int skipZero, double* sumI, unsigned* steps, int* path, int maxPath, double maxRayLength) {
  const double maxLen = (maxRayLength > 0 && maxRayLength < 1e300) ? maxRayLength : INFINITY;
  const int nx = g->hi[0] - g->lo[0], ny = g->hi[1] - g->lo[1];
  for (int t = 0; t < n; ++t) {
    const int* c = cells + 3 * t;
    double r[3];
    rayDxOf(*g, c, orgs + 3 * t, r);
    const long long idx = (long long)(c[0] - g->lo[0]) + (long long)nx * ((c[1] - g->lo[1]) + (long long)ny * (c[2] - g->lo[2]));
    PathSink sink{path ? path + (long long)t * maxPath * 4 : nullptr, maxPath, 0};
    unsigned s = 0;
    const double* d = dirs + 3 * t;
    if (kind == 0)
      sumI[t] = march_single_level<float, Rec8, true>(static_cast<const Rec8*>(g->rec), idx, nx, nx * ny, (float)d[0], (float)d[1], (float)d[2],
                  (float)r[0], (float)r[1], (float)r[2], (float)g->dx[0], (float)g->dx[1], (float)g->dx[2], (float)threshold,
                  allowReflect != 0, skipZero != 0, s, path ? &sink : nullptr, c[0], c[1], c[2],
                  (float)(maxLen * (double)Num<float>::LEN_SCALE));
    else
      sumI[t] = march_single_level<double, Rec16, true>(static_cast<const Rec16*>(g->rec), idx, nx, nx * ny, d[0], d[1], d[2], r[0], r[1], r[2],
                  g->dx[0], g->dx[1], g->dx[2], threshold, allowReflect != 0, skipZero != 0, s, path ? &sink : nullptr, c[0], c[1], c[2],
                  maxLen);
    steps[t] = s;
  }
}

// production FP32 fast path (march_single_level_fast), same inputs as hs_trace_1L kind 0
void hs_trace_1L_fast(const HsLevel* g, int n, const int* cells, const double* dirs, const double* orgs, double threshold,
                      int allowReflect, int skipZero, double* sumI, unsigned* steps) {
  const int nx = g->hi[0] - g->lo[0], ny = g->hi[1] - g->lo[1];
  for (int t = 0; t < n; ++t) {
    const int* c = cells + 3 * t;
    double r[3];
    rayDxOf(*g, c, orgs + 3 * t, r);
    const int idx = (c[0] - g->lo[0]) + nx * ((c[1] - g->lo[1]) + ny * (c[2] - g->lo[2]));
    unsigned s = 0;
    const double* d = dirs + 3 * t;
    sumI[t] = march_single_level_fast(static_cast<const Rec8*>(g->rec), idx, nx, nx * ny, (float)d[0], (float)d[1], (float)d[2],
                                      (float)r[0], (float)r[1], (float)r[2], (float)g->dx[0], (float)g->dx[1], (float)g->dx[2],
                                      (float)threshold, allowReflect != 0, skipZero != 0, s, 8);
    steps[t] = s;
  }
}

void hs_trace_ML(const HsLevel* lv, int maxLevels, const int* roiLo, const int* roiHi, const int* regLo, const int* regHi, int kind, int n,
                 const int* cells, const double* dirs, const double* orgs, double threshold, int allowReflect, int skipZero,
                 double* sumI, unsigned* steps, int* path, int maxPath) {
  OnionGeom<4> G;
  std::memset(&G, 0, sizeof G);
  G.maxLevels = maxLevels;
  const int F = maxLevels - 1;
  for (int L = 0; L < maxLevels; ++L) {
    LevelGeom& g = G.lev[L];
    g.rec = lv[L].rec;
    for (int d = 0; d < 3; ++d) {
      g.lo[d] = lv[L].lo[d]; g.n[d] = lv[L].hi[d] - lv[L].lo[d]; g.intLo[d] = lv[L].intLo[d]; g.intHi[d] = lv[L].intHi[d];
      g.rr[d] = lv[L].rr[d]; g.regLo[d] = regLo[3 * L + d]; g.regHi[d] = regHi[3 * L + d]; g.dx[d] = lv[L].dx[d];
      g.anchor[d] = lv[L].anchor[d]; g.off[d] = lv[L].anchor[d] - lv[F].anchor[d];
    }
  }
  for (int d = 0; d < 3; ++d) { G.roiLo[d] = roiLo[d]; G.roiHi[d] = roiHi[d]; }
  for (int t = 0; t < n; ++t) {
    const int* c = cells + 3 * t;
    double r[3];
    rayDxOf(lv[F], c, orgs + 3 * t, r);
    PathSink sink{path ? path + (long long)t * maxPath * 4 : nullptr, maxPath, 0};
    unsigned s = 0;
    const double* d = dirs + 3 * t;
    if (kind == 0)
      sumI[t] = march_multi_level<float, Rec8, 4, true>(G, c[0], c[1], c[2], (float)d[0], (float)d[1], (float)d[2], (float)r[0], (float)r[1],
                  (float)r[2], (float)threshold, allowReflect != 0, skipZero != 0, s, path ? &sink : nullptr);
    else
      sumI[t] = march_multi_level<double, Rec16, 4, true>(G, c[0], c[1], c[2], d[0], d[1], d[2], r[0], r[1], r[2], threshold,
                  allowReflect != 0, skipZero != 0, s, path ? &sink : nullptr);
    steps[t] = s;
  }
}

// production 2-level Data-Onion fast path (march_onion2_fast): builds the shell-wrapped ROI copy the way roi_copy_kernel
// does and marches the same rays as hs_trace_ML kind 0
void hs_trace_DO2_fast(const HsLevel* lv, const int* roiLo, const int* roiHi, const int* regLo, const int* regHi, int n,
                       const int* cells, const double* dirs, const double* orgs, double threshold, int allowReflect, int skipZero,
                       double* sumI, unsigned* steps) {
  Do2Setup S;
  do2_setup(lv, roiLo, roiHi, regLo, regHi, S);
  const OnionGeom<4>& G = S.G;
  const RoiBox& roi = S.roi;
  for (int t = 0; t < n; ++t) {
    const int* c = cells + 3 * t;
    double r[3];
    rayDxOf(lv[1], c, orgs + 3 * t, r);
    unsigned s = 0;
    const double* d = dirs + 3 * t;
    sumI[t] = march_onion2_fast<Rec8, 4>(G, roi, c[0], c[1], c[2], (float)d[0], (float)d[1], (float)d[2], (float)r[0], (float)r[1],
                                         (float)r[2], (float)threshold, allowReflect != 0, skipZero != 0, s, 8);
    steps[t] = s;
  }
}

void hs_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  const u32x4 r = philox4x32_10(u32x4{ctr[0], ctr[1], ctr[2], ctr[3]}, key[0], key[1]);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
void hs_keyed_permutation(uint32_t n, const uint32_t key[4], uint32_t* out) {
  const u32x4 k{key[0], key[1], key[2], key[3]};
  for (uint32_t x = 0; x < n; ++x) out[x] = keyed_permutation(x, n, k);
}
void hs_mt_first5(uint32_t seed, uint32_t out[5]) { mt19937_first<5>(seed, out); }
void hs_mt_first4(uint32_t seed, uint32_t out[4]) { mt19937_first<4>(seed, out); }

// isotropic scattering march (march_single_level_scatter) with the regression-mode generator: MT19937 seeded with `seeds[t]`, no output
// consumed yet -- the state oracle_updateSumI_seeded starts from.  kind 0 = fp32/Rec8, 1 = fp64/Rec16
void hs_trace_1L_scatter(const HsLevel* g, int kind, int n, const int* cells, const double* dirs, const double* orgs, const unsigned* seeds,
                         double threshold, int allowReflect, double sigmaScat, double maxRayLength, int floatProps, double* sumI, unsigned* steps) {
  const double maxLen = (maxRayLength > 0 && maxRayLength < 1e300) ? maxRayLength : INFINITY;
  const int nx = g->hi[0] - g->lo[0], ny = g->hi[1] - g->lo[1];
  for (int t = 0; t < n; ++t) {
    const int* c = cells + 3 * t;
    const long long idx = (long long)(c[0] - g->lo[0]) + (long long)nx * ((c[1] - g->lo[1]) + (long long)ny * (c[2] - g->lo[2]));
    ScatRng rng;
    rng.mode = 1; rng.key0 = rng.key1 = rng.c0 = rng.c2 = rng.c3 = rng.ray = rng.draw = 0;
    rng.mtSeed = seeds[t]; rng.mtPos = 0;
    unsigned s = 0;
    if (kind == 0)
      sumI[t] = march_single_level_scatter<float, Rec8>(static_cast<const Rec8*>(g->rec), idx, nx, nx * ny, c[0], c[1], c[2], dirs + 3 * t, orgs + 3 * t,
                  g->anchor, g->dx, (float)threshold, allowReflect != 0, s, (float)(maxLen * (double)Num<float>::LEN_SCALE), false, sigmaScat, rng);
    else
      sumI[t] = march_single_level_scatter<double, Rec16>(static_cast<const Rec16*>(g->rec), idx, nx, nx * ny, c[0], c[1], c[2], dirs + 3 * t, orgs + 3 * t,
                  g->anchor, g->dx, threshold, allowReflect != 0, s, maxLen, floatProps != 0, sigmaScat, rng);
    steps[t] = s;
  }
}

// two cells per lane (march_single_level_dual): the rays of cell c and of cell c + (0,1,0) with the SAME direction and in-cell origin
void hs_trace_1L_dual(const HsLevel* g, int n, const int* cells, const double* dirs, const double* orgs, double threshold, int allowReflect,
                      int skipZero, int rejoin, int pairK, double* sumA, double* sumB, unsigned* stepsA, unsigned* stepsB) {
  const int nx = g->hi[0] - g->lo[0], ny = g->hi[1] - g->lo[1];
  for (int t = 0; t < n; ++t) {
    const int* c = cells + 3 * t;
    double r[3];
    rayDxOf(*g, c, orgs + 3 * t, r);
    const int idx = (c[0] - g->lo[0]) + nx * ((c[1] - g->lo[1]) + ny * (c[2] - g->lo[2]));
    unsigned sa = 0, sb = 0;
    float a, b;
    const double* d = dirs + 3 * t;
    const int offB = pairK ? nx * ny : nx;
    if (rejoin)
      march_single_level_dual<true>(static_cast<const Rec8*>(g->rec), idx, offB, nx, nx * ny, (float)d[0], (float)d[1], (float)d[2], (float)r[0],
                                    (float)r[1], (float)r[2], (float)g->dx[0], (float)g->dx[1], (float)g->dx[2], (float)threshold,
                                    allowReflect != 0, skipZero != 0, sa, sb, 8, a, b);
    else
      march_single_level_dual<false>(static_cast<const Rec8*>(g->rec), idx, offB, nx, nx * ny, (float)d[0], (float)d[1], (float)d[2], (float)r[0],
                                     (float)r[1], (float)r[2], (float)g->dx[0], (float)g->dx[1], (float)g->dx[2], (float)threshold,
                                     allowReflect != 0, skipZero != 0, sa, sb, 8, a, b);
    sumA[t] = a; sumB[t] = b; stepsA[t] = sa; stepsB[t] = sb;
  }
}
// four cells per lane (march_single_level_quad): the rays of the 2 x 2 cells c, c + (0,1,0), c + (0,0,1), c + (0,1,1) with the SAME direction
// and in-cell origin; steps = the sum over the four rays (what the kernel counts)
void hs_trace_1L_quad(const HsLevel* g, int n, const int* cells, const double* dirs, const double* orgs, double threshold, int allowReflect,
                      int skipZero, double* sums /* [n][4] */, unsigned* steps /* [n] */) {
  const int nx = g->hi[0] - g->lo[0], ny = g->hi[1] - g->lo[1];
  for (int t = 0; t < n; ++t) {
    const int* c = cells + 3 * t;
    double r[3];
    rayDxOf(*g, c, orgs + 3 * t, r);
    const int idx = (c[0] - g->lo[0]) + nx * ((c[1] - g->lo[1]) + ny * (c[2] - g->lo[2]));
    unsigned st = 0;
    float sm[4];
    const double* d = dirs + 3 * t;
    march_single_level_quad(static_cast<const Rec8*>(g->rec), idx, nx, nx * ny, (float)d[0], (float)d[1], (float)d[2], (float)r[0], (float)r[1],
                            (float)r[2], (float)g->dx[0], (float)g->dx[1], (float)g->dx[2], (float)threshold, allowReflect != 0, skipZero != 0,
                            st, 8, sm);
    for (int q = 0; q < 4; ++q) sums[4 * t + q] = sm[q];
    steps[t] = st;
  }
}
// 2-level Data Onion, four cells per lane (march_onion2_quad): the rays of the 2 x 2 cells c, c+(0,1,0), c+(0,0,1), c+(0,1,1) with the SAME
// direction and in-cell origin; sums [n][4], steps = total of the four rays
void hs_trace_DO2_quad(const HsLevel* lv, const int* roiLo, const int* roiHi, const int* regLo, const int* regHi, int n,
                       const int* cells, const double* dirs, const double* orgs, double threshold, int allowReflect, int skipZero,
                       double* sums, unsigned* steps) {
  Do2Setup S;
  do2_setup(lv, roiLo, roiHi, regLo, regHi, S);
  for (int t = 0; t < n; ++t) {
    const int* c = cells + 3 * t;
    double r[3];
    rayDxOf(lv[1], c, orgs + 3 * t, r);
    unsigned st = 0;
    float sm[4];
    const double* d = dirs + 3 * t;
    march_onion2_quad<Rec8, 4>(S.G, S.roi, c[0], c[1], c[2], (float)d[0], (float)d[1], (float)d[2], (float)r[0], (float)r[1], (float)r[2],
                               (float)threshold, allowReflect != 0, skipZero != 0, st, 8, sm);
    for (int q = 0; q < 4; ++q) sums[4 * t + q] = sm[q];
    steps[t] = st;
  }
}

// ---- the default solve's march: blocked half-packed copy (blk_build_kernel) + march_leg_blk / BlkRay (indep.cuh)
static void blk_build_host(const HsLevel& g, std::vector<uint32_t>& blk, BlkGeom& bg) {
  const Rec8* recs = static_cast<const Rec8*>(g.rec);
  int n[3], nb[3];
  for (int d = 0; d < 3; ++d) { n[d] = g.hi[d] - g.lo[d]; nb[d] = (n[d] + 4 + 1) / 2; bg.lo[d] = g.lo[d] - 2; }
  blk.assign((size_t)nb[0] * nb[1] * nb[2] * 8, 0x80000000u);      // guard: a wall with abskg 0, sigmaT4 0
  for (int b2 = 0; b2 < nb[2]; ++b2) for (int b1 = 0; b1 < nb[1]; ++b1) for (int b0 = 0; b0 < nb[0]; ++b0)
    for (int c = 0; c < 8; ++c) {
      const int x = 2 * b0 + (c & 1) - 2, y = 2 * b1 + ((c >> 1) & 1) - 2, z = 2 * b2 + (c >> 2) - 2;
      if (x < 0 || x >= n[0] || y < 0 || y >= n[1] || z < 0 || z >= n[2]) continue;
      const Rec8 r = recs[(long long)x + (long long)n[0] * (y + (long long)n[1] * z)];
      blk[((size_t)b0 + (size_t)nb[0] * (b1 + (size_t)nb[1] * b2)) * 8 + c] =
          ((uint32_t)float_to_half_bits(r.k) << 16) | float_to_half_bits(r.s * HALF_S_SCALE);
    }
  bg.blk = blk.data(); bg.bx = nb[0]; bg.bxy = nb[0] * nb[1];
}

// legsPerCall 0: march_single_level_blk (the whole ray); k > 0: blk_ray_legs(k) again and again, the ray's state going through the 12
// words divq_sorted_kernel<RESORT> parks between two calls
void hs_trace_1L_blk(const HsLevel* g, int n, const int* cells, const double* dirs, const double* orgs, double threshold, int allowReflect,
                     int skipZero, int legsPerCall, double* sumI, unsigned* steps) {
  std::vector<uint32_t> store;
  BlkGeom bg;
  blk_build_host(*g, store, bg);
  const int sy8 = 8 * bg.bx, sz8 = 8 * bg.bxy;
  const float cx = (float)g->dx[0], cy = (float)g->dx[1], cz = (float)g->dx[2];
  for (int t = 0; t < n; ++t) {
    const int* c = cells + 3 * t;
    double r[3];
    rayDxOf(*g, c, orgs + 3 * t, r);
    const int ux = c[0] - bg.lo[0], uy = c[1] - bg.lo[1], uz = c[2] - bg.lo[2];
    const unsigned par = (unsigned)((ux & 1) | ((uy & 1) << 1) | ((uz & 1) << 2));
    const int w0 = 8 * ((ux >> 1) + bg.bx * (uy >> 1) + bg.bxy * (uz >> 1)) + (int)par;
    const double* d = dirs + 3 * t;
    unsigned s = 0;
    if (legsPerCall <= 0) {
      sumI[t] = (1.0f / HALF_S_SCALE) * march_single_level_blk(bg.blk, w0, par, sy8, sz8, (float)d[0], (float)d[1], (float)d[2], (float)r[0],
                                                               (float)r[1], (float)r[2], cx, cy, cz, (float)threshold, allowReflect != 0,
                                                               skipZero != 0, s);
    } else if (!(1.0f > (float)threshold)) {
      sumI[t] = 0.0;
    } else {
      BlkRay ry;
      float ddx, ddy, ddz;
      blk_ray_start(ry, bg.blk, w0, par, sy8, sz8, (float)d[0], (float)d[1], (float)d[2], (float)r[0], (float)r[1], (float)r[2], cx, cy, cz,
                    ddx, ddy, ddz);
      while (!blk_ray_legs(ry, bg.blk, ddx, ddy, ddz, sy8, sz8, (float)threshold, allowReflect != 0, skipZero != 0, legsPerCall)) {
        // park and resume: tMax, word index | deltas, tau offset | exp(-tau), steps, sum, weight
        const unsigned w[12] = {__float_as_uint(ry.d.tx), __float_as_uint(ry.d.ty), __float_as_uint(ry.d.tz), (unsigned)ry.d.idx,
                                (unsigned)ry.b.dlx, (unsigned)ry.b.dly, (unsigned)ry.b.dlz, __float_as_uint(ry.q.A),
                                __float_as_uint(ry.q.ePrev), __float_as_uint(ry.q.nsteps), __float_as_uint(ry.sum), __float_as_uint(ry.fs)};
        BlkRay z;
        z.d.tx = __uint_as_float(w[0]); z.d.ty = __uint_as_float(w[1]); z.d.tz = __uint_as_float(w[2]); z.d.idx = (int)w[3];
        z.b.dlx = (int)w[4]; z.b.dly = (int)w[5]; z.b.dlz = (int)w[6];
        blk_decode(z.q, bg.blk[z.d.idx]);
        z.q.A = __uint_as_float(w[7]); z.q.ePrev = __uint_as_float(w[8]); z.q.nsteps = __uint_as_float(w[9]);
        z.q.tau = 0.0f; z.q.leg = 0.0f;
        z.sum = __uint_as_float(w[10]); z.fs = __uint_as_float(w[11]);
        ry = z;
      }
      s = (unsigned)ry.q.nsteps;
      sumI[t] = (1.0f / HALF_S_SCALE) * ry.sum;
    }
    steps[t] = s;
  }
}

// ---- Data Onion on blocked boxes (march_onion_blk): level 0 whole, every level above it the box of onion_copy_box (rmcrt_b200.cu) --
// the ROI / region plus the level's extra cells where it ends at the domain boundary -- wrapped in markers (roi_blk_build_kernel)
void hs_trace_DO_blk(const HsLevel* lv, int maxLevels, const int* roiLo, const int* roiHi, const int* regLo, const int* regHi, int n,
                     const int* cells, const double* dirs, const double* orgs, double threshold, int allowReflect, int skipZero,
                     double* sumI, unsigned* steps) {
  OnionGeom<4> G;
  std::memset(&G, 0, sizeof G);
  G.maxLevels = maxLevels;
  const int F = maxLevels - 1;
  for (int L = 0; L < maxLevels; ++L) {
    LevelGeom& g = G.lev[L];
    g.rec = lv[L].rec;
    for (int d = 0; d < 3; ++d) {
      g.lo[d] = lv[L].lo[d]; g.n[d] = lv[L].hi[d] - lv[L].lo[d]; g.intLo[d] = lv[L].intLo[d]; g.intHi[d] = lv[L].intHi[d];
      g.rr[d] = lv[L].rr[d]; g.regLo[d] = regLo[3 * L + d]; g.regHi[d] = regHi[3 * L + d]; g.dx[d] = lv[L].dx[d];
      g.anchor[d] = lv[L].anchor[d]; g.off[d] = lv[L].anchor[d] - lv[F].anchor[d];
    }
  }
  for (int d = 0; d < 3; ++d) { G.roiLo[d] = roiLo[d]; G.roiHi[d] = roiHi[d]; }
  OnionBlk ob;
  std::memset(&ob, 0, sizeof ob);
  std::vector<uint32_t> store[4];
  {
    BlkGeom bg0;
    blk_build_host(lv[0], store[0], bg0);
    ob.blk[0] = bg0.blk; ob.bx[0] = bg0.bx; ob.bxy[0] = bg0.bxy;
    for (int d = 0; d < 3; ++d) ob.lo[0][d] = bg0.lo[d];
  }
  for (int L = 1; L <= F; ++L) {
    const LevelGeom& g = G.lev[L];
    const Rec8* recs = static_cast<const Rec8*>(g.rec);
    int lo[3], hi[3], nb[3];
    for (int d = 0; d < 3; ++d) {
      const int rlo = L == F ? roiLo[d] : g.regLo[d], rhi = L == F ? roiHi[d] : g.regHi[d];
      lo[d] = rlo <= g.intLo[d] ? g.lo[d] : rlo;
      hi[d] = rhi >= g.intHi[d] ? g.lo[d] + g.n[d] : rhi;
      ob.lo[L][d] = lo[d] - 1 - 2;
      nb[d] = (hi[d] - lo[d] + 2 + 4 + 1) / 2;
    }
    store[L].assign((size_t)nb[0] * nb[1] * nb[2] * 8, 0xFC000000u);      // marker: wall flag, abskg = inf
    for (int z = lo[2]; z < hi[2]; ++z) for (int y = lo[1]; y < hi[1]; ++y) for (int x = lo[0]; x < hi[0]; ++x) {
      const Rec8 r = recs[(long long)(x - g.lo[0]) + (long long)g.n[0] * ((y - g.lo[1]) + (long long)g.n[1] * (z - g.lo[2]))];
      const int ux = x - ob.lo[L][0], uy = y - ob.lo[L][1], uz = z - ob.lo[L][2];
      store[L][((size_t)(ux >> 1) + (size_t)nb[0] * ((uy >> 1) + (size_t)nb[1] * (uz >> 1))) * 8 + ((ux & 1) | ((uy & 1) << 1) | ((uz & 1) << 2))] =
          ((uint32_t)float_to_half_bits(r.k) << 16) | float_to_half_bits(r.s * HALF_S_SCALE);
    }
    ob.blk[L] = store[L].data(); ob.bx[L] = nb[0]; ob.bxy[L] = nb[0] * nb[1];
  }
  for (int t = 0; t < n; ++t) {
    const int* c = cells + 3 * t;
    double r[3];
    rayDxOf(lv[F], c, orgs + 3 * t, r);
    const int ux = c[0] - ob.lo[F][0], uy = c[1] - ob.lo[F][1], uz = c[2] - ob.lo[F][2];
    const int w0 = 8 * ((ux >> 1) + ob.bx[F] * (uy >> 1) + ob.bxy[F] * (uz >> 1)) + ((ux & 1) | ((uy & 1) << 1) | ((uz & 1) << 2));
    const double* d = dirs + 3 * t;
    unsigned s = 0;
    sumI[t] = (1.0f / HALF_S_SCALE) * march_onion_blk<4>(G, ob, c[0], c[1], c[2], w0, (float)d[0], (float)d[1], (float)d[2], (float)r[0],
                                                         (float)r[1], (float)r[2], (float)threshold, allowReflect != 0, skipZero != 0, s);
    steps[t] = s;
  }
}
}
