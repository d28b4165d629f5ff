/*
 * rmcrt_oracle.cpp -- CPU restatement of the reference RMCRT hot path (see rmcrt_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: checker for tests/ and smoke(), CPU baseline for bench.py.
 * All arithmetic is FP64 in the same operation order as the reference so that, with
 * randomSeed=false (per-ray MT19937 reseed), results reproduce the reference's
 * deterministic regression mode up to libm round-off.
 *
 * Reference paths are relative to /root/reference/src;
 * RMCRT/ = CCA/Components/Models/Radiation/RMCRT/.
 */
#include "rmcrt_oracle.h"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

//----------------------------------------------------------------------------------------
// MT19937, restating Core/Math/MersenneTwister.h:
//   initialize() 339-352, reload() 355-368, twist helpers 155-163, randInt() 217-231,
//   rand() 172-179, randDblExc() 190-197, FIXED_RANDOM_NUM 93,174-176,192-194.
//----------------------------------------------------------------------------------------
class MT {
 public:
  explicit MT(bool fixed = false) : fixed_(fixed) { seed(5489u); }
  void seed(uint32_t s) {
    st_[0] = s;
    for (int i = 1; i < N; ++i) st_[i] = 1812433253u * (st_[i - 1] ^ (st_[i - 1] >> 30)) + (uint32_t)i;
    regenerate();
  }
  uint32_t u32() {
    if (left_ == 0) regenerate();
    --left_;
    uint32_t y = st_[next_++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    return y ^ (y >> 18);
  }
  // integer in [0,n]  (MersenneTwister.h:233-250)
  uint32_t randInt(uint32_t n) {
    uint32_t used = n;
    used |= used >> 1; used |= used >> 2; used |= used >> 4; used |= used >> 8; used |= used >> 16;
    uint32_t i;
    do i = u32() & used; while (i > n);
    return i;
  }
  bool fixed() const { return fixed_; }
  // real in [0,1]
  double rand() { return fixed_ ? 0.3 : double(u32()) * (1.0 / 4294967295.0); }
  // real in (0,1)
  double randDblExc() { return fixed_ ? 0.3 : (double(u32()) + 0.5) * (1.0 / 4294967296.0); }

 private:
  enum { N = 624, M = 397 };
  static uint32_t mix(uint32_t m, uint32_t s0, uint32_t s1) {
    uint32_t y = (s0 & 0x80000000u) | (s1 & 0x7fffffffu);
    return m ^ (y >> 1) ^ ((s1 & 1u) ? 0x9908b0dfu : 0u);
  }
  void regenerate() {
    int k = 0;
    for (; k < N - M; ++k) st_[k] = mix(st_[k + M], st_[k], st_[k + 1]);
    for (; k < N - 1; ++k) st_[k] = mix(st_[k + M - N], st_[k], st_[k + 1]);
    st_[N - 1] = mix(st_[M - 1], st_[N - 1], st_[0]);
    left_ = N;
    next_ = 0;
  }
  uint32_t st_[N];
  int left_ = 0, next_ = 0;
  bool fixed_;
};

//----------------------------------------------------------------------------------------
// Array3 view: x-fastest window [lo,hi)   (Core/Grid/Variables/Array3Data.h:93,212-220)
//----------------------------------------------------------------------------------------
template <class T>
struct View {
  const T* p;
  int lo[3], n[3];
  inline const T& operator()(const int c[3]) const {
    return p[(int64_t)(c[0] - lo[0]) + (int64_t)n[0] * ((c[1] - lo[1]) + (int64_t)n[1] * (c[2] - lo[2]))];
  }
  inline const T& at(int i, int j, int k) const {
    return p[(int64_t)(i - lo[0]) + (int64_t)n[0] * ((j - lo[1]) + (int64_t)n[1] * (k - lo[2]))];
  }
};
template <class T>
View<T> makeView(const void* p, const int lo[3], const int hi[3]) {
  View<T> v;
  v.p = static_cast<const T*>(p);
  for (int d = 0; d < 3; ++d) { v.lo[d] = lo[d]; v.n[d] = hi[d] - lo[d]; }
  return v;
}

template <class T>
struct LevelT {
  View<T> abskg, sigmaT4;
  View<int32_t> cellType;
  double dx[3], anchor[3];
  int rr[3];
  int intLo[3], intHi[3];
  explicit LevelT(const oracle_level& L) {
    abskg = makeView<T>(L.abskg, L.lo, L.hi);
    sigmaT4 = makeView<T>(L.sigmaT4, L.lo, L.hi);
    cellType = makeView<int32_t>(L.cellType, L.lo, L.hi);
    for (int d = 0; d < 3; ++d) {
      dx[d] = L.dx[d]; anchor[d] = L.anchor[d]; rr[d] = L.refineRatio[d];
      intLo[d] = L.interiorLo[d]; intHi[d] = L.interiorHi[d];
    }
  }
  // Level::getCellPosition (Core/Grid/Level.cc:473-477): anchor + dcell*v + dcell*0.5
  inline void cellPosition(const int c[3], double pos[3]) const {
    for (int d = 0; d < 3; ++d) pos[d] = anchor[d] + dx[d] * c[d] + dx[d] * 0.5;
  }
};

// Level::mapCellToCoarser (Core/Grid/Level.cc:1285-1307) with level_offset=1, including its
// negative-index treatment: truncating division plus (int)fmod(idx, rr).
inline void mapToCoarser(const int idx[3], const int rr[3], int out[3]) {
  for (int d = 0; d < 3; ++d) {
    int q = idx[d] / rr[d];
    int off = 0;
    if (idx[d] < 0 && rr[d] > 1) off = (int)std::fmod((double)idx[d], (double)rr[d]);
    out[d] = q + off;
  }
}
// Level::mapCellToFiner (Level.cc:1311-1332)
inline void mapToFiner(const int idx[3], const int rr[3], int out[3]) {
  for (int d = 0; d < 3; ++d) out[d] = idx[d] * rr[d] + ((idx[d] < 0 && rr[d] > 1) ? 1 : 0);
}

// RMCRTCommon::raySignStep (RMCRTCommon.cc:482-494)
inline void signStep(const double v[3], double sign[3], int step[3]) {
  for (int d = 0; d < 3; ++d) {
    double me = std::copysign(1.0, v[d]);
    sign[d] = std::max(0.0, me);
    step[d] = int(me);
  }
}

// DDA axis choice, RMCRTCommon.cc:696-708 (same in Ray.cc:2462-2476): strict '<', ties fall to Z / Y
inline int pickAxis(const double t[3]) {
  if (t[0] < t[1]) return (t[0] < t[2]) ? 0 : 2;
  return (t[1] < t[2]) ? 1 : 2;
}

struct PathRec {
  int* buf; int stride; int cap; int n;
  inline void push(const int c[3], int L) {
    if (buf && n < cap) { int* q = buf + (int64_t)n * stride; q[0] = c[0]; q[1] = c[1]; q[2] = c[2]; if (stride > 3) q[3] = L; }
    ++n;
  }
};

const double FUZZ = 1e-12;  // RMCRT/RMCRTCommon.cc:49

//----------------------------------------------------------------------------------------
// RMCRTCommon::updateSumI<T>  (RMCRT/RMCRTCommon.cc:607-904), no RAY_SCATTER
//----------------------------------------------------------------------------------------
inline void findRayDirection(MT& mt, const oracle_params& P, const int origin[3], int iRay, double dir[3]);

// mt: the task's generator; only used by the RAY_SCATTER code (P.sigmaScat > 0), which draws from it inside the march
template <class T>
void updateSumI(const oracle_params& P, const LevelT<T>& lv, double dir[3], const double rayOrigin[3],
                const int origin[3], uint64_t& nSteps, double& sumI, PathRec* rec, MT* mt = nullptr) {
  int cur[3] = {origin[0], origin[1], origin[2]};
  int prev[3] = {cur[0], cur[1], cur[2]};
  int step[3];
  double sign[3];
  double inv[3] = {1.0 / dir[0], 1.0 / dir[1], 1.0 / dir[2]};
  signStep(dir, sign, step);  // 638 (called with ray_direction)

  double cc[3];
  lv.cellPosition(origin, cc);  // 640
  double tMax[3], tDelta[3];
  for (int d = 0; d < 3; ++d) {
    double rayDx = rayOrigin[d] - (cc[d] - 0.5 * lv.dx[d]);     // 643-646
    tMax[d] = (sign[d] * lv.dx[d] - rayDx) * inv[d];             // 649-652
    tDelta[d] = std::fabs(inv[d]) * lv.dx[d];                    // 655
  }

  bool inDomain = true;
  double tMaxPrev = 0, intensity = 1.0, fs = 1.0;
  double tau = 0, expPrev = 1.0;
  double rayLength = 0.0;            // 666
  // RAY_SCATTER (670-676): the free path to the first scattering event
  const bool scatter = P.sigmaScat > 0 && mt != nullptr;
  double rayLengthScatter = 0.0;     // 665
  double loc[3] = {rayOrigin[0], rayOrigin[1], rayOrigin[2]};   // ray_location (667)
  const double scatCoeff = std::max(P.sigmaScat, 1e-99);
  double scatLength = scatter ? -std::log(mt->randDblExc()) / scatCoeff : 0.0;

  while (intensity > P.threshold && rayLength < P.maxRayLength) {  // 680
    int ax = -9;
    while (inDomain && rayLength < P.maxRayLength) {  // 684
      prev[0] = cur[0]; prev[1] = cur[1]; prev[2] = cur[2];
      const T kPrev = lv.abskg(prev);      // 690
      const T sPrev = lv.sigmaT4(prev);    // 691
      ax = pickAxis(tMax);                 // 696-708
      cur[ax] += step[ax];                 // 712
      double ds = tMax[ax] - tMaxPrev;     // 713
      tMaxPrev = tMax[ax];
      tMax[ax] += tDelta[ax];
      if (ds > -FUZZ && ds < FUZZ) ds += FUZZ;  // 717-720
      rayLength += ds;                     // 723
      rayLengthScatter += ds;              // 724
      for (int d = 0; d < 3; ++d) loc[d] = loc[d] + (ds * dir[d]);   // 726-728
      inDomain = (lv.cellType(cur) == P.flowCell);   // 729
      tau += kPrev * ds;                   // 731
      ++nSteps;                            // 733
      if (rec) rec->push(cur, 0);
      const double e = std::exp(-tau);     // 782
      sumI += sPrev * (expPrev - e) * fs;  // 786
      expPrev = e;
      if (scatter && rayLengthScatter > scatLength && inDomain) {        // 790-836
        scatLength = -std::log(mt->randDblExc()) / scatCoeff;            // a new free path for every event
        findRayDirection(*mt, P, cur, -9, dir);                          // findRayDirection(mTwister, cur): iRay defaults to -9
        for (int d = 0; d < 3; ++d) inv[d] = 1.0 / dir[d];
        const int stepOld = step[ax];
        signStep(dir, sign, step);
        if (step[ax] * stepOld < 0) { cur[0] = prev[0]; cur[1] = prev[1]; cur[2] = prev[2]; }   // back scattering: return to prevCell
        double ccs[3];
        lv.cellPosition(cur, ccs);
        for (int d = 0; d < 3; ++d) {
          const double rayDx = loc[d] - (ccs[d] - 0.5 * lv.dx[d]);
          tMax[d] = (sign[d] * lv.dx[d] - rayDx) * inv[d];
          tDelta[d] = std::fabs(inv[d]) * lv.dx[d];
        }
        tMaxPrev = 0;
        rayLengthScatter = 0;              // allow for multiple scattering events per ray
      }
    }
    T wallEps = lv.abskg(cur);             // 846
    if (wallEps > 1.0) wallEps = 1.0;
    intensity = std::exp(-tau);            // 852
    sumI += wallEps * lv.sigmaT4(cur) * intensity;  // 854 (not multiplied by fs: SURVEY Q1)
    intensity = intensity * fs;            // 856
    if (!P.allowReflect) intensity = 0;    // 884
    if (intensity > P.threshold && P.allowReflect) {  // 898
      // RMCRTCommon::reflect (581-602)
      fs = fs * (1 - (double)lv.abskg(cur));   // reflect() takes `const double abskg` (581-588): the subtraction is in double also for T = float
      cur[0] = prev[0]; cur[1] = prev[1]; cur[2] = prev[2];
      inDomain = true;
      step[ax] *= -1;
      sign[ax] *= -1;
      dir[ax] *= -1;
    }
  }
}

inline bool containsCell(const int lo[3], const int hi[3], const int c[3], int d) {  // Ray.cc:1578-1586
  return lo[d] <= c[d] && hi[d] > c[d];
}

//----------------------------------------------------------------------------------------
// Ray::updateSumI_ML<T>  (RMCRT/Ray.cc:2349-2627), virtual ROI off
//----------------------------------------------------------------------------------------
template <class T>
void updateSumI_ML(const oracle_params& P, const std::vector<LevelT<T>>& lv, int maxLevels,
                   const double bbMin[3], const double bbMax[3],
                   const int roiLo[3], const int roiHi[3], const int* regLo, const int* regHi,
                   double dir[3], const double rayOrigin[3], const int origin[3],
                   uint64_t& nSteps, double& sumI, PathRec* rec) {
  int L = maxLevels - 1, prevLev = L;
  int cur[3] = {origin[0], origin[1], origin[2]};
  int prev[3] = {cur[0], cur[1], cur[2]};
  int step[3];
  double sign[3];
  const double inv[3] = {1.0 / dir[0], 1.0 / dir[1], 1.0 / dir[2]};  // 2395: never refreshed on reflection
  signStep(dir, sign, step);

  double ccOrigin[3];
  lv[L].cellPosition(origin, ccOrigin);  // 2413
  double tMax[3];
  for (int d = 0; d < 3; ++d) {
    double rayDx = rayOrigin[d] - (ccOrigin[d] - 0.5 * lv[L].dx[d]);  // 2417-2419
    tMax[d] = (sign[d] * lv[L].dx[d] - rayDx) * inv[d];               // 2423-2425
  }
  std::vector<double> tDelta(3 * maxLevels);
  for (int l = maxLevels - 1; l > -1; --l)
    for (int d = 0; d < 3; ++d) tDelta[3 * l + d] = std::fabs(inv[d]) * lv[l].dx[d];  // 2427-2433

  bool inDomain = true, onFine = true;
  double tMaxPrev[3] = {0, 0, 0};
  double oldLen = 0, intensity = 1.0, fs = 1.0, tau = 0, expPrev = 1.0;
  double loc[3] = {rayOrigin[0], rayOrigin[1], rayOrigin[2]};
  double cc[3] = {ccOrigin[0], ccOrigin[1], ccOrigin[2]};

  while (intensity > P.threshold) {  // 2453
    int ax = -9;
    while (inDomain) {
      prev[0] = cur[0]; prev[1] = cur[1]; prev[2] = cur[2];
      prevLev = L;
      ax = pickAxis(tMax);    // 2462-2476
      cur[ax] += step[ax];    // 2479

      lv[L].cellPosition(cur, cc);  // 2486, on the pre-jump level
      inDomain = (cc[0] >= bbMin[0] && cc[1] >= bbMin[1] && cc[2] >= bbMin[2] &&
                  cc[0] <= bbMax[0] && cc[1] <= bbMax[1] && cc[2] <= bbMax[2]);  // BBox::inside, Core/Geometry/BBox.h:96

      const bool outsideROI = !containsCell(roiLo, roiHi, cur, ax);                   // 2489
      const bool outsideRegion = !containsCell(regLo + 3 * L, regHi + 3 * L, cur, ax);  // 2490
      const bool jumpFine = (onFine && outsideROI && inDomain);                       // 2492
      const bool jumpCoarse = (!onFine && outsideRegion && (L > 0) && inDomain);      // 2493

      if (jumpFine) {  // 2505-2516
        int c2[3];
        mapToCoarser(cur, lv[L].rr, c2);
        cur[0] = c2[0]; cur[1] = c2[1]; cur[2] = c2[2];
        L = L - 1;
        onFine = false;
      } else if (jumpCoarse) {  // 2517-2527
        int c2[3];
        mapToCoarser(cur, lv[L].rr, c2);
        cur[0] = c2[0]; cur[1] = c2[1]; cur[2] = c2[2];
        L = L - 1;
      }

      const double ds = tMax[ax] - oldLen;  // 2534
      oldLen = tMax[ax];
      tMaxPrev[0] = tMax[0]; tMaxPrev[1] = tMax[1]; tMaxPrev[2] = tMax[2];
      tMax[ax] = tMax[ax] + tDelta[3 * L + ax];
      for (int d = 0; d < 3; ++d) loc[d] = loc[d] + ds * dir[d];  // 2539-2541

      if (jumpFine || jumpCoarse) {  // 2545-2550
        const double rayDxLevel = loc[ax] - (cc[ax] - 0.5 * lv[L].dx[ax]);
        const double tTmp = (sign[ax] * lv[L].dx[ax] - rayDxLevel) * inv[ax];
        tMax[0] = tMaxPrev[0]; tMax[1] = tMaxPrev[1]; tMax[2] = tMaxPrev[2];
        tMax[ax] += tTmp;
      }

      inDomain = inDomain && (lv[L].cellType(cur) == P.flowCell);  // 2554
      tau += lv[prevLev].abskg(prev) * ds;                        // 2557
      ++nSteps;
      if (rec) rec->push(cur, L);
      const double e = std::exp(-tau);                             // 2564
      sumI += lv[prevLev].sigmaT4(prev) * (expPrev - e) * fs;      // 2588
      expPrev = e;
    }
    double wallEps = lv[L].abskg(cur);  // 2594
    if (wallEps > 1.0) wallEps = 1.0;
    intensity = std::exp(-tau);
    sumI += wallEps * lv[L].sigmaT4(cur) * intensity;  // 2602
    intensity = intensity * fs;
    if (!P.allowReflect) intensity = 0;  // 2607-2609
    if (intensity > P.threshold && P.allowReflect) {  // 2621-2624; L is left unchanged
      fs = fs * (1 - (double)lv[L].abskg(cur));   // reflect(..., const double abskg, ...): double arithmetic also for T = float
      cur[0] = prev[0]; cur[1] = prev[1]; cur[2] = prev[2];
      inDomain = true;
      step[ax] *= -1;
      sign[ax] *= -1;
      dir[ax] *= -1;
    }
  }
}

// RMCRTCommon::findRayDirection (RMCRTCommon.cc:499-538)
inline void findRayDirection(MT& mt, const oracle_params& P, const int origin[3], int iRay, double dir[3]) {
  if (!P.isSeedRandom) mt.seed((uint32_t)((origin[0] + origin[1] + origin[2]) * iRay + 1));  // 504-506
  const double pm1 = 2.0 * mt.randDblExc() - 1.0 + DBL_EPSILON;
  const double r = std::sqrt(1.0 - pm1 * pm1);
  const double theta = 2.0 * M_PI * mt.randDblExc();
  dir[0] = r * std::cos(theta);
  dir[1] = r * std::sin(theta);
  dir[2] = pm1;
}

// RMCRTCommon::randVector (RMCRTCommon.cc:998-1024): modern Fisher-Yates with randInt(i); FIXED_RANDOM_NUM: 0.3*i
inline void randVector(std::vector<int>& a, MT& mt, const oracle_params& P, const int cell[3]) {
  const int max = (int)a.size();
  for (int i = 0; i < max; ++i) a[i] = i;
  if (!P.isSeedRandom) mt.seed((uint32_t)(cell[0] + cell[1] + cell[2]));   // 1011-1013
  for (int i = max - 1; i > 0; --i) {
    const int r = mt.fixed() ? (int)(0.3 * i) : (int)mt.randInt((uint32_t)i);
    std::swap(a[i], a[r]);
  }
}

// Ray::findRayDirectionHyperCube (Ray.cc:1438-1460)
inline void findRayDirectionHyperCube(MT& mt, const oracle_params& P, const int origin[3], int iRay, int bin_i, int bin_j,
                                      double dir[3]) {
  if (!P.isSeedRandom) mt.seed((uint32_t)((origin[0] + origin[1] + origin[2]) * iRay + 1));
  const double pm1 = 2.0 * (mt.randDblExc() + (double)bin_i) / P.nDivQRays - 1.0;
  const double r = std::sqrt(1.0 - pm1 * pm1);
  const double phi = 2.0 * M_PI * (mt.randDblExc() + (double)bin_j) / P.nDivQRays;
  dir[0] = r * std::cos(phi);
  dir[1] = r * std::sin(phi);
  dir[2] = pm1;
}

// RMCRTCommon::ray_Origin (RMCRTCommon.cc:544-576); draws in x,y,z order
inline void rayOriginInCell(MT& mt, const double cc[3], const double dx[3], bool ccRays, double o[3]) {
  if (!ccRays) {
    double off[3];
    off[0] = mt.rand() * dx[0];
    off[1] = mt.rand() * dx[1];
    off[2] = mt.rand() * dx[2];
    if (off[0] > dx[0] || off[1] > dx[1] || off[2] > dx[2]) { off[0] = 0.5 * dx[0]; off[1] = 0.5 * dx[1]; off[2] = 0.5 * dx[2]; }
    for (int d = 0; d < 3; ++d) o[d] = cc[d] - 0.5 * dx[d] + off[d];
  } else {
    o[0] = cc[0]; o[1] = cc[1]; o[2] = cc[2];
  }
}

// RMCRT face enum, RMCRTCommon.h:261 (its "-x +x" comment is wrong)
enum { EAST = 0, WEST = 1, NORTH = 2, SOUTH = 3, TOP = 4, BOT = 5 };
// Ray.cc:112-125
const int kDirIndexOrder[6][3] = {{2, 1, 0}, {2, 1, 0}, {0, 2, 1}, {0, 2, 1}, {0, 1, 2}, {0, 1, 2}};
const int kDirSignSwap[6][3] = {{-1, 1, 1}, {1, 1, 1}, {1, -1, 1}, {1, 1, 1}, {1, 1, -1}, {1, 1, 1}};
// Ray.cc:672
const int kUintahFace[6] = {WEST, EAST, SOUTH, NORTH, BOT, TOP};

// Ray::has_a_boundary (Ray.cc:1517-1572): neighbours tested in order W,E,S,N,B,T with `celltype+1 != 0`
inline int hasBoundary(const View<int32_t>& ct, const int c[3], int faces[6]) {
  int n = 0;
  if (ct.at(c[0] - 1, c[1], c[2]) + 1) faces[n++] = WEST;
  if (ct.at(c[0] + 1, c[1], c[2]) + 1) faces[n++] = EAST;
  if (ct.at(c[0], c[1] - 1, c[2]) + 1) faces[n++] = SOUTH;
  if (ct.at(c[0], c[1] + 1, c[2]) + 1) faces[n++] = NORTH;
  if (ct.at(c[0], c[1], c[2] - 1) + 1) faces[n++] = BOT;
  if (ct.at(c[0], c[1], c[2] + 1) + 1) faces[n++] = TOP;
  return n;
}

// Ray::rayDirection_cellFace (Ray.cc:1362-1390)
inline void rayDirectionCellFace(MT& mt, const oracle_params& P, const int origin[3], int face, int iRay,
                                 double dir[3], double& cosTheta) {
  if (!P.isSeedRandom) mt.seed((uint32_t)((origin[0] + origin[1] + origin[2]) * iRay + 1));
  const double phi = 2 * M_PI * mt.rand();
  const double theta = std::acos(mt.rand());
  cosTheta = std::cos(theta);
  double tmp[3] = {std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), cosTheta};
  for (int d = 0; d < 3; ++d) dir[d] = tmp[kDirIndexOrder[face][d]] * kDirSignSwap[face][d];
}

// Ray::rayDirectionHyperCube_cellFace (Ray.cc:1398-1432)
inline void rayDirectionHyperCubeCellFace(MT& mt, const oracle_params& P, const int origin[3], int face, int iRay,
                                          double dir[3], double& cosTheta, int bin_i, int bin_j) {
  if (!P.isSeedRandom) mt.seed((uint32_t)((origin[0] + origin[1] + origin[2]) * iRay + 1));
  cosTheta = (mt.randDblExc() + (double)bin_i) / P.nFluxRays;
  const double theta = std::acos(cosTheta);
  const double phi = 2.0 * M_PI * (mt.randDblExc() + (double)bin_j) / P.nFluxRays;
  cosTheta = std::cos(theta);
  double tmp[3] = {std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), cosTheta};
  for (int d = 0; d < 3; ++d) dir[d] = tmp[kDirIndexOrder[face][d]] * kDirSignSwap[face][d];
}

// Ray::rayLocation_cellFace (Ray.cc:1465-1513)
inline void rayLocationCellFace(MT& mt, int face, const double dx[3], const double cc[3], double o[3]) {
  const double c0[3] = {cc[0] - 0.5 * dx[0], cc[1] - 0.5 * dx[1], cc[2] - 0.5 * dx[2]};
  switch (face) {
    case WEST:  o[0] = c0[0];          o[1] = c0[1] + mt.rand() * dx[1]; o[2] = c0[2] + mt.rand() * dx[2]; break;
    case EAST:  o[0] = c0[0] + dx[0];  o[1] = c0[1] + mt.rand() * dx[1]; o[2] = c0[2] + mt.rand() * dx[2]; break;
    case SOUTH: o[0] = c0[0] + mt.rand() * dx[0]; o[1] = c0[1];          o[2] = c0[2] + mt.rand() * dx[2]; break;
    case NORTH: o[0] = c0[0] + mt.rand() * dx[0]; o[1] = c0[1] + dx[1];  o[2] = c0[2] + mt.rand() * dx[2]; break;
    case BOT:   o[0] = c0[0] + mt.rand() * dx[0]; o[1] = c0[1] + mt.rand() * dx[1]; o[2] = c0[2];          break;
    case TOP:   o[0] = c0[0] + mt.rand() * dx[0]; o[1] = c0[1] + mt.rand() * dx[1]; o[2] = c0[2] + dx[2];  break;
  }
}

struct OutWin {
  int lo[3], n[3];
  inline int64_t idx(const int c[3]) const {
    return (int64_t)(c[0] - lo[0]) + (int64_t)n[0] * ((c[1] - lo[1]) + (int64_t)n[1] * (c[2] - lo[2]));
  }
};

// Ray::rayTrace<T> patch body (Ray.cc:556-788)
template <class T>
void rayTracePatch(const oracle_params& P, const LevelT<T>& lv, const int pLo[3], const int pHi[3], uint32_t mtSeed,
                   double* divQ, double* volq, double* bflux, const OutWin& ow, uint64_t& nSteps, uint64_t& nFluxSteps) {
  MT mt(P.fixedRandom != 0);
  mt.seed(mtSeed);  // stands in for MTRand() = seed(hash(time,clock)), Ray.cc:539

  const bool lhc = P.rayDirSampleAlgo == 1;
  if (P.solveBoundaryFlux) {  // 645-730
    std::vector<int> rand_i(lhc ? P.nFluxRays : 0);     // 649
    for (int k = pLo[2]; k < pHi[2]; ++k) for (int j = pLo[1]; j < pHi[1]; ++j) for (int i = pLo[0]; i < pHi[0]; ++i) {
      const int origin[3] = {i, j, k};
      if (lv.cellType(origin) != P.flowCell) continue;  // 654-656
      int faces[6];
      const int nf = hasBoundary(lv.cellType, origin, faces);
      double* S = bflux + 7 * ow.idx(origin);
      S[6] = nf > 0 ? 1.0 : 0.0;  // boundFlux[origin].p = has_a_boundary(...), 664
      double cc[3];
      lv.cellPosition(origin, cc);
      for (int f = 0; f < nf; ++f) {
        const int rayFace = faces[f];
        double sumI = 0, sumProjI = 0, sumIprev = 0, sumCos = 0;
        if (lhc) randVector(rand_i, mt, P, origin);      // 679-681
        for (int iRay = 0; iRay < P.nFluxRays; ++iRay) {
          double dir[3], o[3], cosT;
          if (lhc) rayDirectionHyperCubeCellFace(mt, P, origin, rayFace, iRay, dir, cosT, rand_i[iRay], iRay);   // 692-695
          else rayDirectionCellFace(mt, P, origin, rayFace, iRay, dir, cosT);
          rayLocationCellFace(mt, rayFace, lv.dx, cc, o);
          updateSumI<T>(P, lv, dir, o, origin, nFluxSteps, sumI, nullptr, &mt);
          sumProjI += cosT * (sumI - sumIprev);  // 704
          sumCos += cosT;
          sumIprev = sumI;
        }
        sumProjI = sumProjI * (double)P.nFluxRays / sumCos / 2.0;             // 712
        S[kUintahFace[rayFace]] = sumProjI * 2 * M_PI / (double)P.nFluxRays;  // 716-717
      }
    }
  }

  if (P.solveDivQ) {  // 742-788
    std::vector<int> rand_i(lhc ? P.nDivQRays : 0);     // 740
    for (int k = pLo[2]; k < pHi[2]; ++k) for (int j = pLo[1]; j < pHi[1]; ++j) for (int i = pLo[0]; i < pHi[0]; ++i) {
      const int origin[3] = {i, j, k};
      if (lv.cellType(origin) != P.flowCell) continue;
      double sumI = 0, cc[3];
      lv.cellPosition(origin, cc);
      if (lhc) randVector(rand_i, mt, P, origin);        // 750-752
      for (int iRay = 0; iRay < P.nDivQRays; ++iRay) {
        double dir[3], o[3];
        if (lhc) findRayDirectionHyperCube(mt, P, origin, iRay, rand_i[iRay], iRay, dir);   // 760-761
        else findRayDirection(mt, P, origin, iRay, dir);
        rayOriginInCell(mt, cc, lv.dx, P.CCRays != 0, o);
        updateSumI<T>(P, lv, dir, o, origin, nSteps, sumI, nullptr, &mt);
      }
      const int64_t q = ow.idx(origin);
      divQ[q] = -4.0 * M_PI * lv.abskg(origin) * (lv.sigmaT4(origin) - (sumI / P.nDivQRays));  // 775
      volq[q] = 4.0 * M_PI * (sumI / P.nDivQRays);                                            // 778
    }
  }
}

// Radiometer::rayDirection_VR (Radiometer.cc:505-560)
inline void rayDirectionVR(MT& mt, const oracle_params& P, const int origin[3], int iRay, const oracle_radiometer& rad,
                           double dir[3], double& cosThetaRay) {
  if (!P.isSeedRandom) mt.seed((uint32_t)((origin[0] + origin[1] + origin[2]) * iRay + 1));   // 512-514
  const double theta_rotate = rad.theta_rotate, theta_viewAngle = rad.theta_viewAngle, xi_rotate = rad.xi_rotate,
               phi_rotate = rad.phi_rotate, range = rad.range;
  const double R1 = mt.randDblExc();
  const double R2 = mt.randDblExc();
  const double phi_ray = 2 * M_PI * R1;                                        // 526
  const double theta_ray = std::acos(std::cos(theta_viewAngle) + range * R2);  // 530
  cosThetaRay = std::cos(theta_ray);
  const double x = std::sin(theta_ray) * std::cos(phi_ray);                    // 534-536
  const double y = std::sin(theta_ray) * std::sin(phi_ray);
  const double z = cosThetaRay;
  dir[0] = x * std::cos(theta_rotate) * std::cos(xi_rotate) +                  // 539-552
           y * (-std::cos(phi_rotate) * std::sin(xi_rotate) + std::sin(phi_rotate) * std::sin(theta_rotate) * std::cos(xi_rotate)) +
           z * (std::sin(phi_rotate) * std::sin(xi_rotate) + std::cos(phi_rotate) * std::sin(theta_rotate) * std::cos(xi_rotate));
  dir[1] = x * std::cos(theta_rotate) * std::sin(xi_rotate) +
           y * (std::cos(phi_rotate) * std::cos(xi_rotate) + std::sin(phi_rotate) * std::sin(theta_rotate) * std::sin(xi_rotate)) +
           z * (-std::sin(phi_rotate) * std::cos(xi_rotate) + std::cos(phi_rotate) * std::sin(theta_rotate) * std::sin(xi_rotate));
  dir[2] = x * -std::sin(theta_rotate) + y * std::sin(phi_rotate) * std::cos(theta_rotate) + z * std::cos(phi_rotate) * std::cos(theta_rotate);
}

// Radiometer::radiometerFlux<T> (Radiometer.cc:428-500)
template <class T>
void radiometerPatch(const oracle_params& P, const LevelT<T>& lv, const oracle_radiometer* rads, int nRads, const int pLo[3],
                     const int pHi[3], uint32_t mtSeed, T* VRFlux, T* VRIntensity, const OutWin& ow, uint64_t& nSteps) {
  MT mt(P.fixedRandom != 0);
  mt.seed(mtSeed);
  for (int r = 0; r < nRads; ++r) {
    const oracle_radiometer& rad = rads[r];
    int lo[3], hi[3];
    bool hit = true;
    for (int d = 0; d < 3; ++d) {                      // doesIntersect + Max/Min (451-454)
      lo[d] = std::max(pLo[d], rad.cellLo[d]);
      hi[d] = std::min(pHi[d], rad.cellHi[d]);
      hit = hit && lo[d] < hi[d];
    }
    if (!hit) continue;
    for (int k = lo[2]; k < hi[2]; ++k) for (int j = lo[1]; j < hi[1]; ++j) for (int i = lo[0]; i < hi[0]; ++i) {
      const int c[3] = {i, j, k};
      double sumI = 0, sumProjI = 0, sumIprev = 0, cc[3];
      lv.cellPosition(c, cc);
      const int nRays = rad.nRays;
      for (int iRay = 0; iRay < nRays; ++iRay) {
        double o[3], dir[3], cosVR;
        rayOriginInCell(mt, cc, lv.dx, true, o);        // useCCRays = true (472-473)
        rayDirectionVR(mt, P, c, iRay, rad, dir, cosVR);
        updateSumI<T>(P, lv, dir, o, c, nSteps, sumI, nullptr, &mt);
        sumProjI += cosVR * (sumI - sumIprev);          // 483
        sumIprev = sumI;
      }
      const int64_t q = ow.idx(c);
      VRFlux[q] = (T)sumProjI * rad.solidAngle / nRays;   // 492: the cast binds to sumProjI
      VRIntensity[q] = (T)sumProjI / nRays;               // 493
    }
  }
}

}  // namespace

//========================================================================================
// C entry points
//========================================================================================
extern "C" {

void oracle_default_params(oracle_params* p) {
  p->nDivQRays = 10; p->nFluxRays = 1; p->threshold = 0.01; p->allowReflect = 1; p->CCRays = 0;
  p->isSeedRandom = 1; p->solveDivQ = 1; p->solveBoundaryFlux = 0; p->fixedRandom = 0; p->flowCell = -1;
  p->dataOnionFluxFlowGuard = 0;
  p->rayDirSampleAlgo = 0;
  p->maxRayLength = DBL_MAX;   // RMCRTCommon.cc:117
  p->sigmaScat = 0;
  p->roiFixed = 0;
  for (int d = 0; d < 3; ++d) { p->roiFixedLo[d] = 0; p->roiFixedHi[d] = 0; }
}

// Radiometer::problemSetup (Radiometer.cc:214-260)
void oracle_radiometer_setup(double viewAngle, const double vec[3], int isNormalVector, oracle_radiometer* r) {
  double n[3] = {vec[0], vec[1], vec[2]};
  if (isNormalVector) {                                                         // 118-120: unitNormalVec = normalVec/normalVec.length()
    const double l = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    for (int d = 0; d < 3; ++d) n[d] = n[d] / l;
  }
  for (int d = 0; d < 3; ++d) if (n[d] == 0) n[d] = 1e-16;                      // 216-220
  const double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  r->phi_rotate = 0;                                                            // 238
  r->theta_rotate = std::acos(n[2] / len);                                      // 240
  double xi = std::acos(n[0] / std::sqrt(n[0] * n[0] + n[1] * n[1]));           // 242
  if (n[0] < 0 && n[1] < 0) xi = (M_PI / 2 + xi);                               // 246-248
  if (n[0] > 0 && n[1] < 0) xi = (2 * M_PI - xi);                               // 249-251
  r->xi_rotate = xi;
  r->theta_viewAngle = viewAngle / 360 * M_PI;                                  // 258
  r->range = 1 - std::cos(r->theta_viewAngle);                                  // 259
  r->solidAngle = 2 * M_PI * r->range;                                          // 260
}

void oracle_radiometerFlux(const oracle_params* p, const oracle_level* lv, const oracle_radiometer* rads, int nRads,
                           const int patchLo[3], const int patchHi[3], uint32_t mtSeed, void* VRFlux, void* VRIntensity,
                           const int outLo[3], const int outHi[3], uint64_t* nRaySteps) {
  OutWin ow;
  for (int d = 0; d < 3; ++d) { ow.lo[d] = outLo[d]; ow.n[d] = outHi[d] - outLo[d]; }
  uint64_t n = 0;
  if (lv->isFloat) { LevelT<float> L(*lv); radiometerPatch<float>(*p, L, rads, nRads, patchLo, patchHi, mtSeed, (float*)VRFlux, (float*)VRIntensity, ow, n); }
  else             { LevelT<double> L(*lv); radiometerPatch<double>(*p, L, rads, nRads, patchLo, patchHi, mtSeed, (double*)VRFlux, (double*)VRIntensity, ow, n); }
  if (nRaySteps) *nRaySteps = n;
}

void oracle_randVector(const oracle_params* p, uint32_t mtSeed, const int cell[3], int n, int* out) {
  MT mt(p->fixedRandom != 0);
  mt.seed(mtSeed);
  std::vector<int> a(n);
  randVector(a, mt, *p, cell);
  for (int i = 0; i < n; ++i) out[i] = a[i];
}

void oracle_mt_u32(uint32_t seed, int n, uint32_t* out) { MT m; m.seed(seed); for (int i = 0; i < n; ++i) out[i] = m.u32(); }
void oracle_mt_rand(uint32_t seed, int n, double* out) { MT m; m.seed(seed); for (int i = 0; i < n; ++i) out[i] = m.rand(); }
void oracle_mt_randDblExc(uint32_t seed, int n, double* out) { MT m; m.seed(seed); for (int i = 0; i < n; ++i) out[i] = m.randDblExc(); }

void oracle_mapCellToCoarser(const int idx[3], const int rr[3], int out[3]) { mapToCoarser(idx, rr, out); }

void oracle_updateSumI(const oracle_params* p, const oracle_level* lv, const double dir[3], const double rayOrigin[3],
                       const int originCell[3], double* sumI, uint64_t* nRaySteps, int* path, int maxPath, int* nPath) {
  double d[3] = {dir[0], dir[1], dir[2]};
  PathRec rec{path, 3, maxPath, 0};
  uint64_t n = 0;
  double s = 0;
  if (lv->isFloat) { LevelT<float> L(*lv); updateSumI<float>(*p, L, d, rayOrigin, originCell, n, s, &rec); }
  else             { LevelT<double> L(*lv); updateSumI<double>(*p, L, d, rayOrigin, originCell, n, s, &rec); }
  *sumI = s; *nRaySteps = n;
  if (nPath) *nPath = rec.n;
}

void oracle_rayTrace(const oracle_params* p, const oracle_level* lv, const int patchLo[3], const int patchHi[3], uint32_t mtSeed,
                     double* divQ, double* radiationVolq, double* boundFlux, const int outLo[3], const int outHi[3],
                     uint64_t* nRaySteps, uint64_t* nFluxRaySteps) {
  OutWin ow;
  for (int d = 0; d < 3; ++d) { ow.lo[d] = outLo[d]; ow.n[d] = outHi[d] - outLo[d]; }
  uint64_t n = 0, nf = 0;
  if (lv->isFloat) { LevelT<float> L(*lv); rayTracePatch<float>(*p, L, patchLo, patchHi, mtSeed, divQ, radiationVolq, boundFlux, ow, n, nf); }
  else             { LevelT<double> L(*lv); rayTracePatch<double>(*p, L, patchLo, patchHi, mtSeed, divQ, radiationVolq, boundFlux, ow, n, nf); }
  if (nRaySteps) *nRaySteps = n;
  if (nFluxRaySteps) *nFluxRaySteps = nf;
}

// Ray::computeExtents patch_based (Ray.cc:1312-1357).  findCellIndexRange = level window incl. extra cells
// (oracle_level.lo/hi of a getLevel array); findInteriorCellIndexRange = interiorLo/Hi.
// NOTE: for the fine level pass lo/hi = the whole fine level's index range here, not an ROI window.
void oracle_computeExtents(const oracle_level* levels, int maxLevels, const int patchLo[3], const int patchHi[3],
                           const int haloCells[3], int roiLo[3], int roiHi[3], int* regionLo, int* regionHi) {
  oracle_computeExtents_roi(levels, maxLevels, patchLo, patchHi, haloCells, nullptr, nullptr, roiLo, roiHi, regionLo, regionHi);
}

void oracle_computeExtents_roi(const oracle_level* levels, int maxLevels, const int patchLo[3], const int patchHi[3],
                               const int haloCells[3], const int fixedLo[3], const int fixedHi[3], int roiLo[3], int roiHi[3],
                               int* regionLo, int* regionHi) {
  const int F = maxLevels - 1;
  for (int d = 0; d < 3; ++d) {
    const int lo = fixedLo ? fixedLo[d] : patchLo[d] - haloCells[d];   // fixed: 1303-1306; patch_based: 1316-1320
    const int hi = fixedHi ? fixedHi[d] : patchHi[d] + haloCells[d];
    roiLo[d] = std::max(lo, levels[F].lo[d]);                          // 1327-1328
    roiHi[d] = std::min(hi, levels[F].hi[d]);
    const int ec = levels[F].interiorLo[d] - levels[F].lo[d];          // fineLevel->getExtraCells()
    regionLo[3 * F + d] = roiLo[d] + ec;                               // 1333-1335
    regionHi[3 * F + d] = roiHi[d] - ec;
    regionLo[d] = levels[0].interiorLo[d];                             // 1338
    regionHi[d] = levels[0].interiorHi[d];
  }
  for (int L = maxLevels - 2; L > 0; --L) {                            // 1340-1356
    int lo[3], hi[3];
    mapToCoarser(regionLo + 3 * (L + 1), levels[L].refineRatio, lo);   // level->mapCellToCoarser uses level L's ratio
    mapToCoarser(regionHi + 3 * (L + 1), levels[L].refineRatio, hi);
    for (int d = 0; d < 3; ++d) {
      regionLo[3 * L + d] = std::max(lo[d] - haloCells[d], levels[L].interiorLo[d]);
      regionHi[3 * L + d] = std::min(hi[d] + haloCells[d], levels[L].interiorHi[d]);
    }
  }
}

}  // extern "C"

namespace {

template <class T>
void dataOnionPatch(const oracle_params& P, const oracle_level* levels, int maxLevels, const int pLo[3], const int pHi[3],
                    const int halo[3], uint32_t mtSeed, double* divQ, double* volq, double* bflux, const OutWin& ow,
                    uint64_t& nSteps, uint64_t& nFluxSteps, const int* subLo, const int* subHi) {
  // subLo/subHi (test sampling only): the ROI / region extents come from the PATCH [pLo,pHi) as in the task, but only the origin
  // cells of [subLo,subHi) inside it are traced -- a slab of a production-size patch in seconds instead of the whole patch
  std::vector<LevelT<T>> lv;
  lv.reserve(maxLevels);
  for (int l = 0; l < maxLevels; ++l) lv.emplace_back(levels[l]);
  const int F = maxLevels - 1;

  int roiLo[3], roiHi[3];
  std::vector<int> regLo(3 * maxLevels), regHi(3 * maxLevels);
  oracle_computeExtents_roi(levels, maxLevels, pLo, pHi, halo, P.roiFixed ? P.roiFixedLo : nullptr, P.roiFixed ? P.roiFixedHi : nullptr,
                            roiLo, roiHi, regLo.data(), regHi.data());

  // level_0->getInteriorSpatialRange (Ray.cc:1023-1024)
  double bbMin[3], bbMax[3];
  for (int d = 0; d < 3; ++d) {
    bbMin[d] = lv[0].anchor[d] + lv[0].dx[d] * lv[0].intLo[d];
    bbMax[d] = lv[0].anchor[d] + lv[0].dx[d] * lv[0].intHi[d];
  }

  MT mt(P.fixedRandom != 0);
  mt.seed(mtSeed);  // Ray.cc:979 / 1041-1043
  if (subLo && subHi) { pLo = subLo; pHi = subHi; }     // extents are computed: from here on [pLo,pHi) is only the cell loop

  const bool lhc = P.rayDirSampleAlgo == 1;
  if (P.solveBoundaryFlux) {  // 1091-1170, no flow-cell guard
    std::vector<int> rand_i(lhc ? P.nFluxRays : 0);     // 1096
    for (int k = pLo[2]; k < pHi[2]; ++k) for (int j = pLo[1]; j < pHi[1]; ++j) for (int i = pLo[0]; i < pHi[0]; ++i) {
      int origin[3] = {i, j, k};
      if (P.dataOnionFluxFlowGuard && lv[F].cellType(origin) != P.flowCell) continue;
      int faces[6];
      const int nf = hasBoundary(lv[F].cellType, origin, faces);
      double* S = bflux + 7 * ow.idx(origin);
      S[6] = nf > 0 ? 1.0 : 0.0;
      double cc[3];
      lv[F].cellPosition(origin, cc);
      for (int f = 0; f < nf; ++f) {
        const int rayFace = faces[f];
        double sumI = 0, sumProjI = 0, sumIprev = 0, sumCos = 0;
        if (lhc) randVector(rand_i, mt, P, origin);      // 1122-1124
        for (int iRay = 0; iRay < P.nFluxRays; ++iRay) {
          double dir[3], o[3], cosT;
          if (lhc) rayDirectionHyperCubeCellFace(mt, P, origin, rayFace, iRay, dir, cosT, rand_i[iRay], iRay);   // 1134-1136
          else rayDirectionCellFace(mt, P, origin, rayFace, iRay, dir, cosT);
          rayLocationCellFace(mt, rayFace, lv[F].dx, cc, o);
          updateSumI_ML<T>(P, lv, maxLevels, bbMin, bbMax, roiLo, roiHi, regLo.data(), regHi.data(), dir, o, origin,
                           nFluxSteps, sumI, nullptr);
          sumProjI += cosT * (sumI - sumIprev);
          sumCos += cosT;
          sumIprev = sumI;
        }
        sumProjI = sumProjI * (double)P.nFluxRays / sumCos / 2.0;
        S[kUintahFace[rayFace]] = sumProjI * 2 * M_PI / (double)P.nFluxRays;
      }
    }
  }

  if (P.solveDivQ) {  // 1177-1241
    std::vector<int> rand_i(lhc ? P.nDivQRays : 0);     // 1183
    for (int k = pLo[2]; k < pHi[2]; ++k) for (int j = pLo[1]; j < pHi[1]; ++j) for (int i = pLo[0]; i < pHi[0]; ++i) {
      int origin[3] = {i, j, k};
      if (lv[F].cellType(origin) != P.flowCell) continue;
      double cc[3], sumI = 0;
      lv[F].cellPosition(origin, cc);
      if (lhc) randVector(rand_i, mt, P, origin);        // 1196-1198
      for (int iRay = 0; iRay < P.nDivQRays; ++iRay) {
        double dir[3], o[3];
        if (lhc) findRayDirectionHyperCube(mt, P, origin, iRay, rand_i[iRay], iRay, dir);   // 1207-1208
        else findRayDirection(mt, P, origin, iRay, dir);
        rayOriginInCell(mt, cc, lv[F].dx, P.CCRays != 0, o);
        updateSumI_ML<T>(P, lv, maxLevels, bbMin, bbMax, roiLo, roiHi, regLo.data(), regHi.data(), dir, o, origin,
                         nSteps, sumI, nullptr);
      }
      const int64_t q = ow.idx(origin);
      divQ[q] = -4.0 * M_PI * lv[F].abskg(origin) * (lv[F].sigmaT4(origin) - (sumI / P.nDivQRays));  // 1226
      volq[q] = 4.0 * M_PI * (sumI / P.nDivQRays);                                                  // 1229
    }
  }
}

}  // namespace

extern "C" {

void oracle_updateSumI_ML(const oracle_params* p, const oracle_level* levels, int maxLevels, const int roiLo[3], const int roiHi[3],
                          const int* regionLo, const int* regionHi, const double dir[3], const double rayOrigin[3],
                          const int originCell[3], double* sumI, uint64_t* nRaySteps, int* path, int maxPath, int* nPath) {
  double d[3] = {dir[0], dir[1], dir[2]};
  PathRec rec{path, 4, maxPath, 0};
  uint64_t n = 0;
  double s = 0;
  double bbMin[3], bbMax[3];
  for (int k = 0; k < 3; ++k) {
    bbMin[k] = levels[0].anchor[k] + levels[0].dx[k] * levels[0].interiorLo[k];
    bbMax[k] = levels[0].anchor[k] + levels[0].dx[k] * levels[0].interiorHi[k];
  }
  if (levels[0].isFloat) {
    std::vector<LevelT<float>> lv; for (int l = 0; l < maxLevels; ++l) lv.emplace_back(levels[l]);
    updateSumI_ML<float>(*p, lv, maxLevels, bbMin, bbMax, roiLo, roiHi, regionLo, regionHi, d, rayOrigin, originCell, n, s, &rec);
  } else {
    std::vector<LevelT<double>> lv; for (int l = 0; l < maxLevels; ++l) lv.emplace_back(levels[l]);
    updateSumI_ML<double>(*p, lv, maxLevels, bbMin, bbMax, roiLo, roiHi, regionLo, regionHi, d, rayOrigin, originCell, n, s, &rec);
  }
  *sumI = s; *nRaySteps = n;
  if (nPath) *nPath = rec.n;
}

void oracle_rayTrace_dataOnion(const oracle_params* p, const oracle_level* levels, int maxLevels, const int patchLo[3],
                               const int patchHi[3], const int haloCells[3], uint32_t mtSeed, double* divQ,
                               double* radiationVolq, double* boundFlux, const int outLo[3], const int outHi[3],
                               uint64_t* nRaySteps, uint64_t* nFluxRaySteps) {
  OutWin ow;
  for (int d = 0; d < 3; ++d) { ow.lo[d] = outLo[d]; ow.n[d] = outHi[d] - outLo[d]; }
  uint64_t n = 0, nf = 0;
  if (levels[0].isFloat) dataOnionPatch<float>(*p, levels, maxLevels, patchLo, patchHi, haloCells, mtSeed, divQ, radiationVolq, boundFlux, ow, n, nf, nullptr, nullptr);
  else                   dataOnionPatch<double>(*p, levels, maxLevels, patchLo, patchHi, haloCells, mtSeed, divQ, radiationVolq, boundFlux, ow, n, nf, nullptr, nullptr);
  if (nRaySteps) *nRaySteps = n;
  if (nFluxRaySteps) *nFluxRaySteps = nf;
}

// the same task body restricted to the origin cells of [subLo,subHi) of the patch (at-size parity sampling, tests/golden/make_atsize_golden.py)
void oracle_rayTrace_dataOnion_sub(const oracle_params* p, const oracle_level* levels, int maxLevels, const int patchLo[3],
                                   const int patchHi[3], const int haloCells[3], const int subLo[3], const int subHi[3], uint32_t mtSeed,
                                   double* divQ, double* radiationVolq, double* boundFlux, const int outLo[3], const int outHi[3],
                                   uint64_t* nRaySteps, uint64_t* nFluxRaySteps) {
  OutWin ow;
  for (int d = 0; d < 3; ++d) { ow.lo[d] = outLo[d]; ow.n[d] = outHi[d] - outLo[d]; }
  uint64_t n = 0, nf = 0;
  if (levels[0].isFloat) dataOnionPatch<float>(*p, levels, maxLevels, patchLo, patchHi, haloCells, mtSeed, divQ, radiationVolq, boundFlux, ow, n, nf, subLo, subHi);
  else                   dataOnionPatch<double>(*p, levels, maxLevels, patchLo, patchHi, haloCells, mtSeed, divQ, radiationVolq, boundFlux, ow, n, nf, subLo, subHi);
  if (nRaySteps) *nRaySteps = n;
  if (nFluxRaySteps) *nFluxRaySteps = nf;
}

// RMCRTCommon::sigmaT4<T> (RMCRTCommon.cc:336-341)
void oracle_sigmaT4(const double* temp, void* out, int isFloat, int64_t n, double sigma) {
  const double sigma_over_pi = sigma / M_PI;
  for (int64_t i = 0; i < n; ++i) {
    const double T2 = temp[i] * temp[i];
    const double v = sigma_over_pi * T2 * T2;
    if (isFloat) static_cast<float*>(out)[i] = (float)v; else static_cast<double*>(out)[i] = v;
  }
}

// RMCRT_Test::initialize, benchmark 1 (Examples/RMCRT_Test.cc:646-680) on an N^3 unit cube with one layer of
// extra cells holding the Dirichlet BCs of RMCRT_bm1_1L.ups:28-93 (abskg 1, color 0, cellType 8).
void oracle_burns_christon(int N, double* abskg, double* temperature, int32_t* cellType, double T_gas, double abskg_wall,
                           double T_wall, int wallType) {
  const int M = N + 2;
  const double dx = 1.0 / N;
  for (int k = -1; k <= N; ++k) for (int j = -1; j <= N; ++j) for (int i = -1; i <= N; ++i) {
    const int64_t q = (int64_t)(i + 1) + (int64_t)M * ((j + 1) + (int64_t)M * (k + 1));
    const bool wall = (i < 0 || j < 0 || k < 0 || i >= N || j >= N || k >= N);
    if (wall) {
      abskg[q] = abskg_wall; temperature[q] = T_wall; cellType[q] = wallType;
    } else {
      abskg[q] = 0.90 * (1.0 - 2.0 * std::fabs((i - (N - 1.0) / 2.0) * dx))
                      * (1.0 - 2.0 * std::fabs((j - (N - 1.0) / 2.0) * dx))
                      * (1.0 - 2.0 * std::fabs((k - (N - 1.0) / 2.0) * dx)) + 0.1;  // 676-679
      temperature[q] = T_gas; cellType[q] = -1;
    }
  }
}

// coarsenDriver_std via fineToCoarseOperator (Core/Grid/AMR_CoarsenRefine.cc:42-86), computesAve = true,
// interior coarse cells only (Ray.cc:2141-2145); accumulation in T, x-fastest over the fine block.
void oracle_coarsen_interior(const void* fine, int isFloat, const int fLo[3], const int fHi[3], void* coarse,
                             const int cLo[3], const int cHi[3], const int cIntLo[3], const int cIntHi[3], const int rr[3]) {
  const double invRR = 1.0 / ((double)(rr[0] * rr[1] * rr[2]));
  const int fn[3] = {fHi[0] - fLo[0], fHi[1] - fLo[1], fHi[2] - fLo[2]};
  const int cn[3] = {cHi[0] - cLo[0], cHi[1] - cLo[1], cHi[2] - cLo[2]};
  (void)cn;
  for (int k = cIntLo[2]; k < cIntHi[2]; ++k) for (int j = cIntLo[1]; j < cIntHi[1]; ++j) for (int i = cIntLo[0]; i < cIntHi[0]; ++i) {
    const int c[3] = {i, j, k};
    int fs[3];
    mapToFiner(c, rr, fs);
    const int64_t cq = (int64_t)(i - cLo[0]) + (int64_t)(cHi[0] - cLo[0]) * ((j - cLo[1]) + (int64_t)(cHi[1] - cLo[1]) * (k - cLo[2]));
    if (isFloat) {
      float acc = 0.0f;
      for (int kk = 0; kk < rr[2]; ++kk) for (int jj = 0; jj < rr[1]; ++jj) for (int ii = 0; ii < rr[0]; ++ii)
        acc += static_cast<const float*>(fine)[(int64_t)(fs[0] + ii - fLo[0]) + (int64_t)fn[0] * ((fs[1] + jj - fLo[1]) + (int64_t)fn[1] * (fs[2] + kk - fLo[2]))];
      static_cast<float*>(coarse)[cq] = (float)(acc * invRR);
    } else {
      double acc = 0.0;
      for (int kk = 0; kk < rr[2]; ++kk) for (int jj = 0; jj < rr[1]; ++jj) for (int ii = 0; ii < rr[0]; ++ii)
        acc += static_cast<const double*>(fine)[(int64_t)(fs[0] + ii - fLo[0]) + (int64_t)fn[0] * ((fs[1] + jj - fLo[1]) + (int64_t)fn[1] * (fs[2] + kk - fLo[2]))];
      static_cast<double*>(coarse)[cq] = acc * invRR;
    }
  }
}

// Ray::computeCellType (Ray.cc:2272-2345): integer sum of fine cellType (computesAve=false), then *inv_RR and rounding logic
void oracle_computeCellType(const int32_t* fine, const int fLo[3], const int fHi[3], int32_t* coarse, const int cLo[3],
                            const int cHi[3], const int cIntLo[3], const int cIntHi[3], const int rr[3], int logic) {
  const double invRR = 1.0 / ((double)(rr[0] * rr[1] * rr[2]));
  const int FLOWCELL = -1, INTRUSION = 10;
  const int fn[3] = {fHi[0] - fLo[0], fHi[1] - fLo[1], fHi[2] - fLo[2]};
  for (int k = cIntLo[2]; k < cIntHi[2]; ++k) for (int j = cIntLo[1]; j < cIntHi[1]; ++j) for (int i = cIntLo[0]; i < cIntHi[0]; ++i) {
    const int c[3] = {i, j, k};
    int fs[3];
    mapToFiner(c, rr, fs);
    int acc = 0;
    for (int kk = 0; kk < rr[2]; ++kk) for (int jj = 0; jj < rr[1]; ++jj) for (int ii = 0; ii < rr[0]; ++ii)
      acc += fine[(int64_t)(fs[0] + ii - fLo[0]) + (int64_t)fn[0] * ((fs[1] + jj - fLo[1]) + (int64_t)fn[1] * (fs[2] + kk - fLo[2]))];
    const double tmp = (double)acc * invRR;
    int v = (int)tmp;
    if (tmp != FLOWCELL) {
      if (logic == 1) v = (std::fabs(tmp) / (double)INTRUSION <= (1 - invRR)) ? FLOWCELL : INTRUSION;  // ROUNDDOWN
      if (logic == 0) v = (std::fabs(tmp) / (double)INTRUSION >= invRR) ? INTRUSION : FLOWCELL;        // ROUNDUP
    }
    coarse[(int64_t)(i - cLo[0]) + (int64_t)(cHi[0] - cLo[0]) * ((j - cLo[1]) + (int64_t)(cHi[1] - cLo[1]) * (k - cLo[2]))] = v;
  }
}

//______________________________________________________________________ SURVEY 8f-4: RMCRT_coarseLevel, Ray::refine_Q
// normalizedDistance_CC (Core/Grid/AMR.cc:207-235)
static void normalizedDistance_CC(int rr, std::vector<double>& nd) {
  for (int i = 0; i < rr; i++) nd[i] = 0.0;
  if (rr > 1) {
    const int half = rr / 2;
    const double hdx = 0.5 * (1.0 / rr);
    int count = rr - 1;
    for (int i = 0; i < half; i++) { nd[i] = -count * hdx; count -= 2; }
    count = half - 1;
    for (int i = half; i < rr; i++) { nd[i] = count >= 0 ? std::fabs(nd[count]) : 0.0; count -= 1; }   // odd ratios index nd[-1] in the reference (AMR.cc:230-233); 0 here
  }
}
static inline int signOf(double d) { return d < 0 ? -1 : 1; }                 // Core/Math/MiscMath.h:63-66
static inline int roundUp(double d) {                                         // MiscMath.h:195-207 (d >= 0 here)
  return ((d - (int)d) == 0) ? (int)d : (int)(d + 1);
}

// Ray::refine_Q for one scalar field of one fine patch (Ray.cc:1916-1934): selectInterpolator (Core/Grid/AMR.h:708-745) over fine
// cells [fl,fh) = the patch's interior cells (getCoarseLevelRange, AMR.cc:132-133).
//   order 0: piecewiseConstantInterpolation (AMR.h:49-61); 1: linearInterpolation (106-167); 2: quadraticInterpolation (341-431).
// qC is x-fastest over the coarse window [cLo,cHi) (the getRegion box: coarse range of the patch +- 1 cell); gridLo/gridHi =
// coarseLevel->findCellIndexRange() incl. extra cells.  qF is x-fastest over [oLo,oHi) and only [fl,fh) is written.
void oracle_refine_Q(const double* qC, const int cLo[3], const int cHi[3], const int gridLo[3], const int gridHi_[3],
                     const int rr[3], int order, const int fl[3], const int fh[3], double* qF, const int oLo[3], const int oHi[3]) {
  const int64_t cnx = cHi[0] - cLo[0], cny = cHi[1] - cLo[1];
  const int64_t onx = oHi[0] - oLo[0], ony = oHi[1] - oLo[1];
  auto C = [&](int i, int j, int k) -> double { return qC[(i - cLo[0]) + cnx * ((j - cLo[1]) + cny * (int64_t)(k - cLo[2]))]; };
  std::vector<double> ndx(rr[0]), ndy(rr[1]), ndz(rr[2]);
  normalizedDistance_CC(rr[0], ndx); normalizedDistance_CC(rr[1], ndy); normalizedDistance_CC(rr[2], ndz);
  const int gridHi[3] = {gridHi_[0] - 1, gridHi_[1] - 1, gridHi_[2] - 1};      // AMR.h:353
  for (int kf = fl[2]; kf < fh[2]; ++kf) for (int jf = fl[1]; jf < fh[1]; ++jf) for (int i_f = fl[0]; i_f < fh[0]; ++i_f) {
    const int f[3] = {i_f, jf, kf};
    int c[3];
    mapToCoarser(f, rr, c);
    double out;
    if (order == 0) {
      out = C(c[0], c[1], c[2]);
    } else if (order == 1) {
      const int rel[3] = {f[0] - c[0] * rr[0], f[1] - c[1] * rr[1], f[2] - c[2] * rr[2]};
      const double dist[3] = {ndx[rel[0]], ndy[rel[1]], ndz[rel[2]]};
      int i = signOf(dist[0]), j = signOf(dist[1]), k = signOf(dist[2]);
      i *= roundUp(std::fabs(dist[0])); j *= roundUp(std::fabs(dist[1])); k *= roundUp(std::fabs(dist[2]));
      const double x = std::fabs(dist[0]), y = std::fabs(dist[1]), z = std::fabs(dist[2]);
      const double w0 = (1.0 - x) * (1.0 - y), w1 = x * (1.0 - y), w2 = y * (1.0 - x), w3 = x * y;
      const double p1 = w0 * C(c[0], c[1], c[2]) + w1 * C(c[0] + i, c[1], c[2]) + w2 * C(c[0], c[1] + j, c[2]) + w3 * C(c[0] + i, c[1] + j, c[2]);
      const double p2 = w0 * C(c[0], c[1], c[2] + k) + w1 * C(c[0] + i, c[1], c[2] + k) + w2 * C(c[0], c[1] + j, c[2] + k) + w3 * C(c[0] + i, c[1] + j, c[2] + k);
      out = (1.0 - z) * p1 + z * p2;
    } else {
      int b[3], rel[3];
      for (int d = 0; d < 3; ++d) {
        int shift = 0;
        if ((c[d] - gridLo[d]) == 0) shift = 1;
        if ((gridHi[d] - c[d]) == 0) shift = -1;
        b[d] = c[d] + shift;
        // AMR.h:386 takes f - baseCell*rr; with a shifted base cell that leaves [0,rr) and the reference reads its distance table out
        // of range.  Unreachable from refine_Q (interior fine cells never map to a coarse extra cell); keep the in-range index.
        rel[d] = shift ? f[d] - c[d] * rr[d] : f[d] - b[d] * rr[d];
      }
      const double x = ndx[rel[0]], y = ndy[rel[1]], z = ndz[rel[2]];
      const double wx[3] = {0.5 * x * (x - 1.0), -(x + 1.0) * (x - 1.0), 0.5 * x * (x + 1.0)};
      const double wy[3] = {0.5 * y * (y - 1.0), -(y + 1.0) * (y - 1.0), 0.5 * y * (y + 1.0)};
      const double wz[3] = {0.5 * z * (z - 1.0), -(z + 1.0) * (z - 1.0), 0.5 * z * (z + 1.0)};
      double plane[3];
      for (int pz = 0; pz < 3; ++pz) {
        const int k = pz - 1;
        double acc = (wx[0] * wy[0]) * C(b[0] - 1, b[1] - 1, b[2] + k);
        acc += (wx[1] * wy[0]) * C(b[0], b[1] - 1, b[2] + k);
        acc += (wx[2] * wy[0]) * C(b[0] + 1, b[1] - 1, b[2] + k);
        acc += (wx[0] * wy[1]) * C(b[0] - 1, b[1], b[2] + k);
        acc += (wx[1] * wy[1]) * C(b[0], b[1], b[2] + k);
        acc += (wx[2] * wy[1]) * C(b[0] + 1, b[1], b[2] + k);
        acc += (wx[0] * wy[2]) * C(b[0] - 1, b[1] + 1, b[2] + k);
        acc += (wx[1] * wy[2]) * C(b[0], b[1] + 1, b[2] + k);
        acc += (wx[2] * wy[2]) * C(b[0] + 1, b[1] + 1, b[2] + k);
        plane[pz] = acc;
      }
      out = wz[0] * plane[0] + wz[1] * plane[1] + wz[2] * plane[2];
    }
    qF[(f[0] - oLo[0]) + onx * ((f[1] - oLo[1]) + ony * (int64_t)(f[2] - oLo[2]))] = out;
  }
}

//______________________________________________________________________ SURVEY 8f-2: the Arches caller's property tasks
// RMCRT_Radiation::sigmaT4<T> (Arches/SourceTerms/RMCRT.cc:939-1016) over n cells (the patch incl. extra cells).
// partGas_absk / partGas_temp: nPartGas arrays; [0] is the gas phase (radiation_temperature).
void oracle_arches_sigmaT4(const double* gasVolFrac, const double* const* partGas_absk, const double* const* partGas_temp,
                           int nPartGas, int doPartRadiation, void* out, int isFloat, int64_t n, double sigma) {
  const double sigma_over_pi = sigma / M_PI;
  for (int64_t c = 0; c < n; ++c) {
    double v;
    if (!doPartRadiation || !(gasVolFrac[c] > 1e-16)) {
      const double T_sqrd = partGas_temp[0][c] * partGas_temp[0][c];
      v = sigma_over_pi * T_sqrd * T_sqrd;
    } else {
      double sumT = 0.0, sumAbsk = 0.0;
      for (int i = 0; i < nPartGas; i++) {
        const double T_sqrd = partGas_temp[i][c] * partGas_temp[i][c];
        sumT += T_sqrd * T_sqrd * partGas_absk[i][c];
        sumAbsk += partGas_absk[i][c];
      }
      v = 0.0;
      if (sumAbsk > 1e-16) v = sigma_over_pi * sumT / sumAbsk;
    }
    if (isFloat) static_cast<float*>(out)[c] = (float)v; else static_cast<double*>(out)[c] = v;
  }
}

// RMCRT_Radiation::sumAbsk<T> (RMCRT.cc:1062-1130): arrays x-fastest over the patch window [lo,hi) incl. extra cells; interior
// cells [intLo,intHi); every cell outside the interior is a boundary-face cell (ExtraPlusEdgeCells) and gets the hard-wired 1.
void oracle_arches_sumAbsk(const double* gasVolFrac, const double* const* partGas_absk, int nPartGas, const int lo[3],
                           const int hi[3], const int intLo[3], const int intHi[3], void* out, int isFloat) {
  const int64_t nx = hi[0] - lo[0], ny = hi[1] - lo[1];
  for (int k = lo[2]; k < hi[2]; ++k) for (int j = lo[1]; j < hi[1]; ++j) for (int i = lo[0]; i < hi[0]; ++i) {
    const int64_t c = (i - lo[0]) + nx * ((j - lo[1]) + ny * (int64_t)(k - lo[2]));
    const bool interior = i >= intLo[0] && i < intHi[0] && j >= intLo[1] && j < intHi[1] && k >= intLo[2] && k < intHi[2];
    double v = 0.0;
    if (interior) {
      for (int p = 0; p < nPartGas; p++) {
        if (gasVolFrac[c] > 1e-16) v += partGas_absk[p][c]; else v = 1.0;
      }
    } else {
      v = 1.0;
    }
    if (isFloat) static_cast<float*>(out)[c] = (float)v; else static_cast<double*>(out)[c] = v;
  }
}

// RMCRT_Radiation::stencilToDBLs (RMCRT.cc:1389-1423): Stencil7 {w,e,s,n,b,t,p} -> radiationFlux{E,W,N,S,T,B}
void oracle_arches_stencilToDBLs(const double* boundFlux7, int64_t n, double* E, double* W, double* N, double* S, double* T, double* B) {
  for (int64_t c = 0; c < n; ++c) {
    const double* me = boundFlux7 + 7 * c;
    E[c] = me[1]; W[c] = me[0]; N[c] = me[3]; S[c] = me[2]; T[c] = me[5]; B[c] = me[4];
  }
}

// the ray generation of the divQ loop (Ray.cc:752-760), exposed so it can be compared with the reference's own findRayDirection / ray_Origin
// (oracle/_ref/libref_rmcrt.so): rays iRay = 0..nRays-1 of one cell on one MT seeded with mtSeed; dirs / origins: [nRays][3]
void oracle_generate_rays(const oracle_params* p, const oracle_level* lv, const int cell[3], int nRays, uint32_t mtSeed, double* dirs, double* origins) {
  MT mt(p->fixedRandom != 0);
  mt.seed(mtSeed);
  double cc[3];
  for (int d = 0; d < 3; ++d) cc[d] = lv->anchor[d] + lv->dx[d] * cell[d] + lv->dx[d] * 0.5;     // Level.cc:473-477
  for (int i = 0; i < nRays; ++i) {
    findRayDirection(mt, *p, cell, i, dirs + 3 * i);
    rayOriginInCell(mt, cc, lv->dx, p->CCRays != 0, origins + 3 * i);
  }
}

// one ray with the task generator in a known state (RAY_SCATTER draws from it inside the march): MT seeded with mtSeed
void oracle_updateSumI_seeded(const oracle_params* p, const oracle_level* lv, const double dir[3], const double rayOrigin[3],
                              const int originCell[3], uint32_t mtSeed, double* sumI, uint64_t* nRaySteps) {
  double d[3] = {dir[0], dir[1], dir[2]};
  uint64_t n = 0;
  double s = 0;
  MT mt(p->fixedRandom != 0);
  mt.seed(mtSeed);
  if (lv->isFloat) { LevelT<float> L(*lv); updateSumI<float>(*p, L, d, rayOrigin, originCell, n, s, nullptr, &mt); }
  else             { LevelT<double> L(*lv); updateSumI<double>(*p, L, d, rayOrigin, originCell, n, s, nullptr, &mt); }
  *sumI = s; *nRaySteps = n;
}

}  // extern "C"
