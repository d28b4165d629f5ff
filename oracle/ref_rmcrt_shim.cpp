// ref_rmcrt_shim.cpp -- TEST INFRASTRUCTURE: compiles the reference's OWN per-ray routines
//   RMCRTCommon::raySignStep / findRayDirection / ray_Origin / reflect / updateSumI<T>   (RMCRT/RMCRTCommon.cc:482-904)
//   Ray::containsCell / Ray::updateSumI_ML<T>                                            (RMCRT/Ray.cc:1578-1586, 2349-2627)
// cut out of the reference tree at build time by oracle/extract_ref.py into oracle/_ref/ref_rmcrt_extracted.inc, together with the
// reference's own Core/Math/MersenneTwister.h (-I/root/reference/src), into oracle/_ref/libref_rmcrt.so (git-ignored, travels to the
// GPU box) and, with -DRAY_SCATTER (the reference's --enable-ray-scatter configuration), into oracle/_ref/libref_rmcrt_scatter.so.  tests/test_oracle_ref.py compares the oracle's restatement with these bit for bit.
//
// What is NOT the reference's here: the value types below.  The reference's functions use a handful of members of IntVector, Vector,
// Point, BBox, Level, constCCVariable and InternalError; the real classes pull in the whole Core library (and MPI through the grid),
// so this file supplies look-alikes with exactly those members.  Their arithmetic is component-wise double arithmetic
// (Core/Geometry/Vector.h: operator/ and operator* of two Vectors are component-wise, Abs() is fabs per component), cell position is
// Level.cc:473-477 (anchor + dcell*v + dcell*0.5, in that order), mapCellToCoarser is Level.cc:1285-1307, BBox::inside is BBox.h:96 -- the same restatements the oracle makes, so a
// bit-level agreement pins every line of the marching routines themselves.
#include <Core/Math/MersenneTwister.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace Uintah {

class IntVector {
 public:
  IntVector() : v_{0, 0, 0} {}
  IntVector(int x, int y, int z) : v_{x, y, z} {}
  int x() const { return v_[0]; }
  int y() const { return v_[1]; }
  int z() const { return v_[2]; }
  int& operator[](int i) { return v_[i]; }
  int operator[](int i) const { return v_[i]; }
 private:
  int v_[3];
};
inline std::ostream& operator<<(std::ostream& o, const IntVector& v) { return o << "[int " << v.x() << ", " << v.y() << ", " << v.z() << "]"; }
inline IntVector operator+(const IntVector& a, const IntVector& b) { return IntVector(a.x() + b.x(), a.y() + b.y(), a.z() + b.z()); }
inline IntVector operator-(const IntVector& a, const IntVector& b) { return IntVector(a.x() - b.x(), a.y() - b.y(), a.z() - b.z()); }
inline IntVector Max(const IntVector& a, const IntVector& b) { return IntVector(std::max(a.x(), b.x()), std::max(a.y(), b.y()), std::max(a.z(), b.z())); }
inline IntVector Min(const IntVector& a, const IntVector& b) { return IntVector(std::min(a.x(), b.x()), std::min(a.y(), b.y()), std::min(a.z(), b.z())); }

class Vector {
 public:
  Vector() : v_{0, 0, 0} {}
  explicit Vector(double a) : v_{a, a, a} {}
  Vector(double x, double y, double z) : v_{x, y, z} {}
  double x() const { return v_[0]; }
  double y() const { return v_[1]; }
  double z() const { return v_[2]; }
  void x(double a) { v_[0] = a; }
  void y(double a) { v_[1] = a; }
  void z(double a) { v_[2] = a; }
  double& operator[](int i) { return v_[i]; }
  double operator[](int i) const { return v_[i]; }
  Vector operator/(const Vector& o) const { return Vector(v_[0] / o.v_[0], v_[1] / o.v_[1], v_[2] / o.v_[2]); }   // Vector.h: component-wise
  Vector operator*(const Vector& o) const { return Vector(v_[0] * o.v_[0], v_[1] * o.v_[1], v_[2] * o.v_[2]); }
 private:
  double v_[3];
};
inline Vector Abs(const Vector& v) { return Vector(std::fabs(v.x()), std::fabs(v.y()), std::fabs(v.z())); }
inline std::ostream& operator<<(std::ostream& o, const Vector& v) { return o << "[" << v.x() << ", " << v.y() << ", " << v.z() << "]"; }

class Point {
 public:
  Point() : v_{0, 0, 0} {}
  Point(double x, double y, double z) : v_{x, y, z} {}
  double x() const { return v_[0]; }
  double y() const { return v_[1]; }
  double z() const { return v_[2]; }
  double operator()(int i) const { return v_[i]; }
 private:
  double v_[3];
};

class BBox {      // Core/Geometry/BBox.h:96
 public:
  BBox(const Point& lo, const Point& hi) : cmin(lo), cmax(hi) {}
  bool inside(const Point& p) const {
    return (p.x() >= cmin.x() && p.y() >= cmin.y() && p.z() >= cmin.z() && p.x() <= cmax.x() && p.y() <= cmax.y() && p.z() <= cmax.z());
  }
 private:
  Point cmin, cmax;
};

class Level;
struct LevelHandle {     // Handle<Level>::get_rep()
  const Level* p;
  const Level* get_rep() const { return p; }
  const Level* operator->() const { return p; }
};
typedef LevelHandle LevelP;
class Level {
 public:
  int index = 0;
  int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, intLo[3] = {0, 0, 0}, intHi[3] = {0, 0, 0};   // findCellIndexRange / findInteriorCellIndexRange
  void findCellIndexRange(IntVector& l, IntVector& h) const { l = IntVector(lo[0], lo[1], lo[2]); h = IntVector(hi[0], hi[1], hi[2]); }
  void findInteriorCellIndexRange(IntVector& l, IntVector& h) const { l = IntVector(intLo[0], intLo[1], intLo[2]); h = IntVector(intHi[0], intHi[1], intHi[2]); }
  IntVector getExtraCells() const { return IntVector(intLo[0] - lo[0], intLo[1] - lo[1], intLo[2] - lo[2]); }
  bool hasCoarserLevel() const { return coarser != nullptr; }
  bool containsCell(const IntVector& c) const { return c.x() >= lo[0] && c.y() >= lo[1] && c.z() >= lo[2] && c.x() < hi[0] && c.y() < hi[1] && c.z() < hi[2]; }
  IntVector getCellIndex(const Point& pt) const;       // Level.cc:553-557: floor((p - anchor) / dcell)
  double anchor[3] = {0, 0, 0}, dcell[3] = {1, 1, 1};
  int rr[3] = {1, 1, 1};
  const Level* coarser = nullptr;
  int getIndex() const { return index; }
  Point getCellPosition(const IntVector& c) const {           // Level.cc:473-477: m_anchor + m_dcell * v + m_dcell * 0.5
    return Point(anchor[0] + dcell[0] * c.x() + dcell[0] * 0.5, anchor[1] + dcell[1] * c.y() + dcell[1] * 0.5, anchor[2] + dcell[2] * c.z() + dcell[2] * 0.5);
  }
  IntVector mapCellToCoarser(const IntVector& idx) const {    // Level.cc:1285-1307 (level_offset 1)
    IntVector r;
    for (int d = 0; d < 3; ++d) {
      int ratio = idx[d] / rr[d];
      int offset = 0;
      if (idx[d] < 0 && rr[d] > 1) offset = (int)fmod((double)idx[d], (double)rr[d]);
      r[d] = ratio + offset;
    }
    return r;
  }
  LevelHandle getCoarserLevel() const { return LevelHandle{coarser}; }
};

inline IntVector Level::getCellIndex(const Point& pt) const {
  return IntVector((int)std::floor((pt.x() - anchor[0]) / dcell[0]), (int)std::floor((pt.y() - anchor[1]) / dcell[1]), (int)std::floor((pt.z() - anchor[2]) / dcell[2]));
}
inline std::ostream& operator<<(std::ostream& o, const Point& v) { return o << "[" << v.x() << ", " << v.y() << ", " << v.z() << "]"; }

// what Ray::computeExtents touches of the framework (Ray.cc:1279-1358): a patch's cell range, the grid's levels, two reduction variables
class Patch {
 public:
  IntVector lo, hi;
  IntVector getCellLowIndex() const { return lo; }
  IntVector getCellHighIndex() const { return hi; }
  int getID() const { return 0; }
};
struct VarLabel {};
struct minvec_vartype { Vector v; operator Vector() const { return v; } };
struct maxvec_vartype { Vector v; operator Vector() const { return v; } };
inline IntVector roundNearest(const Vector& v) { return IntVector((int)std::lround(v.x()), (int)std::lround(v.y()), (int)std::lround(v.z())); }
class Grid {
 public:
  std::vector<const Level*> levels;
  LevelP getLevel(int L) const { return LevelP{levels[L]}; }
};
class DataWarehouse {
 public:
  const Grid* grid = nullptr;
  Vector roiLo, roiHi;
  const Grid* getGrid() const { return grid; }
  void get(minvec_vartype& v, const VarLabel*) const { v.v = roiLo; }
  void get(maxvec_vartype& v, const VarLabel*) const { v.v = roiHi; }
};
class ProblemSetupException : public std::runtime_error {
 public:
  ProblemSetupException(const std::string& m, const char*, int) : std::runtime_error(m) {}
};
#define DOUT(cond, msg)
static const bool g_ray_dbg = false;

template <class T>
class constCCVariable {      // Array3<T>::operator[](IntVector): x-fastest over the window (Array3.h, Array3Data.h:93,212-220)
 public:
  const T* data = nullptr;
  int lo[3] = {0, 0, 0}, n[3] = {0, 0, 0};
  const T& operator[](const IntVector& c) const {
    return data[(int64_t)(c.x() - lo[0]) + (int64_t)n[0] * ((c.y() - lo[1]) + (int64_t)n[1] * (c.z() - lo[2]))];
  }
};

class InternalError : public std::runtime_error {
 public:
  InternalError(const std::string& m, const char*, int) : std::runtime_error(m) {}
};

// the members the extracted definitions touch (RMCRTCommon.h:65-315, Ray.h:162-330)
class RMCRTCommon {
 public:
  enum DIR { X = 0, Y = 1, Z = 2, NONE = -9 };
  enum FACE { EAST = 0, WEST = 1, NORTH = 2, SOUTH = 3, TOP = 4, BOT = 5, nFACES = 6 };      // RMCRTCommon.h:261
  void raySignStep(double sign[], int cellStep[], const Vector& inv_direction_vector);
  Vector findRayDirection(MTRand& mTwister, const IntVector& origin, const int iRay = -9);
  void ray_Origin(MTRand& mTwister, const Point CC_pos, const Vector dx, const bool useCCRays, Vector& rayOrigin);
  void reflect(double& fs, IntVector& cur, IntVector& prevCell, const double abskg, bool& in_domain, int& step, double& sign, double& ray_direction);
  template <class T>
  void updateSumI(const Level* level, Vector& ray_direction, Vector& ray_origin, const IntVector& origin, const Vector& Dx,
                  constCCVariable<T>& sigmaT4OverPi, constCCVariable<T>& abskg, constCCVariable<int>& celltype, unsigned long int& nRaySteps,
                  double& sumI, MTRand& mTwister);
  bool isDbgCell(IntVector) { return false; }
  int d_flowCell{-1};
  static double d_threshold, d_sigmaScat, d_maxRayLength;
  static bool d_isSeedRandom, d_allowReflect;
};
double RMCRTCommon::d_threshold = 0.01, RMCRTCommon::d_sigmaScat = 0, RMCRTCommon::d_maxRayLength = DBL_MAX;
bool RMCRTCommon::d_isSeedRandom = true, RMCRTCommon::d_allowReflect = true;

class Ray : public RMCRTCommon {
 public:
  inline bool containsCell(const IntVector& low, const IntVector& high, const IntVector& cell, const int& dir);
  template <class T>
  void updateSumI_ML(Vector& ray_direction, Vector& ray_origin, const IntVector& origin, const std::vector<Vector>& Dx, const BBox& domain_BB,
                     const int maxLevels, const Level* fineLevel, const IntVector& fineLevel_ROI_Lo1, const IntVector& fineLevel_ROI_Hi1,
                     std::vector<IntVector>& regionLo, std::vector<IntVector>& regionHi, std::vector<constCCVariable<T>>& sigmaT4OverPi,
                     std::vector<constCCVariable<T>>& abskg, std::vector<constCCVariable<int>>& cellType, unsigned long int& nRaySteps,
                     double& sumI, MTRand& mTwister);
  bool m_use_virtual_ROI{false};
  IntVector m_virtual_ROI{IntVector(0, 0, 0)};
  IntVector d_haloCells{IntVector(-9, -9, -9)};
  // the shells (Ray.h:187-214, 262-330)
  enum ROI_algo { fixed, dynamic, patch_based, boundedRayLength, entireDomain };
  ROI_algo d_ROI_algo{entireDomain};
  Point d_ROI_minPt, d_ROI_maxPt;
  const VarLabel* d_ROI_LoCellLabel{nullptr};
  const VarLabel* d_ROI_HiCellLabel{nullptr};
  std::map<int, IntVector> d_dirIndexOrder, d_dirSignSwap;
  void initFaceTables();
  void computeExtents(LevelP level_0, const Level* fineLevel, const Patch* patch, const int maxLevels, DataWarehouse* new_dw,
                      IntVector& fineLevel_ROI_Lo, IntVector& fineLevel_ROI_Hi, std::vector<IntVector>& regionLo, std::vector<IntVector>& regionHi);
  void rayDirection_cellFace(MTRand& mTwister, const IntVector& origin, const IntVector& indexOrder, const IntVector& signOrder, const int iRay,
                             Vector& directionVector, double& cosTheta);
  void rayLocation_cellFace(MTRand& mTwister, const int face, const Vector Dx, const Point CC_pos, Vector& rayOrigin);
  bool has_a_boundary(const IntVector& c, constCCVariable<int>& celltype, std::vector<int>& boundaryFaces);
};

using std::vector;
using std::cout;
using std::endl;

#include "_ref/ref_rmcrt_extracted.inc"      // the reference's definitions, verbatim

}  // namespace Uintah

//______________________________________________________________________ C interface (same level / parameter structs as the oracle)
#include "rmcrt_oracle.h"

using namespace Uintah;

namespace {
void setStatics(const oracle_params* p) {
  RMCRTCommon::d_threshold = p->threshold;
  RMCRTCommon::d_allowReflect = p->allowReflect != 0;
  RMCRTCommon::d_isSeedRandom = p->isSeedRandom != 0;
  RMCRTCommon::d_maxRayLength = (p->maxRayLength > 0 && p->maxRayLength < 1e300) ? p->maxRayLength : DBL_MAX;
  RMCRTCommon::d_sigmaScat = p->sigmaScat;
}
void fillLevel(const oracle_level& s, int index, Level& L) {
  L.index = index;
  for (int d = 0; d < 3; ++d) {
    L.anchor[d] = s.anchor[d]; L.dcell[d] = s.dx[d]; L.rr[d] = s.refineRatio[d] > 0 ? s.refineRatio[d] : 1;
    L.lo[d] = s.lo[d]; L.hi[d] = s.hi[d]; L.intLo[d] = s.interiorLo[d]; L.intHi[d] = s.interiorHi[d];
  }
}
template <class T> void fillVar(const oracle_level& s, const void* ptr, constCCVariable<T>& v) {
  v.data = static_cast<const T*>(ptr);
  for (int d = 0; d < 3; ++d) { v.lo[d] = s.lo[d]; v.n[d] = s.hi[d] - s.lo[d]; }
}
template <class T>
void updateSumI_T(const oracle_params* p, const oracle_level* lv, const double dir[3], const double rayOrigin[3], const int cell[3], double* sumI,
                  uint64_t* nSteps, uint32_t mtSeed = 5489u) {
  Ray ray;
  ray.d_flowCell = p->flowCell;
  Level L;
  fillLevel(*lv, 0, L);
  constCCVariable<T> abskg, sig;
  constCCVariable<int> ct;
  fillVar<T>(*lv, lv->abskg, abskg);
  fillVar<T>(*lv, lv->sigmaT4, sig);
  fillVar<int>(*lv, lv->cellType, ct);
  Vector d(dir[0], dir[1], dir[2]), o(rayOrigin[0], rayOrigin[1], rayOrigin[2]);
  const Vector Dx(lv->dx[0], lv->dx[1], lv->dx[2]);
  unsigned long int n = 0;
  double s = 0;
  MTRand mt(mtSeed);
  ray.updateSumI<T>(&L, d, o, IntVector(cell[0], cell[1], cell[2]), Dx, sig, abskg, ct, n, s, mt);
  *sumI = s; *nSteps = n;
}
template <class T>
void updateSumI_ML_T(const oracle_params* p, const oracle_level* lv, int maxLevels, const int roiLo[3], const int roiHi[3], const int* regLo,
                     const int* regHi, const double dir[3], const double rayOrigin[3], const int cell[3], double* sumI, uint64_t* nSteps) {
  Ray ray;
  ray.d_flowCell = p->flowCell;
  std::vector<Level> levels(maxLevels);
  std::vector<Vector> Dx(maxLevels);
  std::vector<constCCVariable<T>> abskg(maxLevels), sig(maxLevels);
  std::vector<constCCVariable<int>> ct(maxLevels);
  std::vector<IntVector> regionLo(maxLevels), regionHi(maxLevels);
  for (int l = 0; l < maxLevels; ++l) {
    fillLevel(lv[l], l, levels[l]);
    levels[l].coarser = l > 0 ? &levels[l - 1] : nullptr;
    Dx[l] = Vector(lv[l].dx[0], lv[l].dx[1], lv[l].dx[2]);
    fillVar<T>(lv[l], lv[l].abskg, abskg[l]);
    fillVar<T>(lv[l], lv[l].sigmaT4, sig[l]);
    fillVar<int>(lv[l], lv[l].cellType, ct[l]);
    regionLo[l] = IntVector(regLo[3 * l], regLo[3 * l + 1], regLo[3 * l + 2]);
    regionHi[l] = IntVector(regHi[3 * l], regHi[3 * l + 1], regHi[3 * l + 2]);
  }
  // Ray.cc:1033-1034: level_0->getInteriorSpatialRange(domain_BB)
  const BBox bb(Point(lv[0].anchor[0] + lv[0].dx[0] * lv[0].interiorLo[0], lv[0].anchor[1] + lv[0].dx[1] * lv[0].interiorLo[1], lv[0].anchor[2] + lv[0].dx[2] * lv[0].interiorLo[2]),
                Point(lv[0].anchor[0] + lv[0].dx[0] * lv[0].interiorHi[0], lv[0].anchor[1] + lv[0].dx[1] * lv[0].interiorHi[1], lv[0].anchor[2] + lv[0].dx[2] * lv[0].interiorHi[2]));
  Vector d(dir[0], dir[1], dir[2]), o(rayOrigin[0], rayOrigin[1], rayOrigin[2]);
  unsigned long int n = 0;
  double s = 0;
  MTRand mt;
  ray.updateSumI_ML<T>(d, o, IntVector(cell[0], cell[1], cell[2]), Dx, bb, maxLevels, &levels[maxLevels - 1], IntVector(roiLo[0], roiLo[1], roiLo[2]),
                       IntVector(roiHi[0], roiHi[1], roiHi[2]), regionLo, regionHi, sig, abskg, ct, n, s, mt);
  *sumI = s; *nSteps = n;
}
}  // namespace

extern "C" {

// returns 0, or 1 when the reference threw InternalError (non-physical ray length)
int ref_updateSumI(const oracle_params* p, const oracle_level* lv, const double dir[3], const double rayOrigin[3], const int originCell[3],
                   double* sumI, uint64_t* nRaySteps) {
  setStatics(p);
  try {
    if (lv->isFloat) updateSumI_T<float>(p, lv, dir, rayOrigin, originCell, sumI, nRaySteps);
    else             updateSumI_T<double>(p, lv, dir, rayOrigin, originCell, sumI, nRaySteps);
  } catch (const InternalError&) { return 1; }
  return 0;
}

// the same with the generator seeded (scatter builds draw from it inside the march)
int ref_updateSumI_seeded(const oracle_params* p, const oracle_level* lv, const double dir[3], const double rayOrigin[3], const int originCell[3],
                          uint32_t mtSeed, double* sumI, uint64_t* nRaySteps) {
  setStatics(p);
  try {
    if (lv->isFloat) updateSumI_T<float>(p, lv, dir, rayOrigin, originCell, sumI, nRaySteps, mtSeed);
    else             updateSumI_T<double>(p, lv, dir, rayOrigin, originCell, sumI, nRaySteps, mtSeed);
  } catch (const InternalError&) { return 1; }
  return 0;
}
// 1 when this library was compiled with -DRAY_SCATTER (the reference's --enable-ray-scatter build)
int ref_has_ray_scatter(void) {
#ifdef RAY_SCATTER
  return 1;
#else
  return 0;
#endif
}

int ref_updateSumI_ML(const oracle_params* p, const oracle_level* levels, int maxLevels, const int roiLo[3], const int roiHi[3], const int* regionLo,
                      const int* regionHi, const double dir[3], const double rayOrigin[3], const int originCell[3], double* sumI, uint64_t* nRaySteps) {
  setStatics(p);
  try {
    if (levels[0].isFloat) updateSumI_ML_T<float>(p, levels, maxLevels, roiLo, roiHi, regionLo, regionHi, dir, rayOrigin, originCell, sumI, nRaySteps);
    else                   updateSumI_ML_T<double>(p, levels, maxLevels, roiLo, roiHi, regionLo, regionHi, dir, rayOrigin, originCell, sumI, nRaySteps);
  } catch (const InternalError&) { return 1; }
  return 0;
}

// the ray generation of the divQ loop (Ray.cc:752-760): findRayDirection(mTwister, origin, iRay) then ray_Origin(mTwister, CC_pos, Dx, CCRays, ..)
// for rays iRay = 0..nRays-1 of one cell on one MTRand seeded with mtSeed; dirs / origins: [nRays][3]
void ref_generate_rays(const oracle_params* p, const oracle_level* lv, const int cell[3], int nRays, uint32_t mtSeed, double* dirs, double* origins) {
  setStatics(p);
  Ray ray;
  Level L;
  fillLevel(*lv, 0, L);
  MTRand mt(mtSeed);
  const IntVector c(cell[0], cell[1], cell[2]);
  const Vector Dx(lv->dx[0], lv->dx[1], lv->dx[2]);
  for (int i = 0; i < nRays; ++i) {
    Vector d = ray.findRayDirection(mt, c, i);
    Vector o;
    ray.ray_Origin(mt, L.getCellPosition(c), Dx, p->CCRays != 0, o);
    for (int k = 0; k < 3; ++k) { dirs[3 * i + k] = d[k]; origins[3 * i + k] = o[k]; }
  }
}

}  // extern "C"

// The divQ loop of Ray::rayTrace<T> (Ray.cc:742-779, naive sampling) over cells [patchLo,patchHi), with every per-ray routine it calls
// (findRayDirection, ray_Origin, updateSumI<T>) being the reference's own code above; only the loop shell and the two assignment lines
// 775/778 are restated here.  One MTRand per call, seeded with mtSeed (the per-task generator of Ray.cc:539).  This is what bench.py
// times as the CPU baseline ("kind": "reference") and what tests/test_oracle_ref.py compares oracle_rayTrace with.
template <class T>
static void rayTrace_divQ_T(const oracle_params* p, const oracle_level* lv, const int pLo[3], const int pHi[3], uint32_t mtSeed, double* divQ,
                            double* volq, const int oLo[3], const int oHi[3], uint64_t* nSteps) {
  Ray ray;
  ray.d_flowCell = p->flowCell;
  Level L;
  fillLevel(*lv, 0, L);
  constCCVariable<T> abskg, sig;
  constCCVariable<int> ct;
  fillVar<T>(*lv, lv->abskg, abskg);
  fillVar<T>(*lv, lv->sigmaT4, sig);
  fillVar<int>(*lv, lv->cellType, ct);
  const Vector Dx(lv->dx[0], lv->dx[1], lv->dx[2]);
  MTRand mTwister(mtSeed);
  unsigned long int size = 0;
  const int64_t onx = oHi[0] - oLo[0], ony = oHi[1] - oLo[1];
  const int d_nDivQRays = p->nDivQRays;
  for (int k = pLo[2]; k < pHi[2]; ++k) for (int j = pLo[1]; j < pHi[1]; ++j) for (int i = pLo[0]; i < pHi[0]; ++i) {   // CellIterator: x fastest
    const IntVector origin(i, j, k);
    if (ct[origin] != ray.d_flowCell) continue;
    double sumI = 0;
    Point CC_pos = L.getCellPosition(origin);
    for (int iRay = 0; iRay < d_nDivQRays; iRay++) {
      Vector direction_vector = ray.findRayDirection(mTwister, origin, iRay);
      Vector rayOrigin;
      ray.ray_Origin(mTwister, CC_pos, Dx, p->CCRays != 0, rayOrigin);
      ray.updateSumI<T>(&L, direction_vector, rayOrigin, origin, Dx, sig, abskg, ct, size, sumI, mTwister);
    }
    const int64_t q = (i - oLo[0]) + onx * ((j - oLo[1]) + ony * (int64_t)(k - oLo[2]));
    divQ[q] = -4.0 * M_PI * abskg[origin] * (sig[origin] - (sumI / d_nDivQRays));
    volq[q] = 4.0 * M_PI * (sumI / d_nDivQRays);
  }
  *nSteps = size;
}

extern "C" int ref_rayTrace_divQ(const oracle_params* p, const oracle_level* lv, const int patchLo[3], const int patchHi[3], uint32_t mtSeed,
                                 double* divQ, double* radiationVolq, const int outLo[3], const int outHi[3], uint64_t* nRaySteps) {
  setStatics(p);
  try {
    if (lv->isFloat) rayTrace_divQ_T<float>(p, lv, patchLo, patchHi, mtSeed, divQ, radiationVolq, outLo, outHi, nRaySteps);
    else             rayTrace_divQ_T<double>(p, lv, patchLo, patchHi, mtSeed, divQ, radiationVolq, outLo, outHi, nRaySteps);
  } catch (const InternalError&) { return 1; }
  return 0;
}


// The boundary-flux loop of Ray::rayTrace<T> (Ray.cc:645-730, naive sampling) with every routine it calls being the reference's own
// code: has_a_boundary, rayDirection_cellFace, rayLocation_cellFace (and the face tables of the constructor), updateSumI<T>.  Only the
// loop shell and the two flux lines (712, 716-717) are restated.  boundFlux: Stencil7 AoS (w,e,s,n,b,t,p), p written as 1.0 / 0.0.
template <class T>
static void rayTrace_boundFlux_T(const oracle_params* p, const oracle_level* lv, const int pLo[3], const int pHi[3], uint32_t mtSeed, double* bflux,
                                 const int oLo[3], const int oHi[3], uint64_t* nSteps) {
  Ray ray;
  ray.d_flowCell = p->flowCell;
  ray.initFaceTables();
  Level L;
  fillLevel(*lv, 0, L);
  constCCVariable<T> abskg, sig;
  constCCVariable<int> ct;
  fillVar<T>(*lv, lv->abskg, abskg);
  fillVar<T>(*lv, lv->sigmaT4, sig);
  fillVar<int>(*lv, lv->cellType, ct);
  const Vector Dx(lv->dx[0], lv->dx[1], lv->dx[2]);
  MTRand mTwister(mtSeed);
  unsigned long int size = 0;
  const int64_t onx = oHi[0] - oLo[0], ony = oHi[1] - oLo[1];
  const int d_nFluxRays = p->nFluxRays;
  for (int k = pLo[2]; k < pHi[2]; ++k) for (int j = pLo[1]; j < pHi[1]; ++j) for (int i = pLo[0]; i < pHi[0]; ++i) {
    const IntVector origin(i, j, k);
    if (ct[origin] != ray.d_flowCell) continue;
    std::vector<int> boundaryFaces;
    double* S = bflux + 7 * ((i - oLo[0]) + onx * ((j - oLo[1]) + ony * (int64_t)(k - oLo[2])));
    S[6] = ray.has_a_boundary(origin, ct, boundaryFaces);
    Point CC_pos = L.getCellPosition(origin);
    for (std::vector<int>::iterator it = boundaryFaces.begin(); it < boundaryFaces.end(); it++) {
      int RayFace = *it;
      int UintahFace[6] = {Ray::WEST, Ray::EAST, Ray::SOUTH, Ray::NORTH, Ray::BOT, Ray::TOP};
      double sumI = 0, sumProjI = 0, sumI_prev = 0, sumCosTheta = 0;
      for (int iRay = 0; iRay < d_nFluxRays; iRay++) {
        Vector direction_vector, rayOrigin;
        double cosTheta;
        ray.rayDirection_cellFace(mTwister, origin, ray.d_dirIndexOrder[RayFace], ray.d_dirSignSwap[RayFace], iRay, direction_vector, cosTheta);
        ray.rayLocation_cellFace(mTwister, RayFace, Dx, CC_pos, rayOrigin);
        ray.updateSumI<T>(&L, direction_vector, rayOrigin, origin, Dx, sig, abskg, ct, size, sumI, mTwister);
        sumProjI += cosTheta * (sumI - sumI_prev);
        sumCosTheta += cosTheta;
        sumI_prev = sumI;
      }
      sumProjI = sumProjI * (double)d_nFluxRays / sumCosTheta / 2.0;
      int face = UintahFace[RayFace];
      S[face] = sumProjI * 2 * M_PI / (double)d_nFluxRays;      // Stencil7::operator[]: w e s n b t p (Stencil7.h:38-76)
    }
  }
  *nSteps = size;
}

extern "C" int ref_rayTrace_boundFlux(const oracle_params* p, const oracle_level* lv, const int patchLo[3], const int patchHi[3], uint32_t mtSeed,
                                      double* boundFlux, const int outLo[3], const int outHi[3], uint64_t* nFluxRaySteps) {
  setStatics(p);
  try {
    if (lv->isFloat) rayTrace_boundFlux_T<float>(p, lv, patchLo, patchHi, mtSeed, boundFlux, outLo, outHi, nFluxRaySteps);
    else             rayTrace_boundFlux_T<double>(p, lv, patchLo, patchHi, mtSeed, boundFlux, outLo, outHi, nFluxRaySteps);
  } catch (const InternalError&) { return 1; }
  return 0;
}

// Ray::computeExtents (Ray.cc:1279-1358), the reference's own code: roiAlgo 0 fixed (roiMin / roiMax points), 1 dynamic (ROI_loCell /
// ROI_hiCell reduction values), 2 patch_based.  returns 0, or 2 when the reference threw ProblemSetupException (fixed ROI off the level)
extern "C" int ref_computeExtents(const oracle_level* levels, int maxLevels, const int patchLo[3], const int patchHi[3], const int haloCells[3],
                                  int roiAlgo, const double roiMin[3], const double roiMax[3], int roiLo[3], int roiHi[3], int* regionLo,
                                  int* regionHi) {
  Ray ray;
  std::vector<Level> lv(maxLevels);
  Grid grid;
  for (int l = 0; l < maxLevels; ++l) {
    fillLevel(levels[l], l, lv[l]);
    lv[l].coarser = l > 0 ? &lv[l - 1] : nullptr;
    grid.levels.push_back(&lv[l]);
  }
  DataWarehouse dw;
  dw.grid = &grid;
  if (roiMin && roiMax) {
    dw.roiLo = Vector(roiMin[0], roiMin[1], roiMin[2]); dw.roiHi = Vector(roiMax[0], roiMax[1], roiMax[2]);
    ray.d_ROI_minPt = Point(roiMin[0], roiMin[1], roiMin[2]); ray.d_ROI_maxPt = Point(roiMax[0], roiMax[1], roiMax[2]);
  }
  ray.d_ROI_algo = roiAlgo == 0 ? Ray::fixed : (roiAlgo == 1 ? Ray::dynamic : Ray::patch_based);
  ray.d_haloCells = IntVector(haloCells[0], haloCells[1], haloCells[2]);
  Patch patch;
  patch.lo = IntVector(patchLo[0], patchLo[1], patchLo[2]); patch.hi = IntVector(patchHi[0], patchHi[1], patchHi[2]);
  IntVector rl, rh;
  std::vector<IntVector> regLo(maxLevels), regHi(maxLevels);
  try {
    ray.computeExtents(LevelP{&lv[0]}, &lv[maxLevels - 1], &patch, maxLevels, &dw, rl, rh, regLo, regHi);
  } catch (const ProblemSetupException&) { return 2; }
  for (int d = 0; d < 3; ++d) { roiLo[d] = rl[d]; roiHi[d] = rh[d]; }
  for (int l = 0; l < maxLevels; ++l) for (int d = 0; d < 3; ++d) { regionLo[3 * l + d] = regLo[l][d]; regionHi[3 * l + d] = regHi[l][d]; }
  return 0;
}
