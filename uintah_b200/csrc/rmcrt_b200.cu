// rmcrt_b200.cu -- kernels and C ABI of the B200-native RMCRT solve (see include/rmcrt_b200.h).
//
// Reference paths: /root/reference/src; RMCRT/ = CCA/Components/Models/Radiation/RMCRT/.
//
// Kernel inventory
//   pack_kernel<T,RecT>        abskg/sigmaT4/cellType (Array3 windows) -> packed cell records   (HBM streaming)
//   finalize_kernel<RecT>      validation + "can any non-flow cell emit?" reduction             (HBM streaming)
//   divq_sorted_kernel<RecT,ONION,RESORT>  THE hot kernel (library default): independent Philox rays of a 32-cell tile sorted by expected
//                              length, marched on the blocked half-packed copy of the level (indep.cuh), per-cell sums in ray order
//                              (L1 miss path / L1 data pipe / L2 -> SM bandwidth bound at 30.6 of 32 lanes)
//   blk_build_kernel / roi_blk_build_kernel  the blocked copies (level; Data-Onion region of interest with its marker shell)
//   divq_kernel<Real,RecT,ML>  one cell per lane on the level's own records: FP64 parity march, regression generators, Latin hyper cube,
//                              scattering, bounded rays, 3+ levels, blockedMarch 0          (L2/L1 gather bound)
//   divq_quad / divq_dual_kernel  opt-in lockstep sampling (rayCoherence 1 / 2)                 (issue bound)
//   flux_mask_kernel<RecT>     Ray::has_a_boundary face masks, boundFlux.p                        (streaming)
//   flux_trace_kernel<...>     boundary-flux rays for the compacted (face, cell) items            (gather bound)
//   debug_rays_kernel<...>     caller-specified rays through the production march code
//   sigmaT4 / coarsen / extra-cell fill / box pack / assemble: property pipeline + multi-GPU replication
#include "../../include/rmcrt_b200.h"
#include "march.cuh"
#include "rng.cuh"
#include "indep.cuh"

#include <cuda_runtime.h>

#include <cfloat>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

using namespace rmcrt;

namespace {

constexpr int MAXL = 4;   // levels supported by the Data-Onion kernels (reference decks use 2-3)
// cells of guard in front of a level in its blocked copy (one block: the speculative load past a wall stays inside).  3 cells, which
// puts the interior of a level with one layer of extra cells -- and with it the origin tiles of divq_sorted_kernel -- on block
// boundaries, measured 0.6 % SLOWER than 2 (profiles/r02v49)
constexpr int BLK_PAD = 2;
#ifndef RMCRT_SORT_MIN_RAYS
// below this many rays per cell the default runs the unsorted kernel on the same blocked copy (nothing to sort); from here to 16 rays
// divq_sorted_kernel sorts up to eight tiles together (launch_trace, profiles/r02v40_few_rays.txt)
#define RMCRT_SORT_MIN_RAYS 3
#endif

// per-level status words living on the device, so that a radiation solve needs no host round trip:
//   [0] non-finite records  [1] flow cells on the outer layer  [2] non-flow cells that can emit (k*s != 0)
//   [3] valid: set by finalize_kernel's last block, cleared by every kernel that modifies the level
//   [4] ticket for "last block"      [5] non-finite inputs seen by the pack kernels
//   [6] reflective non-flow cells (abskg < 1): rays can go on after their first wall (chooses between the one- and two-cells-per-lane kernels)
//   [7] bit f: the outer layer of face f (0 x- 1 x+ 2 y- 3 y+ 4 z- 5 z+) has an absorbing cell (abskg >= 0.5); a face without is a mirror
//       for the ray-length estimate of divq_sorted_kernel
constexpr int LSTAT_WORDS = 8;
constexpr int NCOUNTERS = 5;    // [0] divQ steps [1] flux steps [2] divQ cells [3] flux cells [4] radiometer steps

//======================================================================================================
// device-side helpers
//======================================================================================================
struct RayGen {
  uint32_t key0, key1;    // Philox key (seed)
  uint32_t timestep;
  int isSeedRandom;       // 0: MT19937 per-ray reseed (reference regression mode)
  int fixedRandom;        // every draw 0.3
  int ccRays;
  int coherent;           // direction shared by the warp's 32 origin cells
  int lhc;                // rayDirSampleAlgo == LATIN_HYPER_CUBE (Ray.cc:174-179)
  // LHC, regression modes: the Fisher-Yates permutations of RMCRTCommon::randVector (RMCRTCommon.cc:998-1024) depend only on
  // i+j+k (MT19937 reseeded with it) or on nothing (FIXED_RANDOM_NUM); the host tabulates them: perm[(i+j+k) - sumLo][nRays]
  const int* lhcPerm;
  int lhcSumLo, lhcRows;
};

struct TraceArgs {
  OnionGeom<MAXL> G;      // levels [0, maxLevels); the origin cells live on level maxLevels-1
  int multiLevel;         // 0: single-level march on level G.maxLevels-1
  int patchLo[3], patchN[3];
  long long nCells;       // cells in the patch
  long long nTiles;       // ceil(nCells/32)
  int outLo[3], outN[3];
  double* divQ;
  double* volq;
  double* bflux;
  int nRays;
  double threshold;
  int allowReflect;
  int skipZero;
  int fluxFlowGuard;      // 1L: only flow cells get boundary fluxes (Ray.cc:654-656); DO: no guard (Q3)
  double sigmaScat;       // <sigmaScat>: > 0 = isotropic scattering in the single-level march (generic kernels)
  int gate;               // 0: the kernel always runs; 1: only when rays END at their first wall (no reflective wall on the level, or
                          // allowReflect off); 2: only when they can go on -- see launch_gate()
  RayGen rng;
  unsigned long long* counters;   // [0] divQ steps [1] flux steps [2] divQ cells [3] flux cells
  const unsigned* lstat;          // per-level device status words [MAXL][LSTAT_WORDS] (see finalize_kernel)
  int firstLevel;                 // levels [firstLevel, maxLevels) take part in this trace
  int tcx, tcy, tcz;              // divq_sorted_kernel: compact tiles per patch axis (0: tiles = 32 consecutive cells)
  int tsx, tsy, tsz;              // ... of tsx x tsy x tsz = 32 origin cells
  int recBytes;                   // sizeof(record)
  RoiBox roi;                     // Data-Onion fast path: shell-wrapped copy of the fine level's ROI (rec == nullptr: generic march)
  double maxRayLength;            // boundedRayLength ROI (RMCRTCommon.cc:680,684); +inf when unbounded
  int floatProps;                 // the level was packed from float arrays: the FP64 parity march rounds T*T products to float
  int zeroNonFlow;                // write 0 into the non-flow cells of the patch (zeroInit != 0).  With zeroInit == 0 (modifies_divQ:
                                  // getModifiable, Ray.cc:573-588) the reference `continue`s over them and their values survive (746-748)
  // flux items
  const uint32_t* fluxBlk;   // blocked half-packed copy of the level for the boundary-flux rays (nullptr: the level's own records)
  int fluxBx, fluxBxy, fluxLo[3];
  const int* items;       // compacted item ids: face * nCells + patch-linear cell
  const int* nItems;      // device count
};

__device__ __forceinline__ void patch_cell(const TraceArgs& a, long long t, int& i, int& j, int& k) {
  const int nx = a.patchN[0], ny = a.patchN[1];
  const long long row = t / nx;
  i = a.patchLo[0] + (int)(t - row * nx);
  k = a.patchLo[2] + (int)(row / ny);
  j = a.patchLo[1] + (int)(row - (long long)(k - a.patchLo[2]) * ny);
}

__device__ __forceinline__ long long level_index(const LevelGeom& g, int i, int j, int k) {
  return (long long)(i - g.lo[0]) + (long long)g.n[0] * ((long long)(j - g.lo[1]) + (long long)g.n[1] * (long long)(k - g.lo[2]));
}

// Latin-Hyper-Cube stratum of ray `ray` (RMCRTCommon::randVector + Ray.cc:750-752,760-761): a permutation of [0,nRays).
//  * regression modes: row (i+j+k) of the host-built table (exactly the reference's Fisher-Yates shuffle);
//  * <randomSeed> true: a keyed pseudo-random bijection of [0,nRays) (rng.cuh: keyed_permutation), evaluated statelessly per
//    ray; key = one Philox block per (cell or tile, timestep).  Any bijection gives a valid Latin hyper cube; the key
//    makes it differ from cell to cell.
__device__ __forceinline__ int lhc_bin(const RayGen& rg, int i, int j, int k, int ray, int nRays, unsigned long long permId,
                                       uint32_t stream) {
  if (rg.fixedRandom || !rg.isSeedRandom) {
    int row = rg.fixedRandom ? 0 : (i + j + k) - rg.lhcSumLo;
    row = min(max(row, 0), rg.lhcRows - 1);
    return rg.lhcPerm[(long long)row * nRays + ray];
  }
  const u32x4 key = philox4x32_10(u32x4{(uint32_t)permId, 0xFFFFFFFFu, rg.timestep, ((uint32_t)(permId >> 32) << 8) | stream},
                                  rg.key0, rg.key1);
  return (int)keyed_permutation((uint32_t)ray, (uint32_t)nRays, key);
}

// RMCRTCommon::findRayDirection (RMCRTCommon.cc:499-538) + ray_Origin (544-576), FP64 like the reference.
// cellId / dirId : Philox counters (the cell itself; the warp tile's first cell when coherent).
__device__ __forceinline__ void gen_divq_ray(const RayGen& rg, const LevelGeom& g, int i, int j, int k, int ray, int nRays,
                                             unsigned long long cellId, unsigned long long dirId,
                                             double dir[3], double rayDx[3], double* org = nullptr) {
  double u1, u2, o0, o1, o2;
  if (rg.fixedRandom) {
    u1 = u2 = o0 = o1 = o2 = 0.3;
  } else if (!rg.isSeedRandom) {
    uint32_t v[5];
    mt19937_first<5>((uint32_t)((i + j + k) * ray + 1), v);      // RMCRTCommon.cc:504-506
    u1 = u01_open(v[0]); u2 = u01_open(v[1]);
    o0 = u01_closed(v[2]); o1 = u01_closed(v[3]); o2 = u01_closed(v[4]);
  } else {
    const u32x4 a = philox4x32_10(u32x4{(uint32_t)dirId, (uint32_t)ray, rg.timestep,
                                        ((uint32_t)(dirId >> 32) << 8) | STREAM_DIR}, rg.key0, rg.key1);
    const unsigned long long orgId = rg.coherent == 2 ? dirId : cellId;   // 2: the tile also shares the in-cell origin offset
    const u32x4 b = philox4x32_10(u32x4{(uint32_t)orgId, (uint32_t)ray, rg.timestep,
                                        ((uint32_t)(orgId >> 32) << 8) | STREAM_ORIGIN}, rg.key0, rg.key1);
    u1 = u01_open(a.x); u2 = u01_open(a.y);
    o0 = u01_closed(b.x); o1 = u01_closed(b.y); o2 = u01_closed(b.z);
  }
  double pm1, theta;
  if (rg.lhc) {      // Ray::findRayDirectionHyperCube (Ray.cc:1438-1460): bin_i = rand_i[iRay], bin_j = iRay
    const int bin_i = lhc_bin(rg, i, j, k, ray, nRays, dirId, STREAM_LHC_DIVQ);
    pm1 = 2.0 * (u1 + (double)bin_i) / nRays - 1.0;
    theta = 2.0 * M_PI * (u2 + (double)ray) / nRays;
  } else {
    pm1 = 2.0 * u1 - 1.0 + DBL_EPSILON;
    theta = 2.0 * M_PI * u2;
  }
  const double r = sqrt(1.0 - pm1 * pm1);
  double sn, cs;
  sincos(theta, &sn, &cs);
  dir[0] = r * cs; dir[1] = r * sn; dir[2] = pm1;
  const int c[3] = {i, j, k};
  const double off[3] = {o0, o1, o2};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    // absolute positions as the reference forms them, so the FP64 path rounds identically (Level.cc:473-477)
    const double cc = g.anchor[d] + g.dx[d] * c[d] + g.dx[d] * 0.5;
    const double o = rg.ccRays ? cc : (cc - 0.5 * g.dx[d] + off[d] * g.dx[d]);
    rayDx[d] = o - (cc - 0.5 * g.dx[d]);
    if (org) org[d] = o;        // absolute ray origin (RMCRTCommon.cc:565-567), needed by the scattering march
  }
}

// The same ray in FP32 for the FP32 production march with Philox draws: the march consumes float directions and offsets
// anyway, and the FP64 sincos / sqrt of the reference formula cost as much as ~4 DDA steps per ray.  24-bit uniforms,
// direction (r cos, r sin, z) with z = 2u-1 uniform in (-1,1) and phi = 2 pi u uniform (RMCRTCommon.cc:517-530), origin =
// low corner + u*Dx per axis (553-558).
__device__ __forceinline__ void gen_divq_ray_f32(const RayGen& rg, const LevelGeom& g, int ray, unsigned long long cellId,
                                                 unsigned long long dirId, float dir[3], float rayDx[3]) {
  const u32x4 a = philox4x32_10(u32x4{(uint32_t)dirId, (uint32_t)ray, rg.timestep,
                                      ((uint32_t)(dirId >> 32) << 8) | STREAM_DIR}, rg.key0, rg.key1);
  // in-cell origin: rayCoherence 2 keys it on the tile like the direction, so the two spare words of the direction draw serve (24 bits
  // for x, 16 bits each for y and z: one Philox call per ray instead of two); otherwise a draw of its own keyed on the cell
  u32x4 b;
  if (rg.coherent == 2) {
    b = u32x4{a.z, a.w & 0xFFFF0000u, a.w << 16, 0u};
  } else {
    b = philox4x32_10(u32x4{(uint32_t)cellId, (uint32_t)ray, rg.timestep, ((uint32_t)(cellId >> 32) << 8) | STREAM_ORIGIN}, rg.key0, rg.key1);
  }
  const float S = 1.0f / 16777216.0f;                     // 2^-24
  const float z = fmaf((float)(a.x >> 8) + 0.5f, 2.0f * S, -1.0f);      // (-1,1), never +-1
  const float r = sqrtf(fmaxf(0.0f, fmaf(-z, z, 1.0f)));
  float sn, cs;
  sincospif(((float)(a.y >> 8) + 0.5f) * (2.0f * S), &sn, &cs);
  dir[0] = r * cs; dir[1] = r * sn; dir[2] = z;
  const float off[3] = {(float)(b.x >> 8) * S, (float)(b.y >> 8) * S, (float)(b.z >> 8) * S};
#pragma unroll
  for (int d = 0; d < 3; ++d) rayDx[d] = (rg.ccRays ? 0.5f : off[d]) * (float)g.dx[d];
}

// Ray::rayDirection_cellFace (Ray.cc:1362-1390) + rayLocation_cellFace (1465-1513).
// face in push order of has_a_boundary: 0 W(-x) 1 E(+x) 2 S(-y) 3 N(+y) 4 B(-z) 5 T(+z) == Stencil7 slot.
__device__ __forceinline__ void gen_flux_ray(const RayGen& rg, const LevelGeom& g, int i, int j, int k, int face, int ray,
                                             int nRays, unsigned long long cellId, unsigned long long dirId,
                                             double dir[3], double rayDx[3], double& cosTheta, double* org = nullptr) {
  double u1, u2, p0, p1;
  double e1 = 0.3, e2 = 0.3;     // the (0,1) conversions of the same two words: the LHC variant draws randDblExc() (Ray.cc:1413,1416)
  if (rg.fixedRandom) {
    u1 = u2 = p0 = p1 = 0.3;
  } else if (!rg.isSeedRandom) {
    uint32_t v[4];
    mt19937_first<4>((uint32_t)((i + j + k) * ray + 1), v);      // Ray.cc:1371-1373
    u1 = u01_closed(v[0]); u2 = u01_closed(v[1]); p0 = u01_closed(v[2]); p1 = u01_closed(v[3]);
    e1 = u01_open(v[0]); e2 = u01_open(v[1]);
  } else {
    const uint32_t rw = (uint32_t)ray | ((uint32_t)face << 28);
    const u32x4 a = philox4x32_10(u32x4{(uint32_t)dirId, rw, rg.timestep,
                                        ((uint32_t)(dirId >> 32) << 8) | STREAM_FLUX_DIR}, rg.key0, rg.key1);
    const u32x4 b = philox4x32_10(u32x4{(uint32_t)cellId, rw, rg.timestep,
                                        ((uint32_t)(cellId >> 32) << 8) | STREAM_FLUX_ORIGIN}, rg.key0, rg.key1);
    u1 = u01_closed(a.x); u2 = u01_closed(a.y); p0 = u01_closed(b.x); p1 = u01_closed(b.y);
    e1 = u01_open(a.x); e2 = u01_open(a.y);
  }
  double phi, theta;
  if (rg.lhc) {      // Ray::rayDirectionHyperCube_cellFace (Ray.cc:1398-1432): cosTheta stratum first, then phi
    const int bin_i = lhc_bin(rg, i, j, k, ray, nRays, dirId * 8ull + (unsigned long long)face, STREAM_LHC_FLUX);
    theta = acos((e1 + (double)bin_i) / nRays);
    phi = 2.0 * M_PI * (e2 + (double)ray) / nRays;
  } else {
    phi = 2 * M_PI * u1;
    theta = acos(u2);
  }
  cosTheta = cos(theta);
  double sp, cp;
  sincos(phi, &sp, &cp);
  const double st = sin(theta);
  const double tmp[3] = {st * cp, st * sp, cosTheta};
  // Ray.cc:112-125 re-indexed to push order
  const int axisN = face >> 1;            // normal axis
  const bool hiFace = (face & 1) != 0;    // E, N, T
  // indexOrder: x-faces (2,1,0), y-faces (0,2,1), z-faces (0,1,2); sign -1 on the normal axis of E, N, T
  const int io0 = (axisN == 0) ? 2 : 0;
  const int io1 = (axisN == 1) ? 2 : 1;
  const int io2 = (axisN == 0) ? 0 : ((axisN == 1) ? 1 : 2);
  dir[0] = tmp[io0] * ((axisN == 0 && hiFace) ? -1 : 1);
  dir[1] = tmp[io1] * ((axisN == 1 && hiFace) ? -1 : 1);
  dir[2] = tmp[io2] * ((axisN == 2 && hiFace) ? -1 : 1);
  const int c[3] = {i, j, k};
  double q[2] = {p0, p1};
  int qi = 0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double cc = g.anchor[d] + g.dx[d] * c[d] + g.dx[d] * 0.5;
    const double c0 = cc - 0.5 * g.dx[d];
    double o;
    if (d == axisN) o = hiFace ? (c0 + g.dx[d]) : c0;
    else o = c0 + q[qi++] * g.dx[d];
    rayDx[d] = o - (cc - 0.5 * g.dx[d]);
    if (org) org[d] = o;
  }
}

// the generator state a ray's scattering events start from (rng.cuh: ScatRng).  mtPos = MT19937 outputs the ray's own set-up has
// consumed since its reseed (divQ: 2 direction + 3 origin draws, 2 with CCRays; boundary flux: 2 + 2); tag separates the
// Philox counters of divQ rays (0) from the six flux faces.
__device__ __forceinline__ ScatRng scat_rng(const RayGen& rg, int i, int j, int k, int ray, unsigned long long cellId, int mtPos, uint32_t tag) {
  ScatRng r;
  r.mode = rg.fixedRandom ? 2 : (rg.isSeedRandom ? 0 : 1);
  r.key0 = rg.key0; r.key1 = rg.key1;
  r.c0 = (uint32_t)cellId; r.c2 = rg.timestep; r.c3 = ((uint32_t)(cellId >> 32) << 8) | STREAM_SCATTER;
  r.ray = (uint32_t)ray | tag;
  r.draw = 0;
  r.mtSeed = (uint32_t)((i + j + k) * ray + 1);
  r.mtPos = mtPos;
  return r;
}

// The production solve is launched as TWO kernels of which exactly one does the work, chosen on the device from the level's
// finalize status (no host round trip): the two-cells-per-lane kernel wins when every ray ends at its first wall (its joint leg is the
// whole ray), the one-cell kernel (32 registers, twice the occupancy) when reflective walls send rays on for several more legs
// (C5: 1.01e12 vs 0.94e12 steps/s).  Both are correct for any input; an unfinalized level counts as reflective.
__device__ __forceinline__ bool launch_gate(const TraceArgs& a) {
  if (a.gate == 0) return true;
  const unsigned* ls = a.lstat + (a.G.maxLevels - 1) * LSTAT_WORDS;
  const bool raysGoOn = a.allowReflect != 0 && (ls[3] == 0u || ls[6] != 0u);
  return (a.gate == 2) == raysGoOn;
}

// zero-weight legs may be skipped only when every level in play has been finalized and none of its non-flow cells emits
__device__ __forceinline__ bool skip_allowed(const TraceArgs& a) {
  if (!a.skipZero) return false;
  bool ok = true;
  for (int L = a.firstLevel; L < a.G.maxLevels; ++L)
    ok = ok && (a.lstat[L * LSTAT_WORDS + 3] != 0u) && (a.lstat[L * LSTAT_WORDS + 2] == 0u);
  return ok;
}

template <class Real, class D>
__device__ __forceinline__ Real nz(D v) {   // direction component as Real, never exactly zero
  Real r = (Real)v;
  if (r == Real(0)) r = sign_bit(v) ? -Real(1e-30) : Real(1e-30);
  return r;
}

template <class Real, class RecT, bool ML, bool RECORD, bool FAST = false, class D = double>
__device__ __forceinline__ Real trace_one(const TraceArgs& a, bool skipZero, int i, int j, int k, const D dir[3], const D rayDx[3],
                                          unsigned& steps, PathSink* path, ScatRng* srng = nullptr, const double* org = nullptr) {
  const int F = a.G.maxLevels - 1;
  const LevelGeom& g = a.G.lev[F];
  constexpr Real UNSCALE = (Real)RecInfo<RecT>::unscale;   // half-packed records carry sigmaT4/pi scaled by a power of two
  if constexpr (FAST && !ML && !RECORD && std::is_same<Real, float>::value) {
    return UNSCALE * march_single_level_fast(static_cast<const RecT*>(g.rec), (int)level_index(g, i, j, k), g.n[0], g.n[0] * g.n[1],
                                   nz<float, D>(dir[0]), nz<float, D>(dir[1]), nz<float, D>(dir[2]), (float)rayDx[0], (float)rayDx[1],
                                   (float)rayDx[2], (float)g.dx[0], (float)g.dx[1], (float)g.dx[2], (float)a.threshold,
                                   a.allowReflect != 0, skipZero, steps, a.recBytes);
  } else if constexpr (FAST && ML && !RECORD && std::is_same<Real, float>::value) {
    if (a.roi.rec)   // 2-level Data Onion on the per-patch ROI copy
      return UNSCALE * march_onion2_fast<RecT, MAXL>(a.G, a.roi, i, j, k, nz<float, D>(dir[0]), nz<float, D>(dir[1]), nz<float, D>(dir[2]),
                                                     (float)rayDx[0], (float)rayDx[1], (float)rayDx[2], (float)a.threshold,
                                                     a.allowReflect != 0, skipZero, steps, a.recBytes);
    return UNSCALE * march_multi_level<Real, RecT, MAXL, RECORD>(a.G, i, j, k, nz<Real, D>(dir[0]), nz<Real, D>(dir[1]), nz<Real, D>(dir[2]),
                                                       (Real)rayDx[0], (Real)rayDx[1], (Real)rayDx[2], (Real)a.threshold,
                                                       a.allowReflect != 0, skipZero, steps, path, a.floatProps != 0);
  } else
  if (ML) {
    return UNSCALE * march_multi_level<Real, RecT, MAXL, RECORD>(a.G, i, j, k, nz<Real, D>(dir[0]), nz<Real, D>(dir[1]), nz<Real, D>(dir[2]),
                                                       (Real)rayDx[0], (Real)rayDx[1], (Real)rayDx[2], (Real)a.threshold,
                                                       a.allowReflect != 0, skipZero, steps, path, a.floatProps != 0);
  } else {
    if constexpr (std::is_same<D, double>::value) {
      if (srng)     // <sigmaScat> > 0: isotropic scattering (the reference's RAY_SCATTER build), generic march only
        return UNSCALE * march_single_level_scatter<Real, RecT>(static_cast<const RecT*>(g.rec), level_index(g, i, j, k), g.n[0], g.n[0] * g.n[1],
                                                                i, j, k, dir, org, g.anchor, g.dx, (Real)a.threshold, a.allowReflect != 0, steps,
                                                                (Real)(a.maxRayLength * (double)Num<Real>::LEN_SCALE), a.floatProps != 0,
                                                                a.sigmaScat, *srng);
    }
    return UNSCALE * march_single_level<Real, RecT, RECORD>(static_cast<const RecT*>(g.rec), level_index(g, i, j, k), g.n[0],
                                                  g.n[0] * g.n[1], nz<Real, D>(dir[0]), nz<Real, D>(dir[1]), nz<Real, D>(dir[2]),
                                                  (Real)rayDx[0], (Real)rayDx[1], (Real)rayDx[2], (Real)g.dx[0],
                                                  (Real)g.dx[1], (Real)g.dx[2], (Real)a.threshold, a.allowReflect != 0,
                                                  skipZero, steps, path, i, j, k,
                                                  (Real)(a.maxRayLength * (double)Num<Real>::LEN_SCALE), a.floatProps != 0);
  }
}

//======================================================================================================
// THE hot kernel.  One block = one tile of 32 consecutive (x-fastest) origin cells of the patch x W warps;
// warp w marches rays w, w+W, ... of all 32 cells (lane = cell), so the lanes of a warp are parallel rays
// from neighbouring cells (rayCoherence) and every step's 32 loads fall into a handful of 128-B lines.
// Per-lane sums are combined across the W warps through shared memory in a fixed order (deterministic),
// then lane = cell writes divQ / radiationVolq fully coalesced.
//======================================================================================================
// F32GEN: production ray set-up only (Philox draws, FP32 direction/origin); false: the FP64 set-up that also serves the
// regression modes (MT19937 reseed, FIXED_RANDOM_NUM), Latin-Hyper-Cube sampling and the FP64 march.  Two instantiations
// rather than a branch so the hot kernel carries no FP64 sincos / table code in its 32-register budget.
template <class Real, class RecT, bool ML, bool FAST = false, bool F32GEN = false>
// Register budget: the tuned single-level kernel is built for 32 registers/thread (8 x 256-thread blocks per SM = 64
// resident warps).  Measured at 128^3 x 100 rays: 96 regs 5.5e11, 80 regs 6.5e11, 64 regs 8.0e11, 48 regs 8.5e11,
// 40 regs 8.8e11, 32 regs 9.3e11 steps/s -- latency hiding by occupancy beats the spills, which sit in the per-ray
// set-up outside the PTX leg loop.  The generic / Data-Onion / FP64 instantiations keep 64 registers.
#ifndef RMCRT_FAST_MIN_BLOCKS
#define RMCRT_FAST_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(256, ((FAST && !ML) ? RMCRT_FAST_MIN_BLOCKS : 4)) divq_kernel(const __grid_constant__ TraceArgs a) {
  extern __shared__ double red[];     // [W][32]
  if (!launch_gate(a)) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int F = a.G.maxLevels - 1;
  const LevelGeom& g = a.G.lev[F];
  const RecT* recs = static_cast<const RecT*>(g.rec);
  const bool skipZero = skip_allowed(a);
  unsigned long long stepsTotal = 0, cellsTotal = 0;

  for (long long tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x) {
    const long long t = tile * 32 + lane;
    const bool valid = t < a.nCells;
    int i = 0, j = 0, k = 0;
    bool flow = false;
    typename std::conditional<std::is_same<RecT, Rec16>::value, Rec16, Rec8>::type r0{0, 0};
    unsigned long long cellId = 0, dirId = 0;
    if (valid) {
      patch_cell(a, t, i, j, k);
      cellId = (unsigned long long)level_index(g, i, j, k);
      // epilogue properties of the origin cell: full precision (half-packed levels keep a float copy for exactly this)
      if constexpr (std::is_same<RecT, Rec4>::value) r0 = ldrec(static_cast<const Rec8*>(g.recFull) + cellId);
      else r0 = ldrec(recs + cellId);
      flow = !is_wall(r0.k);
    }
    {
      int i0, j0, k0;
      patch_cell(a, tile * 32, i0, j0, k0);
      dirId = a.rng.coherent ? (unsigned long long)level_index(g, i0, j0, k0) : cellId;
    }
    double acc = 0;
    unsigned steps = 0;
    if (flow) {
      if constexpr (F32GEN && std::is_same<Real, float>::value) {
        for (int ray = warp; ray < a.nRays; ray += W) {      // production: Philox draws, FP32 ray set-up
          float dir[3], rayDx[3];
          gen_divq_ray_f32(a.rng, g, ray, cellId, dirId, dir, rayDx);
          acc += (double)trace_one<Real, RecT, ML, false, FAST, float>(a, skipZero, i, j, k, dir, rayDx, steps, nullptr);
        }
      } else {
        for (int ray = warp; ray < a.nRays; ray += W) {      // regression modes (MT19937 reseed, FIXED_RANDOM_NUM) and the FP64 march
          double dir[3], rayDx[3];
          if constexpr (!FAST && !ML) {
            if (a.sigmaScat > 0) {
              double org[3];
              gen_divq_ray(a.rng, g, i, j, k, ray, a.nRays, cellId, dirId, dir, rayDx, org);
              ScatRng srng = scat_rng(a.rng, i, j, k, ray, cellId, a.rng.ccRays ? 2 : 5, 0u);
              acc += (double)trace_one<Real, RecT, ML, false, FAST>(a, skipZero, i, j, k, dir, rayDx, steps, nullptr, &srng, org);
              continue;
            }
          }
          gen_divq_ray(a.rng, g, i, j, k, ray, a.nRays, cellId, dirId, dir, rayDx);
          acc += (double)trace_one<Real, RecT, ML, false, FAST>(a, skipZero, i, j, k, dir, rayDx, steps, nullptr);
        }
      }
    }
    red[warp * 32 + lane] = acc;
    __syncthreads();
    if (warp == 0 && valid) {
      double sumI = 0;
      for (int w = 0; w < W; ++w) sumI += red[w * 32 + lane];
      const long long q = (long long)(i - a.outLo[0]) +
                          (long long)a.outN[0] * ((long long)(j - a.outLo[1]) + (long long)a.outN[1] * (long long)(k - a.outLo[2]));
      if (flow) {
        const double mean = sumI / a.nRays;
        // Ray.cc:775,778 (DO: 1226,1229)
        a.divQ[q] = -4.0 * M_PI * (double)r0.k * ((double)r0.s - mean);
        a.volq[q] = 4.0 * M_PI * mean;
        ++cellsTotal;
      } else if (a.zeroNonFlow) {
        a.divQ[q] = 0.0;   // non-flow cells keep their initial 0 (Ray.cc:581-582,746-748)
        a.volq[q] = 0.0;
      }
    }
    __syncthreads();
    stepsTotal += steps;
  }
  // one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    stepsTotal += __shfl_down_sync(0xffffffffu, stepsTotal, o);
    cellsTotal += __shfl_down_sync(0xffffffffu, cellsTotal, o);
  }
  if (lane == 0) {
    if (stepsTotal) atomicAdd(a.counters + 0, stepsTotal);
    if (cellsTotal) atomicAdd(a.counters + 2, cellsTotal);
  }
}

//======================================================================================================
// Two cells per lane (rayCoherence == 2 and cellsPerLane == 2): one block = a tile of 32 x-consecutive cells x 2 rows; lane = the pair
// of cells (i, j) and (i, j+1); all 64 cells share the direction and the in-cell origin of ray r, so one DDA per lane drives both
// rays (march_single_level_dual).  Production ray set-up only (Philox, FP32), single level.
//======================================================================================================
// pairK: the pair is stacked in z, cells (i,j,k) and (i,j,k+1), instead of in y -- for REJOIN (see march_single_level_dual): which walls
// the two rays reach at the same step depends on the pair's axis, and the reference's mirror decks put their mirrors on x- and y-.
template <class RecT, bool REJOIN>
__global__ void __launch_bounds__(256, 4) divq_dual_kernel(const __grid_constant__ TraceArgs a, int nxc, int njp, int pairK) {
  extern __shared__ double red[];     // [W][64]
  if (!launch_gate(a)) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int F = a.G.maxLevels - 1;
  const LevelGeom& g = a.G.lev[F];
  const RecT* recs = static_cast<const RecT*>(g.rec);
  const bool skipZero = skip_allowed(a);
  constexpr float UNSCALE = RecInfo<RecT>::unscale;
  const int sy = g.n[0], sz = g.n[0] * g.n[1];
  unsigned long long stepsTotal = 0, cellsTotal = 0;
  // njp = rows of pairs in y: (patchN[1] + 1) / 2 pairs stacked in y, or patchN[1] single rows when the pairs are stacked in z
  const int nkp = pairK ? (a.patchN[2] + 1) / 2 : a.patchN[2];
  const int offB = pairK ? sz : sy;
  const long long nTiles = (long long)nxc * njp * nkp;
  for (long long tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    const int xc = (int)(tile % nxc);
    const long long rest = tile / nxc;
    const int jp = (int)(rest % njp), kz = (int)(rest / njp);
    const int i = a.patchLo[0] + xc * 32 + lane, jA = a.patchLo[1] + (pairK ? jp : 2 * jp), k = a.patchLo[2] + (pairK ? 2 * kz : kz);
    const bool validA = (xc * 32 + lane) < a.patchN[0];
    const bool validB = validA && (pairK ? (2 * kz + 1) < a.patchN[2] : (2 * jp + 1) < a.patchN[1]);
    Rec8 rA{0, 0}, rB{0, 0};
    long long idA = 0;
    bool flowA = false, flowB = false;
    if (validA) {
      idA = level_index(g, i, jA, k);
      if constexpr (std::is_same<RecT, Rec4>::value) rA = ldrec(static_cast<const Rec8*>(g.recFull) + idA); else rA = ldrec(recs + idA);
      flowA = !is_wall(rA.k);
    }
    if (validB) {
      if constexpr (std::is_same<RecT, Rec4>::value) rB = ldrec(static_cast<const Rec8*>(g.recFull) + idA + offB); else rB = ldrec(recs + idA + offB);
      flowB = !is_wall(rB.k);
    }
    const unsigned long long dirId = (unsigned long long)level_index(g, a.patchLo[0] + xc * 32, jA, k);
    double accA = 0, accB = 0;
    unsigned stA = 0, stB = 0;
    if (flowA || flowB) {
      const float cx = (float)g.dx[0], cy = (float)g.dx[1], cz = (float)g.dx[2];
      for (int ray = warp; ray < a.nRays; ray += W) {
        float dir[3], rayDx[3];
        // direction: shared by the tile (dirId); in-cell origin: shared by the tile (rayCoherence 2) or by the lane's two cells only (1)
        gen_divq_ray_f32(a.rng, g, ray, (unsigned long long)idA, dirId, dir, rayDx);
        const float d0 = nz<float, float>(dir[0]), d1 = nz<float, float>(dir[1]), d2 = nz<float, float>(dir[2]);
        if (flowA && flowB) {
          float sA, sB;
          march_single_level_dual<REJOIN>(recs, (int)idA, offB, sy, sz, d0, d1, d2, rayDx[0], rayDx[1], rayDx[2], cx, cy, cz, (float)a.threshold,
                                  a.allowReflect != 0, skipZero, stA, stB, a.recBytes, sA, sB);
          accA += (double)(UNSCALE * sA); accB += (double)(UNSCALE * sB);
        } else if (flowA) {
          accA += (double)(UNSCALE * march_single_level_fast(recs, (int)idA, sy, sz, d0, d1, d2, rayDx[0], rayDx[1], rayDx[2], cx, cy, cz,
                                                             (float)a.threshold, a.allowReflect != 0, skipZero, stA, a.recBytes));
        } else {
          accB += (double)(UNSCALE * march_single_level_fast(recs, (int)idA + offB, sy, sz, d0, d1, d2, rayDx[0], rayDx[1], rayDx[2], cx, cy, cz,
                                                             (float)a.threshold, a.allowReflect != 0, skipZero, stB, a.recBytes));
        }
      }
    }
    red[warp * 64 + lane] = accA;
    red[warp * 64 + 32 + lane] = accB;
    __syncthreads();
    if (warp < 2) {           // warp 0 writes row A, warp 1 row B
      const bool valid = warp == 0 ? validA : validB;
      if (valid) {
        double sumI = 0;
        for (int w = 0; w < W; ++w) sumI += red[w * 64 + warp * 32 + lane];
        const int j = jA + (pairK ? 0 : warp), kc = k + (pairK ? warp : 0);
        const long long q = (long long)(i - a.outLo[0]) +
                            (long long)a.outN[0] * ((long long)(j - a.outLo[1]) + (long long)a.outN[1] * (long long)(kc - a.outLo[2]));
        const Rec8 r0 = warp == 0 ? rA : rB;
        if (warp == 0 ? flowA : flowB) {
          const double mean = sumI / a.nRays;
          a.divQ[q] = -4.0 * M_PI * (double)r0.k * ((double)r0.s - mean);
          a.volq[q] = 4.0 * M_PI * mean;
          ++cellsTotal;
        } else if (a.zeroNonFlow) {
          a.divQ[q] = 0.0;
          a.volq[q] = 0.0;
        }
      }
    }
    __syncthreads();
    stepsTotal += (unsigned long long)stA + stB;
  }
  for (int o = 16; o > 0; o >>= 1) {
    stepsTotal += __shfl_down_sync(0xffffffffu, stepsTotal, o);
    cellsTotal += __shfl_down_sync(0xffffffffu, cellsTotal, o);
  }
  if (lane == 0) {
    if (stepsTotal) atomicAdd(a.counters + 0, stepsTotal);
    if (cellsTotal) atomicAdd(a.counters + 2, cellsTotal);
  }
}

//======================================================================================================
// Four cells per lane (cellsPerLane == 4): one block = a tile of 32 x-consecutive cells x 2 rows x 2 planes; lane = the 2 x 2 block of
// cells (i, j..j+1, k..k+1); the four rays of ray index r are exact translates (shared direction and in-cell origin), so one DDA per
// lane drives all four (march_single_level_quad).  Production ray set-up only (Philox, FP32), single level.
//======================================================================================================
#ifndef RMCRT_QUAD_TILE_X
#define RMCRT_QUAD_TILE_X 16      // default origin cells of a warp in x (rmcrt_params.tileX = 0): measured 32: 1.655e12, 16: 1.712e12, 8: 1.686e12 steps/s at 128^3
#endif
#ifndef RMCRT_QUAD_BLOCKS
#define RMCRT_QUAD_BLOCKS 7       // resident blocks per SM the register budget is cut for (72 registers).  Measured, first quad loop: 4: 1.28e12, 5: 1.36e12, 6: 1.46e12, 7: 1.46e12, 8: 1.44e12 steps/s; shipped loop: 6: 1.79e12, 7: 1.83e12, 8: 1.89e12 at 128^3 (float records); 256^3 (half records) 7: 1.947e12, 8: 1.937e12 -- hence one block more for the float-record single-level kernel
#endif
template <class RecT, bool ONION>
__global__ void __launch_bounds__(128, (!ONION && std::is_same<RecT, Rec8>::value) ? RMCRT_QUAD_BLOCKS + 1 : RMCRT_QUAD_BLOCKS) divq_quad_kernel(const __grid_constant__ TraceArgs a, int nxc, int njp, int nkp, int lxBits) {
  extern __shared__ double red[];     // [W][128]
  if (!launch_gate(a)) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int F = a.G.maxLevels - 1;
  const LevelGeom& g = a.G.lev[F];
  const RecT* recs = static_cast<const RecT*>(g.rec);
  const bool skipZero = skip_allowed(a);
  constexpr float UNSCALE = RecInfo<RecT>::unscale;
  const int sy = g.n[0], sz = g.n[0] * g.n[1];
  unsigned long long stepsTotal = 0, cellsTotal = 0;
  const long long nTiles = (long long)nxc * njp * nkp;
  for (long long tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    const int xc = (int)(tile % nxc);
    const long long rest = tile / nxc;
    const int jp = (int)(rest % njp), kp = (int)(rest / njp);
    // lane -> 2 x 2 block: lx of the tile's TX = 1 << lxBits blocks in x, ly of its 32 / TX block rows in y
    const int TX = 1 << lxBits, lx = lane & (TX - 1), ly = lane >> lxBits, RY = 2 * (32 >> lxBits);
    const int i = a.patchLo[0] + xc * TX + lx, jA = a.patchLo[1] + RY * jp + 2 * ly, kA = a.patchLo[2] + 2 * kp;
    const bool vx = (xc * TX + lx) < a.patchN[0] && (RY * jp + 2 * ly) < a.patchN[1];
    const bool vj = (RY * jp + 2 * ly + 1) < a.patchN[1], vk = (2 * kp + 1) < a.patchN[2];
    const bool valid[4] = {vx, vx && vj, vx && vk, vx && vj && vk};
    const int off[4] = {0, sy, sz, sy + sz};
    bool flow[4];
    const long long idA = vx ? level_index(g, i, jA, kA) : level_index(g, a.patchLo[0], a.patchLo[1], kA);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      flow[c] = valid[c] && !is_wall(ldrec(recs + idA + off[c]).k);
    }
    const unsigned long long dirId = (unsigned long long)level_index(g, a.patchLo[0] + xc * TX, a.patchLo[1] + RY * jp, kA);
    double acc[4] = {0, 0, 0, 0};
    unsigned st = 0;          // ray-cell steps of the lane's four cells
    const bool all4 = flow[0] && flow[1] && flow[2] && flow[3];
    if (flow[0] || flow[1] || flow[2] || flow[3]) {
      const float cx = (float)g.dx[0], cy = (float)g.dx[1], cz = (float)g.dx[2];
      for (int ray = warp; ray < a.nRays; ray += W) {
        float dir[3], rayDx[3];
        // direction: shared by the tile (dirId); in-cell origin: shared by the tile (rayCoherence 2) or by the lane's four cells only (1)
        gen_divq_ray_f32(a.rng, g, ray, (unsigned long long)idA, dirId, dir, rayDx);
        const float d0 = nz<float, float>(dir[0]), d1 = nz<float, float>(dir[1]), d2 = nz<float, float>(dir[2]);
        if (all4) {
          float sm[4];
          if constexpr (ONION)    // 2-level Data Onion: joint leg inside the per-patch ROI copy
            march_onion2_quad<RecT, MAXL>(a.G, a.roi, i, jA, kA, d0, d1, d2, rayDx[0], rayDx[1], rayDx[2], (float)a.threshold,
                                          a.allowReflect != 0, skipZero, st, a.recBytes, sm);
          else
            march_single_level_quad(recs, (int)idA, sy, sz, d0, d1, d2, rayDx[0], rayDx[1], rayDx[2], cx, cy, cz, (float)a.threshold,
                                    a.allowReflect != 0, skipZero, st, a.recBytes, sm);
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[c] += (double)(UNSCALE * sm[c]);
        } else {
          // a wall inside the 2 x 2 block (or a partial block at the patch edge): every flow cell of it marches alone, same ray
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            if (!flow[c]) continue;
            unsigned s1 = 0;
            float v;
            if constexpr (ONION)
              v = march_onion2_fast<RecT, MAXL>(a.G, a.roi, i, jA + (c & 1), kA + (c >> 1), d0, d1, d2, rayDx[0], rayDx[1], rayDx[2],
                                                (float)a.threshold, a.allowReflect != 0, skipZero, s1, a.recBytes);
            else
              v = march_single_level_fast(recs, (int)idA + off[c], sy, sz, d0, d1, d2, rayDx[0], rayDx[1], rayDx[2], cx, cy, cz,
                                          (float)a.threshold, a.allowReflect != 0, skipZero, s1, a.recBytes);
            st += s1;
            if (c == 0) acc[0] += (double)(UNSCALE * v);
            else if (c == 1) acc[1] += (double)(UNSCALE * v);
            else if (c == 2) acc[2] += (double)(UNSCALE * v);
            else acc[3] += (double)(UNSCALE * v);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) red[warp * 128 + c * 32 + lane] = acc[c];
    __syncthreads();
    for (int c = warp; c < 4; c += W) {           // warp c writes the row of cell c
      const bool v = c == 0 ? valid[0] : (c == 1 ? valid[1] : (c == 2 ? valid[2] : valid[3]));
      if (!v) continue;
      double sumI = 0;
      for (int w = 0; w < W; ++w) sumI += red[w * 128 + c * 32 + lane];
      const int j = jA + (c & 1), k = kA + (c >> 1);
      const long long q = (long long)(i - a.outLo[0]) +
                          (long long)a.outN[0] * ((long long)(j - a.outLo[1]) + (long long)a.outN[1] * (long long)(k - a.outLo[2]));
      const long long idc = idA + (c & 1 ? sy : 0) + (c & 2 ? sz : 0);
      Rec8 rc;       // the cell's own record at full precision for the epilogue
      if constexpr (std::is_same<RecT, Rec4>::value) rc = ldrec(static_cast<const Rec8*>(g.recFull) + idc); else rc = ldrec(recs + idc);
      const bool fl = c == 0 ? flow[0] : (c == 1 ? flow[1] : (c == 2 ? flow[2] : flow[3]));
      if (fl) {
        const double mean = sumI / a.nRays;
        a.divQ[q] = -4.0 * M_PI * (double)rc.k * ((double)rc.s - mean);      // Ray.cc:775,778
        a.volq[q] = 4.0 * M_PI * mean;
        ++cellsTotal;
      } else if (a.zeroNonFlow) {
        a.divQ[q] = 0.0;
        a.volq[q] = 0.0;
      }
    }
    __syncthreads();
    stepsTotal += st;
  }
  for (int o = 16; o > 0; o >>= 1) {
    stepsTotal += __shfl_down_sync(0xffffffffu, stepsTotal, o);
    cellsTotal += __shfl_down_sync(0xffffffffu, cellsTotal, o);
  }
  if (lane == 0) {
    if (stepsTotal) atomicAdd(a.counters + 0, stepsTotal);
    if (cellsTotal) atomicAdd(a.counters + 2, cellsTotal);
  }
}

//======================================================================================================
// A/B kernel for north_star's "shared-memory staging of the per-patch origin block" (rmcrt_params.originBrickHalo > 0 together with
// genericMarch; default off): the block of W warps that shares a tile of 32 x-consecutive origin cells first copies the tile's
// neighbourhood -- the tile plus `halo` cells on every side, clipped to the level -- into shared memory; the generic march then takes
// every record inside that brick from shared memory and only the rest from global memory.  Same rays, same arithmetic, same results as
// divq_kernel<float, Rec8, false, false, true>; measured in DESIGN.md section 3.
//======================================================================================================
__global__ void __launch_bounds__(256) divq_brick_kernel(const __grid_constant__ TraceArgs a, int halo) {
  extern __shared__ double red[];     // [W][32] doubles, then the brick records
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int F = a.G.maxLevels - 1;
  const LevelGeom& g = a.G.lev[F];
  const Rec8* recs = static_cast<const Rec8*>(g.rec);
  Rec8* bsm = reinterpret_cast<Rec8*>(red + W * 32);
  const bool skipZero = skip_allowed(a);
  unsigned long long stepsTotal = 0, cellsTotal = 0;
  for (long long tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x) {
    const long long t = tile * 32 + lane;
    const bool valid = t < a.nCells;
    int i = 0, j = 0, k = 0;
    bool flow = false;
    Rec8 r0{0, 0};
    unsigned long long cellId = 0;
    if (valid) {
      patch_cell(a, t, i, j, k);
      cellId = (unsigned long long)level_index(g, i, j, k);
      r0 = ldrec(recs + cellId);
      flow = !is_wall(r0.k);
    }
    // brick = bounding box of the tile's cells (a tile may wrap over patch rows) + halo, clipped to the level window
    int i0, j0, k0, i1, j1, k1;
    patch_cell(a, tile * 32, i0, j0, k0);
    patch_cell(a, min(tile * 32 + 31, a.nCells - 1), i1, j1, k1);
    Brick br;
    const int tlo[3] = {(j0 == j1 && k0 == k1) ? i0 : a.patchLo[0], (k0 == k1) ? min(j0, j1) : a.patchLo[1], k0};
    const int thi[3] = {(j0 == j1 && k0 == k1) ? i1 : a.patchLo[0] + a.patchN[0] - 1, (k0 == k1) ? max(j0, j1) : a.patchLo[1] + a.patchN[1] - 1, k1};
    long long nb = 1;
    for (int d = 0; d < 3; ++d) {
      br.lo[d] = max(tlo[d] - halo, g.lo[d]);
      br.n[d] = min(thi[d] + halo + 1, g.lo[d] + g.n[d]) - br.lo[d];
      nb *= br.n[d];
    }
    br.sm = bsm;
    for (long long q = threadIdx.x; q < nb; q += blockDim.x) {
      const int x = (int)(q % br.n[0]);
      const long long rest = q / br.n[0];
      const int y = (int)(rest % br.n[1]), z = (int)(rest / br.n[1]);
      bsm[q] = ldrec(recs + level_index(g, br.lo[0] + x, br.lo[1] + y, br.lo[2] + z));
    }
    __syncthreads();
    double acc = 0;
    unsigned steps = 0;
    if (flow) {
      for (int ray = warp; ray < a.nRays; ray += W) {
        float dir[3], rayDx[3];
        gen_divq_ray_f32(a.rng, g, ray, cellId, cellId, dir, rayDx);
        acc += (double)march_single_level<float, Rec8, false, true>(recs, (long long)cellId, g.n[0], g.n[0] * g.n[1], nz<float, float>(dir[0]),
                                                                    nz<float, float>(dir[1]), nz<float, float>(dir[2]), rayDx[0], rayDx[1], rayDx[2],
                                                                    (float)g.dx[0], (float)g.dx[1], (float)g.dx[2], (float)a.threshold,
                                                                    a.allowReflect != 0, skipZero, steps, nullptr, i, j, k, INFINITY, false, &br);
      }
    }
    red[warp * 32 + lane] = acc;
    __syncthreads();
    if (warp == 0 && valid) {
      double sumI = 0;
      for (int w = 0; w < W; ++w) sumI += red[w * 32 + lane];
      const long long q = (long long)(i - a.outLo[0]) +
                          (long long)a.outN[0] * ((long long)(j - a.outLo[1]) + (long long)a.outN[1] * (long long)(k - a.outLo[2]));
      if (flow) {
        const double mean = sumI / a.nRays;
        a.divQ[q] = -4.0 * M_PI * (double)r0.k * ((double)r0.s - mean);
        a.volq[q] = 4.0 * M_PI * mean;
        ++cellsTotal;
      } else if (a.zeroNonFlow) {
        a.divQ[q] = 0.0;
        a.volq[q] = 0.0;
      }
    }
    __syncthreads();
    stepsTotal += steps;
  }
  for (int o = 16; o > 0; o >>= 1) {
    stepsTotal += __shfl_down_sync(0xffffffffu, stepsTotal, o);
    cellsTotal += __shfl_down_sync(0xffffffffu, cellsTotal, o);
  }
  if (lane == 0) {
    if (stepsTotal) atomicAdd(a.counters + 0, stepsTotal);
    if (cellsTotal) atomicAdd(a.counters + 2, cellsTotal);
  }
}

//======================================================================================================
// An alternative kernel for the DEFAULT sampling (rayCoherence 0, independent rays per (cell, ray, timestep)): indep.cuh; opt-in
// (rmcrt_params.blockedMarch = 1) because it measured SLOWER than the one-cell-per-lane divq_kernel (70 vs 53 ms at 128^3 x 100 rays,
// profiles/r02v3_divq_indep_128cubed_ncu.md): it halves the global sector requests as designed, but the per-lane slot traffic puts as
// many wavefronts on the L1 data pipe as it saves, and the block bookkeeping doubles the instructions per step.
// One WARP = one tile of 8 x 2 x 2 origin cells, lane = cell; a lane traces ALL nDivQRays rays of its cell one after the other (per-lane
// refill), so nothing is reduced across warps and there is no barrier; a block is four independent warps (32 x 2 x 2 cells).  The
// records come from the blocked half-packed copy of the level (BlkGeom), one 256-bit load per 2 x 2 x 2 block entered.
//======================================================================================================
#ifndef RMCRT_INDEP_BLOCKS
#define RMCRT_INDEP_BLOCKS 8       // resident 128-thread blocks per SM the register budget is cut for
#endif
#ifndef RMCRT_INDEP_REFILL
#define RMCRT_INDEP_REFILL 8       // ray set-up runs when this many lanes of the warp wait for a new ray (or nothing else is left to do)
#endif
#ifndef RMCRT_INDEP_BURST
#define RMCRT_INDEP_BURST 8        // DDA steps between two looks at the warp's refill state
#endif
template <class RecT>
__global__ void __launch_bounds__(128, RMCRT_INDEP_BLOCKS) divq_indep_kernel(const __grid_constant__ TraceArgs a, const BlkGeom bg, int nxc, int njp, int nkp,
                                                                             int refill) {
  __shared__ __align__(16) uint32_t slots[2 * 128 * 4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int F = a.G.maxLevels - 1;
  const LevelGeom& g = a.G.lev[F];
  const bool skipZero = skip_allowed(a);
  constexpr float UNSCALE = 1.0f / HALF_S_SCALE;
  // block -> (xc, jp, kp): 32 x 2 x 2 cells; warp -> its 8 x 2 x 2 tile; lane -> cell
  const long long tile = blockIdx.x;
  const int xc = (int)(tile % nxc);
  const long long rest = tile / nxc;
  const int jp = (int)(rest % njp), kp = (int)(rest / njp);
  const int li = xc * 32 + warp * 8 + (lane & 7), lj = 2 * jp + ((lane >> 3) & 1), lk = 2 * kp + (lane >> 4);
  const bool valid = li < a.patchN[0] && lj < a.patchN[1] && lk < a.patchN[2];
  const int i = a.patchLo[0] + li, j = a.patchLo[1] + lj, k = a.patchLo[2] + lk;
  unsigned long long cellId = 0;
  Rec8 r0{0, 0};
  bool flow = false;
  if (valid) {
    cellId = (unsigned long long)level_index(g, i, j, k);
    // epilogue properties of the origin cell at full precision
    if constexpr (std::is_same<RecT, Rec4>::value) r0 = ldrec(static_cast<const Rec8*>(g.recFull) + cellId);
    else r0 = ldrec(static_cast<const Rec8*>(g.rec) + cellId);
    flow = !is_wall(r0.k);
  }
  const int ux = i - bg.lo[0], uy = j - bg.lo[1], uz = k - bg.lo[2];
  const int B0 = (ux >> 1) + bg.bx * (uy >> 1) + bg.bxy * (uz >> 1);
  const unsigned o0 = ((ux & 1) ? OX : 0u) | ((uy & 1) ? OY : 0u) | ((uz & 1) ? OZ : 0u);
  const uint32_t slot = (uint32_t)__cvta_generic_to_shared(slots) + (uint32_t)tid * 16u;
  const uint32_t* __restrict__ blk = bg.blk;
  const float cx = (float)g.dx[0], cy = (float)g.dx[1], cz = (float)g.dx[2];
  const float thr = (float)a.threshold;
  const bool reflect = a.allowReflect != 0;

  IndepRay r;
  double acc = 0;
  unsigned long long steps = 0;
  int next = 0;              // next ray index of this lane
  bool alive = false;
  const int nRays = (flow && 1.0f > thr) ? a.nRays : 0;      // `while (intensity > d_threshold)`, RMCRTCommon.cc:680
  for (;;) {
    const bool want = !alive && next < nRays;
    const unsigned needm = __ballot_sync(0xffffffffu, want);
    const unsigned livem = __ballot_sync(0xffffffffu, alive);
    if ((needm | livem) == 0u) break;
    if (needm != 0u && (__popc(needm) >= refill || livem == 0u)) {
      if (want) {
        float dir[3], rayDx[3];
        gen_divq_ray_f32(a.rng, g, next, cellId, cellId, dir, rayDx);
        ++next;
        indep_start(r, B0, o0, nz<float, float>(dir[0]), nz<float, float>(dir[1]), nz<float, float>(dir[2]), rayDx[0], rayDx[1], rayDx[2],
                    cx, cy, cz, bg, slot);
        alive = true;
      }
    }
#pragma unroll 1
    for (int m = 0; m < RMCRT_INDEP_BURST; ++m) {
      if (alive) {
        if (!indep_step(r, blk, slot, thr, reflect, skipZero)) {
          alive = false;
          acc += (double)(UNSCALE * r.sum);
          steps += (unsigned long long)r.nsteps;
        }
      }
    }
  }
  unsigned long long cells = 0;
  if (valid) {
    const long long q = (long long)(i - a.outLo[0]) +
                        (long long)a.outN[0] * ((long long)(j - a.outLo[1]) + (long long)a.outN[1] * (long long)(k - a.outLo[2]));
    if (flow) {
      const double mean = acc / a.nRays;
      // emission of the origin cell with the SAME fp16 sigmaT4/pi its rays integrate, so that an isothermal enclosure gives divQ == 0
      // whatever the rounding (abskg multiplies both terms: full precision)
      const uint32_t w0 = __ldg(blk + (size_t)B0 * 8 + (((o0 & OX) ? 1 : 0) | ((o0 & OY) ? 2 : 0) | ((o0 & OZ) ? 4 : 0)));
      const double s0 = (double)(UNSCALE * __half2float(__ushort_as_half((unsigned short)(w0 & 0xFFFFu))));
      a.divQ[q] = -4.0 * M_PI * (double)r0.k * (s0 - mean);      // Ray.cc:775,778
      a.volq[q] = 4.0 * M_PI * mean;
      cells = 1;
    } else if (a.zeroNonFlow) {
      a.divQ[q] = 0.0;
      a.volq[q] = 0.0;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    steps += __shfl_down_sync(0xffffffffu, steps, o);
    cells += __shfl_down_sync(0xffffffffu, cells, o);
  }
  if (lane == 0) {
    if (steps) atomicAdd(a.counters + 0, steps);
    if (cells) atomicAdd(a.counters + 2, cells);
  }
}

//======================================================================================================
// blockedMarch = 2: the tile / warp layout of divq_kernel (32 x-consecutive origin cells per tile, lane = cell, warp w traces rays w, w + W,
// ... of its cells, fixed-order cross-warp sum) with march_single_level_blk: the PTX leg loop on the 2 x 2 x 2 blocked half-packed copy
// of the level, plain 4-byte loads (indep.cuh).  Same rays as divq_kernel (Philox keyed on (cell, ray, timestep)).
//======================================================================================================
template <class RecT>
__global__ void __launch_bounds__(256, RMCRT_FAST_MIN_BLOCKS) divq_blk_kernel(const __grid_constant__ TraceArgs a, const BlkGeom bg) {
  extern __shared__ double red[];     // [W][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int F = a.G.maxLevels - 1;
  const LevelGeom& g = a.G.lev[F];
  const bool skipZero = skip_allowed(a);
  constexpr float UNSCALE = 1.0f / HALF_S_SCALE;
  const uint32_t* __restrict__ blk = bg.blk;
  const int sy8 = 8 * bg.bx, sz8 = 8 * bg.bxy;
  const float cx = (float)g.dx[0], cy = (float)g.dx[1], cz = (float)g.dx[2];
  unsigned long long stepsTotal = 0, cellsTotal = 0;

  for (long long tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x) {
    const long long t = tile * 32 + lane;
    const bool valid = t < a.nCells;
    int i = 0, j = 0, k = 0;
    bool flow = false;
    Rec8 r0{0, 0};
    unsigned long long cellId = 0;
    int w0 = 0;
    unsigned par = 0;
    if (valid) {
      patch_cell(a, t, i, j, k);
      cellId = (unsigned long long)level_index(g, i, j, k);
      if constexpr (std::is_same<RecT, Rec4>::value) r0 = ldrec(static_cast<const Rec8*>(g.recFull) + cellId);
      else r0 = ldrec(static_cast<const Rec8*>(g.rec) + cellId);
      flow = !is_wall(r0.k);
      const int ux = i - bg.lo[0], uy = j - bg.lo[1], uz = k - bg.lo[2];
      par = (unsigned)((ux & 1) | ((uy & 1) << 1) | ((uz & 1) << 2));
      w0 = 8 * ((ux >> 1) + bg.bx * (uy >> 1) + bg.bxy * (uz >> 1)) + (int)par;
    }
    double acc = 0;
    unsigned steps = 0;
    if (flow) {
      for (int ray = warp; ray < a.nRays; ray += W) {
        float dir[3], rayDx[3];
        gen_divq_ray_f32(a.rng, g, ray, cellId, cellId, dir, rayDx);
        acc += (double)(UNSCALE * march_single_level_blk(blk, w0, par, sy8, sz8, nz<float, float>(dir[0]), nz<float, float>(dir[1]),
                                                         nz<float, float>(dir[2]), rayDx[0], rayDx[1], rayDx[2], cx, cy, cz,
                                                         (float)a.threshold, a.allowReflect != 0, skipZero, steps));
      }
    }
    red[warp * 32 + lane] = acc;
    __syncthreads();
    if (warp == 0 && valid) {
      double sumI = 0;
      for (int w = 0; w < W; ++w) sumI += red[w * 32 + lane];
      const long long q = (long long)(i - a.outLo[0]) +
                          (long long)a.outN[0] * ((long long)(j - a.outLo[1]) + (long long)a.outN[1] * (long long)(k - a.outLo[2]));
      if (flow) {
        const double mean = sumI / a.nRays;
        // emission of the origin cell with the SAME fp16 sigmaT4/pi its rays integrate (isothermal enclosure => divQ == 0 whatever the
        // rounding); abskg multiplies both terms: full precision.  Ray.cc:775,778
        const uint32_t ww = __ldg(blk + w0);
        const double s0 = (double)(UNSCALE * __half2float(__ushort_as_half((unsigned short)(ww & 0xFFFFu))));
        a.divQ[q] = -4.0 * M_PI * (double)r0.k * (s0 - mean);
        a.volq[q] = 4.0 * M_PI * mean;
        ++cellsTotal;
      } else if (a.zeroNonFlow) {
        a.divQ[q] = 0.0;
        a.volq[q] = 0.0;
      }
    }
    __syncthreads();
    stepsTotal += steps;
  }
  for (int o = 16; o > 0; o >>= 1) {
    stepsTotal += __shfl_down_sync(0xffffffffu, stepsTotal, o);
    cellsTotal += __shfl_down_sync(0xffffffffu, cellsTotal, o);
  }
  if (lane == 0) {
    if (stepsTotal) atomicAdd(a.counters + 0, stepsTotal);
    if (cellsTotal) atomicAdd(a.counters + 2, cellsTotal);
  }
}

//======================================================================================================
// blockedMarch = 3 (the library default): LENGTH-SORTED independent rays on the blocked copy.  Same rays as divq_kernel / divq_blk_kernel
// (Philox keyed on (cell, ray, timestep)), but the lanes of a warp no longer trace ray r of 32 neighbouring cells -- whose lengths differ
// by the full range (13.4 of 32 lanes active, ncu) -- they trace 32 rays of the tile that are expected to END TOGETHER:
//   0. persistent blocks of 8 warps take tiles of 4 x 4 x 2 origin cells off a counter (compact_tiles);
//   1. every (cell, ray) of the tile's 32 cells x RP rays gets a 9-bit key from the number of DDA steps to the far side of the level along
//      its direction (closed form from the direction and the cell index, ~15 FP instructions; walls inside the domain and the intensity
//      threshold make it an estimate -- correctness never depends on it);
//   2. a counting sort (histogram in shared memory, scan, scatter of the ray ids into global scratch) orders the rays, longest first;
//   3. warps take chunks of 32 consecutive rays off that list (dynamic hand-out: longest chunks first, so the warps of a block end
//      together), regenerate each ray from its Philox counter and march it (march_single_level_blk / march_onion_blk);
//   4. each ray's sum is written to res[ray][cell] (global scratch) and the rays of a cell are added in RAY ORDER in FP64 -- the result
//      does not depend on the sort, the hand-out, the rays per pass or the patch layout, bit for bit.
//======================================================================================================
#ifndef RMCRT_SORTED_BLOCKS
#define RMCRT_SORTED_BLOCKS 4
#endif
#ifndef RMCRT_KEY_VARIANT
#define RMCRT_KEY_VARIANT 2
#endif
#ifndef RMCRT_RESORT_ROUNDS
#define RMCRT_RESORT_ROUNDS 4      // RESORT: rounds of one leg each before whatever is left runs to its end
#endif
constexpr int RESORT_ROUNDS = RMCRT_RESORT_ROUNDS;
constexpr int SORT_BINS = 512;       // 9-bit keys: one length, or two lengths (first leg: 4 bits | rest: 5 bits)
// per-block scratch in GLOBAL memory (L2-resident, accessed with .cg so it never takes L1 lines from the records): res[RP][32] f32 |
// ids[RP * 32] u16 | keys[RP * 32] u16.  Shared memory would do, but every KB of it comes out of the SM's L1, and the march lives on the
// L1 keeping a sector until the ray's next step (23 KB per block = 187 KB per SM measured: L1 hit rate 46 % -> 8 %, solve 47 -> 59 ms).
// RESORT adds state[RP * 32] x 48 B: the rays parked after their first leg.
__host__ __device__ inline size_t sorted_scratch_bytes(int RP, bool resort) { return (size_t)RP * 32 * (resort ? 56 : 8); }
// dynamic shared memory: Philox counter (8 B), word index, threshold length (4 B each) and cell index (12 B) of the 32 cells of each of G tiles
inline size_t sorted_cell_smem(int G) { return (size_t)G * 32 * 28; }
// ONION: the Data Onion (Ray::rayTrace_dataOnion, 2 to 4 levels): origin cells on the fine level, rays march the blocked copy of the
// patch's region of interest and go on, level by level, in the blocked region boxes of the coarser levels (march_onion_blk); bg = the
// ROI box, ob = all boxes.
// RESORT (single level, rays that go on after their first wall: reflective walls, or hot black ones whose zero-weight second leg the
// reference marches): the lanes of a warp wait for each other at the end of every leg, so a sort by whole-ray length does not keep them
// busy (C5 at 256^3: 18 of 32 lanes).  Here the FIRST legs are sorted and marched as usual; a ray that goes on parks its state (48 B) and
// gets a new key from where it stands -- a weightless ray ends at its next wall, so its one remaining leg sorts exactly -- and the parked
// rays are sorted and finished in a second pass over the tile.  The two instantiations are launched back to back and the level's status
// words decide on the device which one does the work (a.gate: 1 = rays end at their first wall, 2 = they can go on).
// MULTI: G > 1 tiles per block (few rays per cell); otherwise G is the constant 1 and the group arithmetic folds away
template <class RecT, bool ONION, bool RESORT = false, bool MULTI = false>
__global__ void __launch_bounds__(256, RMCRT_SORTED_BLOCKS) divq_sorted_kernel(const __grid_constant__ TraceArgs a, const BlkGeom bg,
                                                                               const __grid_constant__ OnionBlk ob, int RP, int Garg,
                                                                               unsigned char* __restrict__ scratch, int* __restrict__ tileCounter) {
  const int G = MULTI ? Garg : 1;
  // G: 32-cell tiles sorted and marched TOGETHER (1 at the BASELINE's 100 rays per cell; up to 8 when a cell has few rays, so that a
  // block still has a few dozen 32-ray chunks for its eight warps -- a block is as slow as its longest chunk).  RP = capacity of a pass in
  // (tile, ray) pairs per lane: the "virtual rays" v = group * nr + ray of the lists below.
  // 2.9 KB of shared memory per block at G = 1: four resident blocks stay inside the 16-KB carve-out, the rest of the SM's 256 KB is L1
  // (1024 bins = 4.9 KB per block move the carve-out to 32 KB: 3 % slower at 128^3)
  __shared__ int hist[SORT_BINS];
  double* part = reinterpret_cast<double*>(hist);      // [8][32] = 2 KB, written in phase 4 when hist is dead
  extern __shared__ __align__(16) unsigned char cellSmem[];      // per group of 32 cells:
  unsigned long long* cellIds = reinterpret_cast<unsigned long long*>(cellSmem);        // [G * 32] Philox counter of the cell
  int* w0s = reinterpret_cast<int*>(cellIds + G * 32);           // [G * 32] word index of the cell, -1: not traced
  float* tThrS = reinterpret_cast<float*>(w0s + G * 32);         // [G * 32] path length at which the intensity threshold ends a ray
  int* ijk = reinterpret_cast<int*>(tThrS + G * 32);             // [G * 32][3] cell index relative to the level's low corner
  __shared__ int ctl[4];             // [0] next chunk [1] list length [2] tile group
  __shared__ int warpTot[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int W = 8;
  unsigned char* my = scratch + (size_t)blockIdx.x * sorted_scratch_bytes(RP, RESORT);
  uint4* parked = reinterpret_cast<uint4*>(my + (size_t)RP * 256);                     // RESORT: [RP * 32][3]
  float* res = reinterpret_cast<float*>(my);                                           // [RP][32]
  unsigned short* ids = reinterpret_cast<unsigned short*>(my + (size_t)RP * 128);      // [RP * 32]
  unsigned short* keys = reinterpret_cast<unsigned short*>(my + (size_t)RP * 192);     // [RP * 32]
  const int F = a.G.maxLevels - 1;
  const LevelGeom& g = a.G.lev[F];
  const bool skipZero = skip_allowed(a);
  constexpr float UNSCALE = 1.0f / HALF_S_SCALE;
  const uint32_t* __restrict__ blk = bg.blk;
  const int sy8 = 8 * bg.bx, sz8 = 8 * bg.bxy;
  const float cx = (float)g.dx[0], cy = (float)g.dx[1], cz = (float)g.dx[2];
  const float icx = 1.0f / cx, icy = 1.0f / cy, icz = 1.0f / cz;
  // faces that do not end a ray: allowReflect and no absorbing cell on the face (finalize_kernel); an unfinalized level has none
  const unsigned* ls = a.lstat + (ONION ? 0 : F) * LSTAT_WORDS;      // (Data Onion: rays end on level 0)
  const unsigned mirrors = (a.allowReflect && ls[3] != 0u && ls[6] != 0u) ? (~ls[7] & 63u) : 0u;
  const float logThr = a.threshold > 0 ? (float)log(1.0 / a.threshold) : 1e30f;
  // Sort key.  The lanes of a warp wait for each other at the end of every LEG (a wall, the ROI exit), so rays that are to run together
  // must agree leg by leg, not just in total.  Single level: the key is that of the FIRST leg, (z-hemisphere of the direction: 1 bit |
  // length: 6 bits | x and y signs: 2 bits) -- the 32 rays of a chunk also head the same way, and so do the chunks the warps of a block
  // march at one time, so that more of their sectors near the origin are shared (+0.8 % at 128^3 over a 9-bit length with the octant in
  // the low bits, profiles/r02v22; another +1.1 % with the hemisphere on top, profiles/r02v57_sort_key.txt) -- and rays that go on
  // are sorted again where they stand (RESORT).  Data Onion: (fine steps inside the ROI: 4 bits | coarse steps: 5 bits) -- a chunk of 32
  // out of a tile's ~3200 rays then spans about a sixteenth of the range of the first part and a sixth of the second.
  // goOn: bounces the estimate follows for a ray that goes on after a wall; 0 = rays end at their first wall
  const int goOn = (a.allowReflect && !ONION && (mirrors != 0u || !skipZero)) ? 6 : 0;
  if (a.gate != 0 && (a.gate == 2) != (goOn != 0)) return;
  const bool twoPart = ONION;
  const float nSum = (float)(g.n[0] + g.n[1] + g.n[2]);
  float scale1 = 63.0f / nSum, scale2 = 0.0f;
  if constexpr (ONION) {     // fine steps inside the ROI | steps on the coarser levels
    float nCoarse = 0.0f;
    for (int l = 0; l < F; ++l) nCoarse += (float)((a.G.lev[l].regHi[0] - a.G.lev[l].regLo[0]) + (a.G.lev[l].regHi[1] - a.G.lev[l].regLo[1]) +
                                                   (a.G.lev[l].regHi[2] - a.G.lev[l].regLo[2]) + 6);
    const float nRoi = (float)((a.G.roiHi[0] - a.G.roiLo[0]) + (a.G.roiHi[1] - a.G.roiLo[1]) + (a.G.roiHi[2] - a.G.roiLo[2]));
    scale1 = 15.0f / (0.6f * nRoi);
    scale2 = 31.0f / (0.6f * nCoarse);
  }
  // exclusive scan over the bins in DESCENDING key order (thread t owns PER consecutive bins): hist[bin] = start of the bin in the
  // sorted list, ctl[1] = list length.  All threads; hist must be complete (a barrier before), is ready after the call.
  auto scan_bins = [&]() {
    constexpr int PER = SORT_BINS / 256;
    int hq[PER], h = 0;
#pragma unroll
    for (int u = 0; u < PER; ++u) { hq[u] = hist[SORT_BINS - 1 - (tid * PER + u)]; h += hq[u]; }
    int incl = h;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warpTot[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += warpTot[w];
    int start = base + incl - h;
#pragma unroll
    for (int u = 0; u < PER; ++u) { hist[SORT_BINS - 1 - (tid * PER + u)] = start; start += hq[u]; }
    if (tid == 255) ctl[1] = base + incl;
    __syncthreads();
  };
  unsigned long long stepsTotal = 0, cellsTotal = 0;
#ifdef RMCRT_PHASE_CLOCKS
  long long tMarch = 0, tAll = clock64(), tMark = 0;
#endif

  for (;;) {
    // tiles are handed out dynamically (rays near the centre of the domain are several times longer than those in its corners)
    if (tid == 0) ctl[2] = atomicAdd(tileCounter, 1);
    __syncthreads();
    const long long tile0 = (long long)ctl[2] * G;
    if (tile0 >= a.nTiles) break;
    // ---- phase 0: the cells of the G tiles (warp g < G: tile tile0 + g, lane = cell)
    int i = 0, j = 0, k = 0;
    bool valid = false, flow = false;
    Rec8 r0{0, 0};
    int w0 = -1;
    if (warp < G) {
      const long long tile = tile0 + warp;
      const long long t = tile * 32 + lane;
      valid = tile < a.nTiles && t < a.nCells;
      if (a.tcx) {      // compact tiles: 8 x 2 x 2 cells, the rays of a tile share more sectors near their origin
        const int tz_ = (int)(tile / ((long long)a.tcx * a.tcy)), ty_ = (int)((tile - (long long)tz_ * a.tcx * a.tcy) / a.tcx);
        const int tx_ = (int)(tile - ((long long)tz_ * a.tcy + ty_) * a.tcx);
        const int li = a.tsx * tx_ + lane % a.tsx, lj = a.tsy * ty_ + (lane / a.tsx) % a.tsy, lk = a.tsz * tz_ + lane / (a.tsx * a.tsy);
        valid = tile < a.nTiles && li < a.patchN[0] && lj < a.patchN[1] && lk < a.patchN[2];
        i = a.patchLo[0] + li; j = a.patchLo[1] + lj; k = a.patchLo[2] + lk;
      }
      unsigned long long cellId = 0;
      if (valid) {
        if (!a.tcx) patch_cell(a, t, i, j, k);
        cellId = (unsigned long long)level_index(g, i, j, k);
        {      // (read once per cell: streaming, so that 8 B per cell of float records do not push the blocked copy out of the L2)
          const float2 v = __ldcs(reinterpret_cast<const float2*>(std::is_same<RecT, Rec4>::value ? g.recFull : g.rec) + cellId);
          r0 = Rec8{v.x, v.y};
        }
        flow = !is_wall(r0.k);
        const int ux = i - bg.lo[0], uy = j - bg.lo[1], uz = k - bg.lo[2];
        if (flow) w0 = 8 * ((ux >> 1) + bg.bx * (uy >> 1) + bg.bxy * (uz >> 1)) + ((ux & 1) | ((uy & 1) << 1) | ((uz & 1) << 2));
      }
      const int gl = warp * 32 + lane;
      cellIds[gl] = cellId;
      w0s[gl] = w0;
      ijk[gl * 3 + 0] = i - g.lo[0]; ijk[gl * 3 + 1] = j - g.lo[1]; ijk[gl * 3 + 2] = k - g.lo[2];
      // ... or the intensity threshold ends a ray first: optical depth ln(1 / threshold) at the origin cell's abskg (RMCRTCommon.cc:680)
      tThrS[gl] = w0 >= 0 ? logThr / fmaxf(fabsf(__half2float(__ushort_as_half((unsigned short)(__ldg(blk + w0) >> 16)))), 1e-20f) : 1e30f;
    }
    double acc = 0;       // warps < G: running sum over the passes
    unsigned steps = 0;
    const int RPr = RP / G;      // rays of a cell per pass
    for (int ray0 = 0; ray0 < a.nRays; ray0 += RPr) {
      const int nr = min(RPr, a.nRays - ray0), V = G * nr;
      for (int q4 = tid; q4 < SORT_BINS; q4 += 256) hist[q4] = 0;
      if (tid == 0) ctl[0] = 0;
      __syncthreads();
      // ---- phase 1: keys + histogram (thread -> cell = lane of every group; (group, ray) pairs warp, warp + W, ...)
      {
          for (int v = warp; v < V; v += W) {
            const int gq = G == 1 ? 0 : v / nr, r = v - gq * nr, gc = gq * 32 + lane;
            if (w0s[gc] < 0) continue;
            const unsigned long long cid = cellIds[gc];
            const float ci = (float)ijk[gc * 3 + 0], cj = (float)ijk[gc * 3 + 1], ck = (float)ijk[gc * 3 + 2];
            const float tThr = tThrS[gc];
            float dir[3], rayDx[3];
            gen_divq_ray_f32(a.rng, g, ray0 + r, cid, cid, dir, rayDx);
            // cell faces to the outer layer of the level along each axis, from the cell centre
            const float ax = fabsf(dir[0]) * icx + 1e-30f, ay = fabsf(dir[1]) * icy + 1e-30f, az = fabsf(dir[2]) * icz + 1e-30f;
            const bool px = dir[0] >= 0.0f, py = dir[1] >= 0.0f, pz = dir[2] >= 0.0f;
            const float rate = ax + ay + az;             // cell faces crossed per unit length
            float e1, e2 = 0.0f;                         // steps of the first leg (Data Onion: inside the ROI), of the rest
            const float nx_ = (px ? (float)(g.n[0] - 1) - ci : ci) - 0.5f;
            const float ny_ = (py ? (float)(g.n[1] - 1) - cj : cj) - 0.5f;
            const float nz_ = (pz ? (float)(g.n[2] - 1) - ck : ck) - 0.5f;
            const float tExit = fminf(fminf(fminf(nx_ / ax, ny_ / ay), nz_ / az), tThr);
            e1 = tExit * rate;
            if constexpr (ONION) {
              // fine steps up to the ROI face, then level by level to the face of that level's region (cells R times larger; level 0:
              // to the wall)
              const float rx_ = (px ? (float)(a.G.roiHi[0] - g.lo[0]) - ci : ci - (float)(a.G.roiLo[0] - g.lo[0]) + 1.0f) - 0.5f;
              const float ry_ = (py ? (float)(a.G.roiHi[1] - g.lo[1]) - cj : cj - (float)(a.G.roiLo[1] - g.lo[1]) + 1.0f) - 0.5f;
              const float rz_ = (pz ? (float)(a.G.roiHi[2] - g.lo[2]) - ck : ck - (float)(a.G.roiLo[2] - g.lo[2]) + 1.0f) - 0.5f;
              float tPrev = fminf(fminf(fminf(rx_ / ax, ry_ / ay), rz_ / az), tExit);
              e1 = tPrev * rate;
              float R0 = 1.0f, R1 = 1.0f, R2 = 1.0f;       // fine cells per cell of level l
              for (int l = F - 1; l >= 0; --l) {
                const LevelGeom& gu = a.G.lev[l + 1];
                const LevelGeom& gl = a.G.lev[l];
                R0 *= (float)gu.rr[0]; R1 *= (float)gu.rr[1]; R2 *= (float)gu.rr[2];
                float tl = tExit;
                if (l > 0) {      // faces of the region of level l, in fine cells from the fine level's low corner (same anchor)
                  const float q0 = px ? (float)gl.regHi[0] * R0 - (float)g.lo[0] - (ci + 0.5f) : (ci + 0.5f) - ((float)gl.regLo[0] * R0 - (float)g.lo[0]);
                  const float q1 = py ? (float)gl.regHi[1] * R1 - (float)g.lo[1] - (cj + 0.5f) : (cj + 0.5f) - ((float)gl.regLo[1] * R1 - (float)g.lo[1]);
                  const float q2 = pz ? (float)gl.regHi[2] * R2 - (float)g.lo[2] - (ck + 0.5f) : (ck + 0.5f) - ((float)gl.regLo[2] * R2 - (float)g.lo[2]);
                  tl = fminf(fminf(fminf(q0 / ax, q1 / ay), q2 / az), tExit);
                }
                e2 += fmaxf(tl - tPrev, 0.0f) * (ax / R0 + ay / R1 + az / R2);
                tPrev = fmaxf(tPrev, tl);
              }
            }
            int key;
            if (twoPart) key = (min(15, (int)(e1 * scale1 + 0.5f)) << 5) | min(31, (int)(e2 * scale2 + 0.5f));
#if RMCRT_KEY_VARIANT == 1      // A/B: octant-major (the chunks a block marches at one time head the same way)
            else key = (((px ? 1 : 0) | (py ? 2 : 0) | (pz ? 4 : 0)) << 6) | min(63, (int)(e1 * scale1 + 0.5f));
#elif RMCRT_KEY_VARIANT == 2    // shipped: hemisphere-major
            else key = ((pz ? 1 : 0) << 8) | (min(63, (int)(e1 * scale1 + 0.5f)) << 2) | (px ? 1 : 0) | (py ? 2 : 0);
#else                           // A/B: (length | octant)
            else key = (min(63, (int)(e1 * scale1 + 0.5f)) << 3) | (px ? 1 : 0) | (py ? 2 : 0) | (pz ? 4 : 0);
#endif
            atomicAdd(&hist[key], 1);
            __stcg(keys + v * 32 + lane, (unsigned short)key);
          }
      }
      __syncthreads();
      // ---- phase 2: start of every bin in the sorted list (descending keys), then scatter
      {
        scan_bins();
        for (int v = warp; v < V; v += W) {
          if (w0s[(G == 1 ? 0 : v / nr) * 32 + lane] < 0) continue;
          const int pos = atomicAdd(&hist[__ldcg(keys + v * 32 + lane)], 1);
          __stcg(ids + pos, (unsigned short)(v * 32 + lane));
        }
      }
      __syncthreads();
      if constexpr (RESORT) {
        // ---- phase 3, in ROUNDS of one leg: round 0 marches the first legs (the list phase 2 sorted); a ray that goes on is parked with a
        // key for what is left of it; the parked rays are sorted and marched for their next leg, and so on until none is left (a weightless
        // ray ends at its next wall; with mirrors a few per cent of the rays see a fourth round)
        const float thr = (float)a.threshold;
        typedef Num<float> N;
        // park: state of a ray standing reflected in d.idx + the key of its remaining legs, followed through the empty box from the cell
        // it stands in (decoded from its word index)
        auto park = [&](const BlkRay& ry, int id, const float (&dir)[3]) {
          parked[id * 3 + 0] = make_uint4(__float_as_uint(ry.d.tx), __float_as_uint(ry.d.ty), __float_as_uint(ry.d.tz), (unsigned)ry.d.idx);
          parked[id * 3 + 1] = make_uint4((unsigned)ry.b.dlx, (unsigned)ry.b.dly, (unsigned)ry.b.dlz, __float_as_uint(ry.q.A));
          parked[id * 3 + 2] = make_uint4(__float_as_uint(ry.q.ePrev), __float_as_uint(ry.q.nsteps), __float_as_uint(ry.sum), __float_as_uint(ry.fs));
          int B = ry.d.idx >> 3;
          const int bz_ = B / bg.bxy; B -= bz_ * bg.bxy;
          const int by_ = B / bg.bx, bx_ = B - by_ * bg.bx;
          float p0 = (float)(2 * bx_ + (ry.d.idx & 1) + bg.lo[0] - g.lo[0]) + 0.5f;
          float p1 = (float)(2 * by_ + ((ry.d.idx >> 1) & 1) + bg.lo[1] - g.lo[1]) + 0.5f;
          float p2 = (float)(2 * bz_ + ((ry.d.idx >> 2) & 1) + bg.lo[2] - g.lo[2]) + 0.5f;
          const float ax = fabsf(dir[0]) * icx + 1e-30f, ay = fabsf(dir[1]) * icy + 1e-30f, az = fabsf(dir[2]) * icz + 1e-30f;
          bool s0 = ry.b.dlx > 0, s1 = ry.b.dly > 0, s2 = ry.b.dlz > 0, dead = ry.fs == 0.0f;
          const bool deadAtStart = dead;
          float tTot = 0.0f, tFirst = 0.0f;
          for (int bnc = 0; bnc < goOn; ++bnc) {
            const float t0 = (s0 ? (float)(g.n[0] - 1) - p0 : p0 - 1.0f) / ax;
            const float t1 = (s1 ? (float)(g.n[1] - 1) - p1 : p1 - 1.0f) / ay;
            const float t2 = (s2 ? (float)(g.n[2] - 1) - p2 : p2 - 1.0f) / az;
            const float t = fmaxf(fminf(fminf(t0, t1), t2), 0.0f);
            const int axh = (t0 <= t1 && t0 <= t2) ? 0 : (t1 <= t2 ? 1 : 2);
            tTot += t;
            if (bnc == 0) tFirst = tTot;
            p0 += (s0 ? t : -t) * ax; p1 += (s1 ? t : -t) * ay; p2 += (s2 ? t : -t) * az;
            const unsigned face = 2u * axh + ((axh == 0 ? s0 : (axh == 1 ? s1 : s2)) ? 1u : 0u);
            if (dead) break;                      // a weightless ray ends at its next wall of any kind (intensity 0)
            if (!((mirrors >> face) & 1u)) {      // an absorbing face: the weight is gone, the ray is not
              if (skipZero) break;
              dead = true;
            }
            if (axh == 0) s0 = !s0; else if (axh == 1) s1 = !s1; else s2 = !s2;
          }
          const float rate = ax + ay + az;
          // weightless rays: ONE leg, 8 bits of length, upper half of the key space; the others (next leg: 3 bits | rest: 5 bits)
          unsigned short key2;
          if (deadAtStart) key2 = (unsigned short)(256 + min(255, (int)(tTot * rate * (255.0f / nSum) + 0.5f)));
          else key2 = (unsigned short)((min(7, (int)(tFirst * rate * (7.0f / (0.6f * nSum)) + 0.5f)) << 5) |
                                       min(31, (int)((tTot - tFirst) * rate * (31.0f / (1.2f * nSum)) + 0.5f)));
          atomicAdd(&hist[key2], 1);
          __stcg(keys + id, key2);
        };
        for (int round = 0;; ++round) {
          // (hist: the bin positions of the list being marched are no longer needed; it collects the keys of the rays parked in this round)
          const int nList = ctl[1];
          __syncthreads();
          for (int q4 = tid; q4 < SORT_BINS; q4 += 256) hist[q4] = 0;
          __syncthreads();
          for (;;) {
            int chunk = 0;
            if (lane == 0) chunk = atomicAdd(&ctl[0], 1);
            chunk = __shfl_sync(0xffffffffu, chunk, 0);
            if (chunk * 32 >= nList) break;
            const int p = chunk * 32 + lane;
            if (p < nList) {
              const int id = __ldcg(ids + p);
              const int gq = G == 1 ? 0 : (id >> 5) / nr, r = (id >> 5) - gq * nr, gc = gq * 32 + (id & 31);
              const unsigned long long cid = cellIds[gc];
              float dir[3], rayDx[3];
              gen_divq_ray_f32(a.rng, g, ray0 + r, cid, cid, dir, rayDx);
              BlkRay ry;
              float ddx, ddy, ddz;
              bool done = true;
              ry.sum = 0.0f; ry.q.nsteps = 0.0f;
              if (round == 0) {
                if (1.0f > thr) {
                  const int rw0 = w0s[gc];
                  blk_ray_start(ry, blk, rw0, (unsigned)(rw0 & 7), sy8, sz8, nz<float, float>(dir[0]), nz<float, float>(dir[1]),
                                nz<float, float>(dir[2]), rayDx[0], rayDx[1], rayDx[2], cx, cy, cz, ddx, ddy, ddz);
                  done = blk_ray_legs(ry, blk, ddx, ddy, ddz, sy8, sz8, thr, a.allowReflect != 0, skipZero, 1);
                }
              } else {
                ddx = fabsf(N::rcp(nz<float, float>(dir[0]))) * cx * N::LEN_SCALE;
                ddy = fabsf(N::rcp(nz<float, float>(dir[1]))) * cy * N::LEN_SCALE;
                ddz = fabsf(N::rcp(nz<float, float>(dir[2]))) * cz * N::LEN_SCALE;
                const uint4 v0 = __ldcg(parked + id * 3 + 0), v1 = __ldcg(parked + id * 3 + 1), v2 = __ldcg(parked + id * 3 + 2);
                ry.d.tx = __uint_as_float(v0.x); ry.d.ty = __uint_as_float(v0.y); ry.d.tz = __uint_as_float(v0.z); ry.d.idx = (int)v0.w;
                ry.b.dlx = (int)v1.x; ry.b.dly = (int)v1.y; ry.b.dlz = (int)v1.z;
                blk_decode(ry.q, __ldg(blk + ry.d.idx));
                ry.q.A = __uint_as_float(v1.w); ry.q.ePrev = __uint_as_float(v2.x); ry.q.nsteps = __uint_as_float(v2.y);
                ry.q.tau = 0.0f; ry.q.leg = 0.0f;
                ry.sum = __uint_as_float(v2.z); ry.fs = __uint_as_float(v2.w);
                // (after RESORT_ROUNDS rounds whatever is left runs to its end in the company it has)
                done = blk_ray_legs(ry, blk, ddx, ddy, ddz, sy8, sz8, thr, a.allowReflect != 0, skipZero, round < RESORT_ROUNDS ? 1 : 0x7fffffff);
              }
              if (done) {
                steps += (unsigned)ry.q.nsteps;
                __stcg(res + id, UNSCALE * ry.sum);
                __stcg(keys + id, (unsigned short)0xFFFFu);
              } else {
                park(ry, id, dir);
              }
            }
          }
          __syncthreads();
          // the rays parked in this round -> the next round's sorted list
          scan_bins();
          if (tid == 0) ctl[0] = 0;
          if (ctl[1] == 0) break;
          for (int v = warp; v < V; v += W) {
            if (w0s[(G == 1 ? 0 : v / nr) * 32 + lane] < 0) continue;
            const unsigned short k2 = __ldcg(keys + v * 32 + lane);
            if (k2 != 0xFFFFu) {
              const int pos = atomicAdd(&hist[k2], 1);
              __stcg(ids + pos, (unsigned short)(v * 32 + lane));
            }
          }
          __syncthreads();
        }
        __syncthreads();
      } else {
      // ---- phase 3: march chunks of 32 rays, longest first
      {
#ifdef RMCRT_PHASE_CLOCKS
        tMark = clock64();
#endif
        const int nList = ctl[1];
        for (;;) {
          int chunk = 0;
          if (lane == 0) chunk = atomicAdd(&ctl[0], 1);
          chunk = __shfl_sync(0xffffffffu, chunk, 0);
          if (chunk * 32 >= nList) break;
          const int p = chunk * 32 + lane;
          if (p < nList) {
            const int id = __ldcg(ids + p);
            const int gq = G == 1 ? 0 : (id >> 5) / nr, r = (id >> 5) - gq * nr, gc = gq * 32 + (id & 31);
            const int rw0 = w0s[gc];
            const unsigned long long cid = cellIds[gc];
            float dir[3], rayDx[3];
            gen_divq_ray_f32(a.rng, g, ray0 + r, cid, cid, dir, rayDx);
            float sumI;
            if constexpr (ONION)
              sumI = march_onion_blk<MAXL>(a.G, ob, ijk[gc * 3 + 0] + g.lo[0], ijk[gc * 3 + 1] + g.lo[1], ijk[gc * 3 + 2] + g.lo[2], rw0,
                                            nz<float, float>(dir[0]), nz<float, float>(dir[1]), nz<float, float>(dir[2]), rayDx[0], rayDx[1],
                                            rayDx[2], (float)a.threshold, a.allowReflect != 0, skipZero, steps);
            else
              sumI = march_single_level_blk(blk, rw0, (unsigned)(rw0 & 7), sy8, sz8, nz<float, float>(dir[0]), nz<float, float>(dir[1]),
                                            nz<float, float>(dir[2]), rayDx[0], rayDx[1], rayDx[2], cx, cy, cz, (float)a.threshold,
                                            a.allowReflect != 0, skipZero, steps);
            __stcg(res + id, UNSCALE * sumI);
          }
        }
#ifdef RMCRT_PHASE_CLOCKS
        tMarch += clock64() - tMark;      // this warp's time inside the chunk loop (the barrier wait after it counts as "outside")
#endif
      }
      __syncthreads();
      }
      // ---- phase 4: rays of a cell added in ray order (warp w: its contiguous share, then the shares in warp order)
      if (G == 1) {
        const int per = (nr + W - 1) / W;
        double s = 0;
        if (w0s[lane] >= 0)
          for (int r = warp * per; r < min(nr, (warp + 1) * per); ++r) s += (double)__ldcg(res + r * 32 + lane);
        part[warp * 32 + lane] = s;
        __syncthreads();
        if (warp == 0)
          for (int w = 0; w < W; ++w) acc += part[w * 32 + lane];
        __syncthreads();       // hist aliases part
      } else {                 // few rays per cell: warp g adds the rays of its own tile
        if (warp < G && w0s[warp * 32 + lane] >= 0) {
          double s = 0;
          for (int r = 0; r < nr; ++r) s += (double)__ldcg(res + (warp * nr + r) * 32 + lane);
          acc += s;
        }
        __syncthreads();
      }
    }
    if (warp < G && valid) {
      const long long q = (long long)(i - a.outLo[0]) +
                          (long long)a.outN[0] * ((long long)(j - a.outLo[1]) + (long long)a.outN[1] * (long long)(k - a.outLo[2]));
      if (flow) {
        const double mean = acc / a.nRays;
        const uint32_t ww = __ldg(blk + w0);      // the fp16 sigmaT4/pi the rays integrate (see divq_blk_kernel); Ray.cc:775,778
        const double s0 = (double)(UNSCALE * __half2float(__ushort_as_half((unsigned short)(ww & 0xFFFFu))));
        __stcs(a.divQ + q, -4.0 * M_PI * (double)r0.k * (s0 - mean));      // streaming stores: the outputs are not read again
        __stcs(a.volq + q, 4.0 * M_PI * mean);
        ++cellsTotal;
      } else if (a.zeroNonFlow) {
        __stcs(a.divQ + q, 0.0);
        __stcs(a.volq + q, 0.0);
      }
    }
    stepsTotal += steps;
    __syncthreads();      // ctl[2] / part / w0s are rewritten for the next tile
  }
#ifdef RMCRT_PHASE_CLOCKS
  if (!RESORT && lane == 0 && (blockIdx.x % 97) == 0)
    printf("block %d warp %d: %.1f %% of %lld cycles inside the chunk loop\n", blockIdx.x, warp, 100.0 * (double)tMarch / (double)(clock64() - tAll),
           clock64() - tAll);
#endif
  for (int o = 16; o > 0; o >>= 1) {
    stepsTotal += __shfl_down_sync(0xffffffffu, stepsTotal, o);
    cellsTotal += __shfl_down_sync(0xffffffffu, cellsTotal, o);
  }
  if (lane == 0) {
    if (stepsTotal) atomicAdd(a.counters + 0, stepsTotal);
    if (cellsTotal) atomicAdd(a.counters + 2, cellsTotal);
  }
}

//======================================================================================================
// boundary flux
//======================================================================================================
// Ray::has_a_boundary (Ray.cc:1517-1572): `celltype[adj]+1 != 0`  <=>  adj is not a flow cell  <=>  record sign bit.
template <class RecT>
__global__ void flux_mask_kernel(const __grid_constant__ TraceArgs a, unsigned char* __restrict__ flags) {
  const int F = a.G.maxLevels - 1;
  const LevelGeom& g = a.G.lev[F];
  const RecT* recs = static_cast<const RecT*>(g.rec);
  unsigned long long nb = 0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < a.nCells; t += (long long)gridDim.x * blockDim.x) {
    int i, j, k;
    patch_cell(a, t, i, j, k);
    const long long c = level_index(g, i, j, k);
    const bool flow = !is_wall(ldrec(recs + c).k);
    unsigned m = 0;
    if (flow || !a.fluxFlowGuard) {
      const long long sy = g.n[0], sz = (long long)g.n[0] * g.n[1];
      m |= is_wall(ldrec(recs + c - 1).k) ? 1u : 0u;
      m |= is_wall(ldrec(recs + c + 1).k) ? 2u : 0u;
      m |= is_wall(ldrec(recs + c - sy).k) ? 4u : 0u;
      m |= is_wall(ldrec(recs + c + sy).k) ? 8u : 0u;
      m |= is_wall(ldrec(recs + c - sz).k) ? 16u : 0u;
      m |= is_wall(ldrec(recs + c + sz).k) ? 32u : 0u;
      const long long q = (long long)(i - a.outLo[0]) +
                          (long long)a.outN[0] * ((long long)(j - a.outLo[1]) + (long long)a.outN[1] * (long long)(k - a.outLo[2]));
      a.bflux[7 * q + 6] = m ? 1.0 : 0.0;    // boundFlux[origin].p = has_a_boundary(...)  (Ray.cc:664)
      nb += m ? 1 : 0;
    }
#pragma unroll
    for (int f = 0; f < 6; ++f) flags[(long long)f * a.nCells + t] = (m >> f) & 1u;
  }
  for (int o = 16; o > 0; o >>= 1) nb += __shfl_down_sync(0xffffffffu, nb, o);
  if ((threadIdx.x & 31) == 0 && nb) atomicAdd(a.counters + 3, nb);
}

// Ordered stream compaction of the (face, cell) flags into item ids (replaces cub::DeviceSelect::Flagged: no library kernel on the path).
// Three small launches: per-chunk counts, one-block exclusive scan of the counts, scatter with an in-chunk ballot scan.  The order of the
// items is that of the flags (face-major, cells x-fastest), so consecutive items are neighbouring cells of one face -- deterministic.
constexpr int COMPACT_CHUNK = 4096;      // flags per block (256 threads x 16)
__global__ void __launch_bounds__(256) compact_count_kernel(const unsigned char* __restrict__ flags, int n, int* __restrict__ chunkCount) {
  __shared__ int wsum[8];
  const int base = blockIdx.x * COMPACT_CHUNK;
  int c = 0;
  for (int t = threadIdx.x; t < COMPACT_CHUNK; t += 256) { const int q = base + t; c += (q < n && flags[q]) ? 1 : 0; }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) { int s = 0; for (int w = 0; w < 8; ++w) s += wsum[w]; chunkCount[blockIdx.x] = s; }
}
__global__ void __launch_bounds__(1024) compact_scan_kernel(int* __restrict__ chunkCount, int nChunks, int* __restrict__ nItems) {
  __shared__ int part[1024];
  // thread t owns a contiguous run of chunks; exclusive scan in place
  const int per = (nChunks + 1023) / 1024;
  const int lo = min(threadIdx.x * per, nChunks), hi = min(lo + per, nChunks);
  int s = 0;
  for (int q = lo; q < hi; ++q) s += chunkCount[q];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) { int run = 0; for (int t = 0; t < 1024; ++t) { const int v = part[t]; part[t] = run; run += v; } *nItems = run; }
  __syncthreads();
  int run = part[threadIdx.x];
  for (int q = lo; q < hi; ++q) { const int v = chunkCount[q]; chunkCount[q] = run; run += v; }
}
__global__ void __launch_bounds__(256) compact_scatter_kernel(const unsigned char* __restrict__ flags, int n, const int* __restrict__ chunkOffset,
                                                              int* __restrict__ items) {
  __shared__ int wbase[8];
  __shared__ int carry;
  const int base = blockIdx.x * COMPACT_CHUNK;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = chunkOffset[blockIdx.x];
  __syncthreads();
  for (int r = 0; r < COMPACT_CHUNK / 256; ++r) {      // rounds of 256 consecutive flags keep the order
    const int q = base + r * 256 + threadIdx.x;
    const bool f = q < n && flags[q];
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (lane == 0) wbase[warp] = __popc(m);
    __syncthreads();
    int off = carry;
    for (int w = 0; w < warp; ++w) off += wbase[w];
    if (f) items[off + __popc(m & ((1u << lane) - 1u))] = q;
    __syncthreads();
    if (threadIdx.x == 0) { int s = 0; for (int w = 0; w < 8; ++w) s += wbase[w]; carry += s; }
    __syncthreads();
  }
}

// lane = one (face, cell) item; consecutive items are neighbouring cells on the same face.
template <class Real, class RecT, bool ML, bool FAST = true>
__global__ void __launch_bounds__(256) flux_trace_kernel(const __grid_constant__ TraceArgs a) {
  extern __shared__ double red[];     // [W][32][2]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int F = a.G.maxLevels - 1;
  const LevelGeom& g = a.G.lev[F];
  const long long nItems = *a.nItems;
  const bool skipZero = skip_allowed(a);
  const long long nTiles = (nItems + 31) / 32;
  unsigned long long stepsTotal = 0;
  for (long long tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    const long long it = tile * 32 + lane;
    const bool valid = it < nItems;
    int i = 0, j = 0, k = 0, face = 0;
    unsigned long long cellId = 0, dirId = 0;
    if (valid) {
      const long long id = a.items[it];
      face = (int)(id / a.nCells);
      patch_cell(a, id - (long long)face * a.nCells, i, j, k);
      cellId = (unsigned long long)level_index(g, i, j, k);
    }
    {
      const long long id0 = a.items[tile * 32];
      const int f0 = (int)(id0 / a.nCells);
      int i0, j0, k0;
      patch_cell(a, id0 - (long long)f0 * a.nCells, i0, j0, k0);
      // a warp straddling two faces still gets per-face directions through `face` in the counter
      dirId = a.rng.coherent ? (unsigned long long)level_index(g, i0, j0, k0) : cellId;
    }
    double sumProj = 0, sumCos = 0;
    unsigned steps = 0;
    if (valid) {
      for (int ray = warp; ray < a.nRays; ray += W) {
        double dir[3], rayDx[3], cosT, dI;
        bool done = false;
        if constexpr (!FAST && !ML) {
          if (a.sigmaScat > 0) {
            double org[3];
            gen_flux_ray(a.rng, g, i, j, k, face, ray, a.nRays, cellId, dirId, dir, rayDx, cosT, org);
            ScatRng srng = scat_rng(a.rng, i, j, k, ray, cellId, 4, (uint32_t)(face + 1) << 28);
            dI = (double)trace_one<Real, RecT, ML, false, FAST>(a, skipZero, i, j, k, dir, rayDx, steps, nullptr, &srng, org);
            done = true;
          }
        }
        if (!done) {
          gen_flux_ray(a.rng, g, i, j, k, face, ray, a.nRays, cellId, dirId, dir, rayDx, cosT);
          bool blocked = false;
          if constexpr (FAST && !ML && std::is_same<Real, float>::value) {
            if (a.fluxBlk) {      // the production solve: the same blocked half-packed copy the divQ rays march (indep.cuh)
              const int ux = i - a.fluxLo[0], uy = j - a.fluxLo[1], uz = k - a.fluxLo[2];
              const unsigned par = (unsigned)((ux & 1) | ((uy & 1) << 1) | ((uz & 1) << 2));
              const int w0 = 8 * ((ux >> 1) + a.fluxBx * (uy >> 1) + a.fluxBxy * (uz >> 1)) + (int)par;
              dI = (double)((1.0f / HALF_S_SCALE) *
                            march_single_level_blk(a.fluxBlk, w0, par, 8 * a.fluxBx, 8 * a.fluxBxy, nz<float, double>(dir[0]), nz<float, double>(dir[1]),
                                                   nz<float, double>(dir[2]), (float)rayDx[0], (float)rayDx[1], (float)rayDx[2], (float)g.dx[0],
                                                   (float)g.dx[1], (float)g.dx[2], (float)a.threshold, a.allowReflect != 0, skipZero, steps));
              blocked = true;
            }
          }
          if (!blocked) dI = (double)trace_one<Real, RecT, ML, false, FAST>(a, skipZero, i, j, k, dir, rayDx, steps, nullptr);
        }
        sumProj += cosT * dI;      // sumProjI += cosTheta * (sumI - sumI_prev)   (Ray.cc:704)
        sumCos += cosT;
      }
    }
    red[(warp * 32 + lane) * 2 + 0] = sumProj;
    red[(warp * 32 + lane) * 2 + 1] = sumCos;
    __syncthreads();
    if (warp == 0 && valid) {
      double sp = 0, sc = 0;
      for (int w = 0; w < W; ++w) { sp += red[(w * 32 + lane) * 2]; sc += red[(w * 32 + lane) * 2 + 1]; }
      sp = sp * (double)a.nRays / sc / 2.0;                              // Ray.cc:712
      const long long q = (long long)(i - a.outLo[0]) +
                          (long long)a.outN[0] * ((long long)(j - a.outLo[1]) + (long long)a.outN[1] * (long long)(k - a.outLo[2]));
      a.bflux[7 * q + face] = sp * 2 * M_PI / (double)a.nRays;           // Ray.cc:716-717
    }
    __syncthreads();
    stepsTotal += steps;
  }
  for (int o = 16; o > 0; o >>= 1) stepsTotal += __shfl_down_sync(0xffffffffu, stepsTotal, o);
  if (lane == 0 && stepsTotal) atomicAdd(a.counters + 1, stepsTotal);
}

//======================================================================================================
// virtual radiometer (Radiometer.cc:428-560).  A radiometer is a handful of cells with thousands of rays each, so the
// parallel axis is the ray: one block per radiometer cell, thread t marches rays t, t+B, ...; per-thread FP64 partial sums
// are combined in a fixed order (deterministic).
//======================================================================================================
struct RadArgs {
  rmcrt_radiometer rad;
  int lo[3], n[3];        // cells of (radiometer box n patch)
  void* flux;
  void* intensity;
  int isFloat;
};

template <class Real, class RecT>
__global__ void __launch_bounds__(256) radiometer_kernel(const __grid_constant__ TraceArgs a, const __grid_constant__ RadArgs ra) {
  __shared__ double red[256];
  __shared__ unsigned long long redSteps[256];
  const int F = a.G.maxLevels - 1;
  const LevelGeom& g = a.G.lev[F];
  const bool skipZero = skip_allowed(a);
  const long long nCells = (long long)ra.n[0] * ra.n[1] * ra.n[2];
  const rmcrt_radiometer& rad = ra.rad;
  // rotation matrix of Radiometer.cc:539-552 (phi_rotate is always 0, kept for fidelity)
  const double cth = cos(rad.theta_rotate), sth = sin(rad.theta_rotate), cxi = cos(rad.xi_rotate), sxi = sin(rad.xi_rotate),
               cph = cos(rad.phi_rotate), sph = sin(rad.phi_rotate);
  const double cosView = cos(rad.theta_viewAngle);
  for (long long cell = blockIdx.x; cell < nCells; cell += gridDim.x) {
    const int ci = (int)(cell % ra.n[0]), cj = (int)((cell / ra.n[0]) % ra.n[1]), ck = (int)(cell / ((long long)ra.n[0] * ra.n[1]));
    const int i = ra.lo[0] + ci, j = ra.lo[1] + cj, k = ra.lo[2] + ck;
    const unsigned long long cellId = (unsigned long long)level_index(g, i, j, k);
    double sumProj = 0;
    unsigned steps = 0;
    for (int ray = threadIdx.x; ray < rad.nRays; ray += blockDim.x) {
      double R1, R2;
      if (a.rng.fixedRandom) R1 = R2 = 0.3;
      else if (!a.rng.isSeedRandom) {
        uint32_t v[2];
        mt19937_first<2>((uint32_t)((i + j + k) * ray + 1), v);      // Radiometer.cc:512-514
        R1 = u01_open(v[0]); R2 = u01_open(v[1]);
      } else {
        const u32x4 w = philox4x32_10(u32x4{(uint32_t)cellId, (uint32_t)ray, a.rng.timestep,
                                            ((uint32_t)(cellId >> 32) << 8) | STREAM_RADIOMETER}, a.rng.key0, a.rng.key1);
        R1 = u01_open(w.x); R2 = u01_open(w.y);
      }
      const double phiRay = 2 * M_PI * R1;                        // 526
      const double thetaRay = acos(cosView + rad.range * R2);     // 530
      const double cosT = cos(thetaRay);
      const double x = sin(thetaRay) * cos(phiRay), y = sin(thetaRay) * sin(phiRay), z = cosT;   // 534-536
      double dir[3];
      dir[0] = x * cth * cxi + y * (-cph * sxi + sph * sth * cxi) + z * (sph * sxi + cph * sth * cxi);
      dir[1] = x * cth * sxi + y * (cph * cxi + sph * sth * sxi) + z * (-sph * cxi + cph * sth * sxi);
      dir[2] = x * -sth + y * sph * cth + z * cph * cth;
      double rayDx[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {                               // ray_Origin with useCCRays = true (Radiometer.cc:472-473)
        const int c = d == 0 ? i : (d == 1 ? j : k);
        const double cc = g.anchor[d] + g.dx[d] * c + g.dx[d] * 0.5;
        rayDx[d] = cc - (cc - 0.5 * g.dx[d]);
      }
      const double dI = (double)trace_one<Real, RecT, false, false, false>(a, skipZero, i, j, k, dir, rayDx, steps, nullptr);
      sumProj += cosT * dI;                                       // 483
    }
    red[threadIdx.x] = sumProj;
    redSteps[threadIdx.x] = steps;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {               // fixed-order tree: deterministic
      if ((int)threadIdx.x < o) { red[threadIdx.x] += red[threadIdx.x + o]; redSteps[threadIdx.x] += redSteps[threadIdx.x + o]; }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const long long q = (long long)(i - a.outLo[0]) +
                          (long long)a.outN[0] * ((long long)(j - a.outLo[1]) + (long long)a.outN[1] * (long long)(k - a.outLo[2]));
      const double sp = red[0];
      if (ra.isFloat) {     // `(T) sumProjI * solidAngle/nRays`: the cast binds to sumProjI (Radiometer.cc:492-493)
        static_cast<float*>(ra.flux)[q] = (float)((double)(float)sp * rad.solidAngle / rad.nRays);
        static_cast<float*>(ra.intensity)[q] = (float)((float)sp / rad.nRays);
      } else {
        static_cast<double*>(ra.flux)[q] = sp * rad.solidAngle / rad.nRays;
        static_cast<double*>(ra.intensity)[q] = sp / rad.nRays;
      }
      if (redSteps[0]) atomicAdd(a.counters + 4, redSteps[0]);
    }
    __syncthreads();
  }
}

//======================================================================================================
// debug rays
//======================================================================================================
template <class Real, class RecT, bool ML>
__global__ void debug_rays_kernel(const __grid_constant__ TraceArgs a, int n, const int* __restrict__ cells,
                                  const double* __restrict__ dirs, const double* __restrict__ origins,
                                  double* __restrict__ sumI, unsigned* __restrict__ stepsOut, int* pathBuf, int maxPath) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int F = a.G.maxLevels - 1;
  const LevelGeom& g = a.G.lev[F];
  const int i = cells[3 * t], j = cells[3 * t + 1], k = cells[3 * t + 2];
  const int c[3] = {i, j, k};
  double dir[3], rayDx[3];
  for (int d = 0; d < 3; ++d) {
    dir[d] = dirs[3 * t + d];
    const double cc = g.anchor[d] + g.dx[d] * c[d] + g.dx[d] * 0.5;   // RMCRTCommon.cc:640-646
    rayDx[d] = origins[3 * t + d] - (cc - 0.5 * g.dx[d]);
  }
  PathSink sink{pathBuf ? pathBuf + (long long)t * maxPath * 4 : nullptr, maxPath, 0};
  unsigned steps = 0;
  const Real v = trace_one<Real, RecT, ML, true>(a, skip_allowed(a), i, j, k, dir, rayDx, steps, pathBuf ? &sink : nullptr);
  sumI[t] = (double)v;
  stepsOut[t] = steps;
}

//======================================================================================================
// property pipeline
//======================================================================================================
template <class RecT> struct RecTraits;
template <> struct RecTraits<Rec8>  { typedef float  real; };
template <> struct RecTraits<Rec16> { typedef double real; };
template <> struct RecTraits<Rec4>  { typedef float  real; };
template <class RecT, class R> __device__ __forceinline__ RecT make_rec(R kSigned, R s) { return RecT{kSigned, s}; }
template <> __device__ __forceinline__ Rec4 make_rec<Rec4, float>(float kSigned, float s) { return encode_rec4(kSigned, s); }
// true (unscaled) values of a record
template <class RecT> __device__ __forceinline__ typename RecInfo<RecT>::decoded load_true(const RecT* p) {
  typename RecInfo<RecT>::decoded r = ldrec(p);
  r.s = r.s * (typename RecInfo<RecT>::real)RecInfo<RecT>::unscale;
  return r;
}

struct Box { int lo[3], n[3]; };
__host__ __device__ inline long long box_cells(const Box& b) { return (long long)b.n[0] * b.n[1] * b.n[2]; }

template <class T, class RecT>
__global__ void pack_kernel(const T* __restrict__ abskg, const T* __restrict__ sig, const int* __restrict__ ct, Box src,
                            Box region /* cells to pack */, RecT* __restrict__ dst, Box dstBox, int flowCell,
                            unsigned* __restrict__ status, Rec8* __restrict__ full, Box fullBox) {
  typedef typename RecTraits<RecT>::real R;
  const long long n = box_cells(region);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / region.n[0];
    const int i = region.lo[0] + (int)(t - row * region.n[0]);
    const int kk = (int)(row / region.n[1]);
    const int j = region.lo[1] + (int)(row - (long long)kk * region.n[1]);
    const int k = region.lo[2] + kk;
    const long long s = (long long)(i - src.lo[0]) + (long long)src.n[0] * ((long long)(j - src.lo[1]) + (long long)src.n[1] * (long long)(k - src.lo[2]));
    const long long d = (long long)(i - dstBox.lo[0]) + (long long)dstBox.n[0] * ((long long)(j - dstBox.lo[1]) + (long long)dstBox.n[1] * (long long)(k - dstBox.lo[2]));
    const T kv = abskg[s], sv = sig[s];
    if (!isfinite((double)kv) || !isfinite((double)sv)) atomicAdd(status + 5, 1u);   // RMCRTCommon.cc:267-271
    R kr = (R)kv;
    kr = kr < R(0) ? -kr : kr;
    if (ct[s] != flowCell) kr = -kr;     // sign bit == "not a flow cell", also for abskg == 0 (-0.0)
    const RecT rec = make_rec<RecT, R>(kr, (R)sv);
    dst[d] = rec;
    if (std::is_same<RecT, Rec4>::value) {
      const auto back = ldrec(&dst[d]);   // fp16 overflow (T > 3270 K or abskg > 65504) must not pass silently
      if (isinf((float)back.k) || isinf((float)back.s)) atomicAdd(status + 5, 1u);
      if (full) {
        const long long f = (long long)(i - fullBox.lo[0]) + (long long)fullBox.n[0] * ((long long)(j - fullBox.lo[1]) + (long long)fullBox.n[1] * (long long)(k - fullBox.lo[2]));
        full[f] = Rec8{(float)kr, (float)sv};
      }
    }
  }
}

template <class RecT>
__global__ void finalize_kernel(const RecT* __restrict__ recs, Box b, unsigned* __restrict__ status) {
  const long long n = box_cells(b);
  unsigned open = 0, emit = 0, bad = 0, refl = 0, absorbing = 0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / b.n[0];
    const int i = (int)(t - row * b.n[0]);
    const int k = (int)(row / b.n[1]);
    const int j = (int)(row - (long long)k * b.n[1]);
    const auto r = ldrec(recs + t);
    const bool wall = is_wall(r.k);
    const bool outer = (i == 0 || j == 0 || k == 0 || i == b.n[0] - 1 || j == b.n[1] - 1 || k == b.n[2] - 1);
    if (outer && !wall) ++open;
    if (wall && (r.k != 0) && (r.s != 0)) ++emit;
    if (wall && fabs((double)r.k) < 1.0) ++refl;        // fs = fs (1 - abskg) stays > 0 after this wall (RMCRTCommon.cc:588)
    if (!isfinite((double)r.k) || !isfinite((double)r.s)) ++bad;
    // faces of the level whose outer layer absorbs (abskg >= 0.5; edges and corners aside): the others send most of a ray's weight on
    const int nOuter = (i == 0) + (j == 0) + (k == 0) + (i == b.n[0] - 1) + (j == b.n[1] - 1) + (k == b.n[2] - 1);
    if (nOuter == 1 && !(wall && fabs((double)r.k) < 0.5))
      absorbing |= i == 0 ? 1u : (i == b.n[0] - 1 ? 2u : (j == 0 ? 4u : (j == b.n[1] - 1 ? 8u : (k == 0 ? 16u : 32u))));
  }
  absorbing = __reduce_or_sync(0xffffffffu, absorbing);
  if (absorbing && (threadIdx.x & 31) == 0) atomicOr(status + 7, absorbing);
  if (bad) atomicAdd(status + 0, bad);
  if (open) atomicAdd(status + 1, open);
  if (emit) atomicAdd(status + 2, emit);
  if (refl) atomicAdd(status + 6, refl);
  // the last block to finish publishes the result (words 0..2 were zeroed before the launch)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(status + 4, 1u);
    if (t == gridDim.x - 1) { __threadfence(); atomicExch(status + 3, 1u); }
  }
}

// Ray::ROI_Extents (Ray.cc:1984-2043) for one patch: magnitude of the first-order central-difference gradient of abskg and sigmaT4/pi
// (compute_Mag_gradient, Core/Grid/AMR.cc:464-483: T inv_dx = (T)0.5 / dx; (q[r] - q[l]) * inv_dx in T; Vector::length() in double; stored
// as T) and the flag `mag_grad_abskg > abskg_threshold || mag_grad_sigmaT4 > sigmaT4_threshold`.  Optional outputs over the out window.
template <class RecT, class T>
__global__ void roi_extents_kernel(const RecT* __restrict__ recs, Box lb, Box patch, double dx0, double dx1, double dx2, double thrA, double thrS,
                                   T* __restrict__ magA, T* __restrict__ magS, int* __restrict__ flagged, Box out, unsigned* __restrict__ any) {
  const long long n = box_cells(patch);
  const T inv[3] = {(T)((T)0.5 / dx0), (T)((T)0.5 / dx1), (T)((T)0.5 / dx2)};
  const long long st[3] = {1, lb.n[0], (long long)lb.n[0] * lb.n[1]};
  unsigned mine = 0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / patch.n[0];
    const int i = patch.lo[0] + (int)(t - row * patch.n[0]), j = patch.lo[1] + (int)(row % patch.n[1]), k = patch.lo[2] + (int)(row / patch.n[1]);
    const long long c = (long long)(i - lb.lo[0]) + st[1] * (long long)(j - lb.lo[1]) + st[2] * (long long)(k - lb.lo[2]);
    double ga[3], gs[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const auto r = load_true(recs + c + st[d]), l = load_true(recs + c - st[d]);
      const T ar = (T)fabs((double)r.k), al = (T)fabs((double)l.k);       // abskg = |k| (the sign bit is the wall flag)
      ga[d] = (double)(T)((T)(ar - al) * inv[d]);
      gs[d] = (double)(T)((T)((T)r.s - (T)l.s) * inv[d]);
    }
    const T ma = (T)sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ga[0], ga[0]), __dmul_rn(ga[1], ga[1])), __dmul_rn(ga[2], ga[2])));
    const T ms = (T)sqrt(__dadd_rn(__dadd_rn(__dmul_rn(gs[0], gs[0]), __dmul_rn(gs[1], gs[1])), __dmul_rn(gs[2], gs[2])));
    const bool f = (double)ma > thrA || (double)ms > thrS;
    mine |= f ? 1u : 0u;
    if (magA) {
      const long long q = (long long)(i - out.lo[0]) + (long long)out.n[0] * ((long long)(j - out.lo[1]) + (long long)out.n[1] * (long long)(k - out.lo[2]));
      magA[q] = ma; magS[q] = ms; flagged[q] = f ? 1 : 0;
    }
  }
  if (__any_sync(0xffffffffu, mine != 0u) && (threadIdx.x & 31) == 0) atomicOr(any, 1u);
}

// blocked half-packed copy of a level (indep.cuh): thread = one 2 x 2 x 2 block of the guard-padded block grid
template <class RecT>
__global__ void blk_build_kernel(const RecT* __restrict__ recs, Box lb, uint32_t* __restrict__ blk, int bx, int by, int bz) {
  const long long n = (long long)bx * by * bz;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int b0 = (int)(t % bx);
    const long long rest = t / bx;
    const int b1 = (int)(rest % by), b2 = (int)(rest / by);
    uint32_t w[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      // block-local coordinate 0 = cell lb.lo - BLK_PAD
      const int x = 2 * b0 + (c & 1) - BLK_PAD, y = 2 * b1 + ((c >> 1) & 1) - BLK_PAD, z = 2 * b2 + (c >> 2) - BLK_PAD;
      uint32_t v = 0x80000000u;                     // guard: a wall with abskg 0, sigmaT4 0
      if (x >= 0 && x < lb.n[0] && y >= 0 && y < lb.n[1] && z >= 0 && z < lb.n[2]) {
        const long long f = (long long)x + (long long)lb.n[0] * ((long long)y + (long long)lb.n[1] * (long long)z);
        if constexpr (std::is_same<RecT, Rec4>::value) v = recs[f].w;
        else { const Rec8 r = ldrec(recs + f); v = encode_rec4(r.k, r.s).w; }
      }
      w[c] = v;
    }
    uint4* o = reinterpret_cast<uint4*>(blk + t * 8);
    o[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

// blocked half-packed copy of the fine level's region of interest for the Data-Onion march (march_onion2_blk): thread = one 2 x 2 x 2 block
// of the box [lo, lo + 2 nb); cells inside roi = the fine level's records, every other cell a marker {wall flag, abskg = inf} -- the
// one-cell shell that ends the ROI leg, and the guard blocks the speculative load past it may touch
template <class RecT>
__global__ void roi_blk_build_kernel(const RecT* __restrict__ fine, Box fb, uint32_t* __restrict__ blk, int bx, int by, int bz, int lo0, int lo1,
                                     int lo2, Box roi) {
  const long long n = (long long)bx * by * bz;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int b0 = (int)(t % bx);
    const long long rest = t / bx;
    const int b1 = (int)(rest % by), b2 = (int)(rest / by);
    uint32_t w[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int x = lo0 + 2 * b0 + (c & 1), y = lo1 + 2 * b1 + ((c >> 1) & 1), z = lo2 + 2 * b2 + (c >> 2);
      uint32_t v = 0xFC000000u;                     // fp16 -inf | 0
      if (x >= roi.lo[0] && x < roi.lo[0] + roi.n[0] && y >= roi.lo[1] && y < roi.lo[1] + roi.n[1] && z >= roi.lo[2] && z < roi.lo[2] + roi.n[2]) {
        const long long f = (long long)(x - fb.lo[0]) + (long long)fb.n[0] * ((long long)(y - fb.lo[1]) + (long long)fb.n[1] * (long long)(z - fb.lo[2]));
        if constexpr (std::is_same<RecT, Rec4>::value) v = fine[f].w;
        else { const Rec8 r = ldrec(fine + f); v = encode_rec4(r.k, r.s).w; }
      }
      w[c] = v;
    }
    uint4* o = reinterpret_cast<uint4*>(blk + t * 8);
    o[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

template <class T>
__global__ void sigmaT4_kernel(const double* __restrict__ temp, T* __restrict__ out, long long n, double sigma_over_pi) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const double T2 = temp[t] * temp[t];
    out[t] = (T)(sigma_over_pi * T2 * T2);   // RMCRTCommon.cc:339-340
  }
}

// Ray::coarsen_Q (Ray.cc:2103-2145; AMR_CoarsenRefine.cc:42-86) + Ray::computeCellType (2272-2345), interior coarse cells.
// computeCellType sums the INTEGER cell types of the RR fine cells (flow = -1) and decides on tmp = sum / RR:
//   tmp == -1 -> flow;  ROUNDUP: intrusion iff |tmp|/10 >= 1/RR (i.e. |sum| >= 10);  ROUNDDOWN: flow iff |tmp|/10 <= 1 - 1/RR.
// The packed record keeps one bit of the cell type, so the sum is rebuilt as -nFlow + wallType * nNonFlow with
// wallType = rmcrt_params.nonFlowCellType (8: the wall / intrusion id of RMCRT_Test.cc:80 and of Arches' walls) and the reference's two
// floating-point tests are then applied verbatim -- e.g. RR = 8 with ONE non-flow fine cell gives sum = 1: the coarse cell stays a flow
// cell under ROUNDUP, as in the reference.  Exact whenever the non-flow cells of a level share one type.
template <class RecT>
__global__ void coarsen_kernel(const RecT* __restrict__ fine, Box fb, RecT* __restrict__ coarse, Box cb, Box cint,
                               int rr0, int rr1, int rr2, int logic, int wallType, Rec8* __restrict__ coarseFull) {
  typedef typename RecTraits<RecT>::real R;
  const long long n = box_cells(cint);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / cint.n[0];
    const int i = cint.lo[0] + (int)(t - row * cint.n[0]);
    const int kk = (int)(row / cint.n[1]);
    const int j = cint.lo[1] + (int)(row - (long long)kk * cint.n[1]);
    const int k = cint.lo[2] + kk;
    // Level::mapCellToFiner (Level.cc:1311-1332)
    const int fi = i * rr0 + ((i < 0 && rr0 > 1) ? 1 : 0), fj = j * rr1 + ((j < 0 && rr1 > 1) ? 1 : 0),
              fk = k * rr2 + ((k < 0 && rr2 > 1) ? 1 : 0);
    R ak = 0, as = 0;
    int nWall = 0;
    for (int c = 0; c < rr2; ++c) for (int b = 0; b < rr1; ++b) for (int a = 0; a < rr0; ++a) {
      const long long f = (long long)(fi + a - fb.lo[0]) + (long long)fb.n[0] * ((long long)(fj + b - fb.lo[1]) + (long long)fb.n[1] * (long long)(fk + c - fb.lo[2]));
      const auto r = load_true(fine + f);
      ak += (r.k < R(0) || is_wall(r.k)) ? -r.k : r.k;   // accumulate in T, x-fastest, like coarsenDriver_std
      as += r.s;
      nWall += is_wall(r.k) ? 1 : 0;
    }
    const double invRR = 1.0 / ((double)(rr0 * rr1 * rr2));
    R kc = (R)(ak * invRR), sc = (R)(as * invRR);
    const int RR = rr0 * rr1 * rr2;
    const double tmp = (double)(-(RR - nWall) + wallType * nWall) * invRR;      // Ray.cc:2320
    bool wall = false;
    if (tmp != -1.0) {                                                          // 2324
      if (logic == 1) wall = !(fabs(tmp) / 10.0 <= (1.0 - invRR));              // ROUNDDOWN, 2327-2333
      else            wall = fabs(tmp) / 10.0 >= invRR;                         // ROUNDUP, 2336-2342
    }
    const long long d = (long long)(i - cb.lo[0]) + (long long)cb.n[0] * ((long long)(j - cb.lo[1]) + (long long)cb.n[1] * (long long)(k - cb.lo[2]));
    coarse[d] = make_rec<RecT, R>(wall ? -kc : kc, sc);
    if (coarseFull) coarseFull[d] = Rec8{(float)(wall ? -kc : kc), (float)sc};
  }
}

// <coarsenExtraCells> true (Ray::coarsen_Q, Ray.cc:2148-2245): the coarse level's extra cells on the faces of the computational domain take
// the average of the fine level's extra cells they cover -- refinement ratio 1 along the face normal (one layer of extra cells on both
// levels), the full ratio in the two tangential directions, faces WITHOUT their edges (Patch::ExtraMinusEdgeCells); accumulated in T and
// scaled by the double 1/RR like the reference.  Edge and corner extra cells keep the 0 of Q_coarse.initialize(0.0); no ray can enter
// them.  The cell type of coarse extra cells is not touched by computeCellType (it loops over interior cells): they stay the non-flow
// cells the caller initialised, i.e. the wall flag is set.
template <class RecT>
__global__ void coarsen_extra_kernel(const RecT* __restrict__ fine, Box fb, Box fint, RecT* __restrict__ coarse, Box cb, Box cint,
                                     int rr0, int rr1, int rr2, Rec8* __restrict__ coarseFull) {
  typedef typename RecTraits<RecT>::real R;
  const long long n = box_cells(cb);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / cb.n[0];
    const int c[3] = {cb.lo[0] + (int)(t - row * cb.n[0]), cb.lo[1] + (int)(row % cb.n[1]), cb.lo[2] + (int)(row / cb.n[1])};
    int nOut = 0, axis = -1;
    bool high = false;
    for (int d = 0; d < 3; ++d) {
      if (c[d] < cint.lo[d]) { ++nOut; axis = d; high = false; }
      else if (c[d] >= cint.lo[d] + cint.n[d]) { ++nOut; axis = d; high = true; }
    }
    if (nOut == 0) continue;                      // interior: coarsen_kernel
    R kc = 0, sc = 0;
    if (nOut == 1) {
      const int rr[3] = {rr0, rr1, rr2};
      int f0[3], cnt[3];
      for (int d = 0; d < 3; ++d) {
        if (d == axis) { f0[d] = high ? fint.lo[d] + fint.n[d] : fint.lo[d] - 1; cnt[d] = 1; }     // the fine extra-cell layer of that face
        else { f0[d] = c[d] * rr[d]; cnt[d] = rr[d]; }
      }
      R ak = 0, as = 0;
      for (int z = 0; z < cnt[2]; ++z) for (int y = 0; y < cnt[1]; ++y) for (int x = 0; x < cnt[0]; ++x) {
        const long long f = (long long)(f0[0] + x - fb.lo[0]) + (long long)fb.n[0] * ((long long)(f0[1] + y - fb.lo[1]) + (long long)fb.n[1] * (long long)(f0[2] + z - fb.lo[2]));
        const auto r = load_true(fine + f);
        ak += (r.k < R(0) || is_wall(r.k)) ? -r.k : r.k;
        as += r.s;
      }
      const double invRR = 1.0 / ((double)(cnt[0] * cnt[1] * cnt[2]));
      kc = (R)(ak * invRR); sc = (R)(as * invRR);
    }
    coarse[t] = make_rec<RecT, R>(-kc, sc);       // wall flag set (also for abskg == 0: -0.0)
    if (coarseFull) coarseFull[t] = Rec8{(float)(-kc), (float)sc};
  }
}

template <class RecT>
__global__ void extra_cells_kernel(RecT* __restrict__ recs, Box b, Box interior, const double* __restrict__ vals /*[3][6]*/,
                                   const int* __restrict__ types, int flowCell, Rec8* __restrict__ full) {
  typedef typename RecTraits<RecT>::real R;
  const long long n = box_cells(b);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / b.n[0];
    const int i = b.lo[0] + (int)(t - row * b.n[0]);
    const int kk = (int)(row / b.n[1]);
    const int j = b.lo[1] + (int)(row - (long long)kk * b.n[1]);
    const int k = b.lo[2] + kk;
    // face order x-,x+,y-,y+,z-,z+ ; later faces overwrite earlier ones on shared edges, like the setBC face loop
    int face = -1;
    if (i < interior.lo[0]) face = 0;
    if (i >= interior.lo[0] + interior.n[0]) face = 1;
    if (j < interior.lo[1]) face = 2;
    if (j >= interior.lo[1] + interior.n[1]) face = 3;
    if (k < interior.lo[2]) face = 4;
    if (k >= interior.lo[2] + interior.n[2]) face = 5;
    if (face < 0) continue;
    R kr = (R)fabs(vals[face]);
    if (types[face] != flowCell) kr = -kr;
    recs[t] = make_rec<RecT, R>(kr, (R)vals[6 + face]);
    if (full) full[t] = Rec8{(float)kr, (float)vals[6 + face]};
  }
}

template <class RecT>
__global__ void assemble_kernel(const unsigned char* __restrict__ gathered, size_t slotBytes, int rank, Box box,
                                RecT* __restrict__ dst, Box dstBox) {
  const RecT* src = reinterpret_cast<const RecT*>(gathered + (size_t)rank * slotBytes);
  const long long n = box_cells(box);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / box.n[0];
    const int i = box.lo[0] + (int)(t - row * box.n[0]);
    const int kk = (int)(row / box.n[1]);
    const int j = box.lo[1] + (int)(row - (long long)kk * box.n[1]);
    const int k = box.lo[2] + kk;
    const long long d = (long long)(i - dstBox.lo[0]) + (long long)dstBox.n[0] * ((long long)(j - dstBox.lo[1]) + (long long)dstBox.n[1] * (long long)(k - dstBox.lo[2]));
    dst[d] = src[t];
  }
}


// Data-Onion fast path: copy of the fine level over [roiLo, roiHi) wrapped in a one-cell shell (and the guard records
// around the box) of marker records {wall flag, abskg = inf}; replaces the reference's per-patch getRegion(ROI) (Ray.cc:1060-1075)
template <class RecT>
__global__ void roi_copy_kernel(const RecT* __restrict__ fine, Box fb, RecT* __restrict__ dst /* incl. guards */, long long guard,
                                Box shellBox, Box roiBox) {
  const long long nBox = box_cells(shellBox), n = nBox + 2 * guard;
  const RecT marker = make_rec<RecT, float>(-INFINITY, 0.0f);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long u = t - guard;
    RecT v = marker;
    if (u >= 0 && u < nBox) {
      const long long row = u / shellBox.n[0];
      const int i = shellBox.lo[0] + (int)(u - row * shellBox.n[0]);
      const int kk = (int)(row / shellBox.n[1]);
      const int j = shellBox.lo[1] + (int)(row - (long long)kk * shellBox.n[1]);
      const int k = shellBox.lo[2] + kk;
      const bool in = i >= roiBox.lo[0] && i < roiBox.lo[0] + roiBox.n[0] && j >= roiBox.lo[1] && j < roiBox.lo[1] + roiBox.n[1] &&
                      k >= roiBox.lo[2] && k < roiBox.lo[2] + roiBox.n[2];
      if (in) v = fine[(long long)(i - fb.lo[0]) + (long long)fb.n[0] * ((long long)(j - fb.lo[1]) + (long long)fb.n[1] * (long long)(k - fb.lo[2]))];
    }
    dst[t] = v;
  }
}

// ---- callers either side of the path: streaming kernels, HBM bound ----
// Ray::refine_Q (Ray.cc:1872-1945) = selectInterpolator (Core/Grid/AMR.h:708-745) for one scalar of one fine patch.
// FP64 without FMA contraction and in the reference's operation order, so the result equals the CPU's bit for bit.
struct RefineArgs {
  Box cw;            // window of the coarse array (getRegion box)
  Box fr;            // fine cells to fill [fl,fh)
  Box ow;            // window of the fine array
  int gridLo[3], gridHiM1[3];   // coarseLevel->findCellIndexRange(), high - 1 (AMR.h:351-353)
  int rr[3];
  int order;
};
// normalizedDistance_CC (Core/Grid/AMR.cc:207-235) for one fine index inside its coarse cell
__device__ __forceinline__ double norm_dist_cc(int rel, int rr) {
  if (rr <= 1) return 0.0;
  const int half = rr / 2;
  const double hdx = __dmul_rn(0.5, 1.0 / (double)rr);
  if (rel < half) return __dmul_rn((double)(-(rr - 1 - 2 * rel)), hdx);
  if (rel >= rr) return 0.0;
  const int m = half - 1 - (rel - half);          // mirror copy; odd ratios keep 0 for the last index
  return m >= 0 ? fabs(__dmul_rn((double)(-(rr - 1 - 2 * m)), hdx)) : 0.0;
}
__device__ __forceinline__ int map_coarser_dev(int i, int rr) {     // Level::mapCellToCoarser (Level.cc:1285-1307)
  int q = i / rr;
  if (i < 0 && rr > 1) q += (int)fmod((double)i, (double)rr);
  return q;
}
__global__ void refine_kernel(const double* __restrict__ qC, double* __restrict__ qF, const __grid_constant__ RefineArgs a) {
  const long long n = box_cells(a.fr);
  const long long cnx = a.cw.n[0], cny = a.cw.n[1];
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / a.fr.n[0];
    const int kk = (int)(row / a.fr.n[1]);
    const int f[3] = {a.fr.lo[0] + (int)(t - row * a.fr.n[0]), a.fr.lo[1] + (int)(row - (long long)kk * a.fr.n[1]), a.fr.lo[2] + kk};
    int c[3];
    for (int d = 0; d < 3; ++d) c[d] = map_coarser_dev(f[d], a.rr[d]);
    auto C = [&](int i, int j, int k) -> double {
      return __ldg(qC + ((i - a.cw.lo[0]) + cnx * ((j - a.cw.lo[1]) + cny * (long long)(k - a.cw.lo[2]))));
    };
    double out;
    if (a.order == 0) {
      out = C(c[0], c[1], c[2]);
    } else if (a.order == 1) {
      double dist[3];
      int off[3];
      for (int d = 0; d < 3; ++d) {
        dist[d] = norm_dist_cc(f[d] - c[d] * a.rr[d], a.rr[d]);
        const double ad = fabs(dist[d]);
        off[d] = (dist[d] < 0 ? -1 : 1) * (((ad - (int)ad) == 0) ? (int)ad : (int)(ad + 1));
        dist[d] = ad;
      }
      const double x = dist[0], y = dist[1], z = dist[2];
      const double w0 = __dmul_rn(1.0 - x, 1.0 - y), w1 = __dmul_rn(x, 1.0 - y), w2 = __dmul_rn(y, 1.0 - x), w3 = __dmul_rn(x, y);
      const int i = off[0], j = off[1], k = off[2];
      const double p1 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(w0, C(c[0], c[1], c[2])), __dmul_rn(w1, C(c[0] + i, c[1], c[2]))),
                                            __dmul_rn(w2, C(c[0], c[1] + j, c[2]))), __dmul_rn(w3, C(c[0] + i, c[1] + j, c[2])));
      const double p2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(w0, C(c[0], c[1], c[2] + k)), __dmul_rn(w1, C(c[0] + i, c[1], c[2] + k))),
                                            __dmul_rn(w2, C(c[0], c[1] + j, c[2] + k))), __dmul_rn(w3, C(c[0] + i, c[1] + j, c[2] + k)));
      out = __dadd_rn(__dmul_rn(1.0 - z, p1), __dmul_rn(z, p2));
    } else {
      int b[3];
      double w[3][3];
      for (int d = 0; d < 3; ++d) {
        int shift = 0;
        if ((c[d] - a.gridLo[d]) == 0) shift = 1;
        if ((a.gridHiM1[d] - c[d]) == 0) shift = -1;
        b[d] = c[d] + shift;
        const int rel = f[d] - b[d] * a.rr[d];
        // a shifted base cell puts the fine cell outside [0,rr): the reference then indexes its table out of range; such cells
        // (coarse cell == an extra cell) never occur for interior fine cells, keep the in-range distance of the unshifted cell
        const double x = norm_dist_cc(shift ? f[d] - c[d] * a.rr[d] : rel, a.rr[d]);
        w[d][0] = __dmul_rn(__dmul_rn(0.5, x), x - 1.0);
        w[d][1] = __dmul_rn(-(x + 1.0), x - 1.0);
        w[d][2] = __dmul_rn(__dmul_rn(0.5, x), x + 1.0);
      }
      double plane[3];
#pragma unroll
      for (int pz = 0; pz < 3; ++pz) {
        const int k = b[2] + pz - 1;
        double acc = __dmul_rn(__dmul_rn(w[0][0], w[1][0]), C(b[0] - 1, b[1] - 1, k));
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(w[0][1], w[1][0]), C(b[0], b[1] - 1, k)));
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(w[0][2], w[1][0]), C(b[0] + 1, b[1] - 1, k)));
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(w[0][0], w[1][1]), C(b[0] - 1, b[1], k)));
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(w[0][1], w[1][1]), C(b[0], b[1], k)));
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(w[0][2], w[1][1]), C(b[0] + 1, b[1], k)));
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(w[0][0], w[1][2]), C(b[0] - 1, b[1] + 1, k)));
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(w[0][1], w[1][2]), C(b[0], b[1] + 1, k)));
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(w[0][2], w[1][2]), C(b[0] + 1, b[1] + 1, k)));
        plane[pz] = acc;
      }
      out = __dadd_rn(__dadd_rn(__dmul_rn(w[2][0], plane[0]), __dmul_rn(w[2][1], plane[1])), __dmul_rn(w[2][2], plane[2]));
    }
    qF[(f[0] - a.ow.lo[0]) + (long long)a.ow.n[0] * ((f[1] - a.ow.lo[1]) + (long long)a.ow.n[1] * (long long)(f[2] - a.ow.lo[2]))] = out;
  }
}

// Arches RMCRT_Radiation::sigmaT4<T> (Arches/SourceTerms/RMCRT.cc:939-1016): gas-only or absk-weighted gas + particle emission
#define RMCRT_MAX_PARTGAS 16
struct PartGasPtrs { const double* absk[RMCRT_MAX_PARTGAS]; const double* temp[RMCRT_MAX_PARTGAS]; int n; };
template <class T>
__global__ void arches_sigmaT4_kernel(const double* __restrict__ gasVolFrac, const __grid_constant__ PartGasPtrs pg, int doPart,
                                      T* __restrict__ out, long long n, double sigma_over_pi) {
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
    double v;
    if (!doPart || !(gasVolFrac[c] > 1e-16)) {          // gas only, or a wall / intrusion cell (1005-1009)
      const double T2 = __dmul_rn(pg.temp[0][c], pg.temp[0][c]);
      v = __dmul_rn(__dmul_rn(sigma_over_pi, T2), T2);
    } else {                                            // 985-1003
      double sumT = 0.0, sumAbsk = 0.0;
      for (int i = 0; i < pg.n; ++i) {
        const double T2 = __dmul_rn(pg.temp[i][c], pg.temp[i][c]);
        const double ak = pg.absk[i][c];
        sumT = __dadd_rn(sumT, __dmul_rn(__dmul_rn(T2, T2), ak));
        sumAbsk = __dadd_rn(sumAbsk, ak);
      }
      v = 0.0;
      if (sumAbsk > 1e-16) v = __ddiv_rn(__dmul_rn(sigma_over_pi, sumT), sumAbsk);
    }
    out[c] = (T)v;
  }
}
// RMCRT_Radiation::sumAbsk<T> (RMCRT.cc:1062-1130)
template <class T>
__global__ void arches_sumAbsk_kernel(const double* __restrict__ gasVolFrac, const __grid_constant__ PartGasPtrs pg, Box w, Box interior,
                                      T* __restrict__ out) {
  const long long n = box_cells(w);
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
    const long long row = c / w.n[0];
    const int kk = (int)(row / w.n[1]);
    const int idx[3] = {w.lo[0] + (int)(c - row * w.n[0]), w.lo[1] + (int)(row - (long long)kk * w.n[1]), w.lo[2] + kk};
    bool in = true;
    for (int d = 0; d < 3; ++d) in = in && idx[d] >= interior.lo[d] && idx[d] < interior.lo[d] + interior.n[d];
    double v = 1.0;                                     // boundary faces, ExtraPlusEdgeCells: hard-wired 1 (1108-1120)
    if (in) {
      v = 0.0;
      const bool gas = gasVolFrac[c] > 1e-16;
      for (int i = 0; i < pg.n; ++i) v = gas ? __dadd_rn(v, pg.absk[i][c]) : 1.0;   // 1093-1104
    }
    out[c] = (T)v;
  }
}
// RMCRT_Radiation::stencilToDBLs (RMCRT.cc:1389-1423): Stencil7 {w,e,s,n,b,t,p} AoS -> six double fields
struct SixPtrs { double* p[6]; };   // E W N S T B
__global__ void stencil_to_dbls_kernel(const double* __restrict__ bf7, long long n, const __grid_constant__ SixPtrs o) {
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
    const double* me = bf7 + 7 * c;
    o.p[0][c] = me[1]; o.p[1][c] = me[0]; o.p[2][c] = me[3]; o.p[3][c] = me[2]; o.p[4][c] = me[5]; o.p[5][c] = me[4];
  }
}

// ---- roofline denominators ----
__global__ void bw_stream_kernel(const uint4* __restrict__ buf, long long n16, int inner, unsigned* sink) {
  unsigned acc = 0;
  for (int it = 0; it < inner; ++it) {
    // rotate the start so consecutive sweeps of one thread do not hit its own L1 lines
    const long long shift = (long long)it * 7919 * blockDim.x;
#pragma unroll 4
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n16; t += (long long)gridDim.x * blockDim.x) {
      long long u = t + shift;
      if (u >= n16) u -= n16 * (u / n16);
      const uint4 v = __ldg(buf + u);
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  }
  if (acc == 0x12345677u) *sink = acc;
}
// pattern 2: L2 -> SM bandwidth proper.  Eight independent 16-B loads in flight per thread, cached in L2 only (ld.global.cg), every warp
// request four whole 128-B lines, consecutive sweeps rotated so nothing is served from L1: what the L2 slices and the crossbar deliver
// when the SMs ask as hard as they can (the denominator of the L2 roofline; B300_MICROARCH.md quotes ~6300 B/clk for the chip)
__global__ void bw_stream_cg_kernel(const uint4* __restrict__ buf, long long n16, int inner, unsigned* sink) {
  unsigned acc = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int it = 0; it < inner; ++it) {
    const long long shift = (long long)it * 7919 * blockDim.x;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n16; t += 8 * stride) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        long long w = t + u * stride + shift;
        if (w >= n16) w -= n16 * (w / n16);
        v[u] = __ldcg(buf + w);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
  }
  if (acc == 0x12345677u) *sink = acc;
}
__global__ void bw_gather_kernel(const float2* __restrict__ buf, unsigned nRec, int perThread, unsigned* sink) {
  unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0;
#pragma unroll 4
  for (int i = 0; i < perThread; ++i) {
    x = x * 1664525u + 1013904223u;                         // LCG: next index independent of loaded data
    const unsigned idx = (unsigned)(((unsigned long long)(x ^ (x >> 15)) * nRec) >> 32);
    const float2 v = __ldg(buf + idx);
    acc += v.x + v.y;
  }
  if (acc == 123.456f) *sink = 1;
}

}  // namespace

//======================================================================================================
// host side
//======================================================================================================
struct LevelState {
  bool defined = false;
  bool finalized = false;
  bool wallsEmit = true;
  rmcrt_level_desc d{};
  int recordKind = 0;
  void* base = nullptr;   // allocation (guard records in front of and behind the level)
  void* rec = nullptr;    // record of cell `lo`
  void* recFull = nullptr;   // recordKind 2: float records of the same cells (epilogue of the divQ kernel)
  long long nCells = 0;
  size_t recBytes = 8;
  bool propsFloat = false;   // the last pack read float property arrays
  // blocked half-packed copy for the independent-ray march (indep.cuh), built on demand after the level last changed
  void* blk = nullptr;
  size_t blkCap = 0;
  bool blkValid = false;
  int blkDims[3] = {0, 0, 0};
};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return -1;
    cap = bytes;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct rmcrt_ctx {
  int device = 0;
  int smCount = 148;
  rmcrt_params p{};
  LevelState lv[RMCRT_MAX_LEVELS];
  unsigned long long* d_counters = nullptr;   // NCOUNTERS
  unsigned* d_lstat = nullptr;                // [RMCRT_MAX_LEVELS][LSTAT_WORDS]
  int* d_nItems = nullptr;
  DevBuf stage[3];      // upload staging
  DevBuf outDivQ, outVolq, outFlux;   // host-API output staging
  DevBuf fluxFlags, fluxItems, compactTmp, dbg[6], misc, roiBuf;
  DevBuf lhcDivQ, lhcFlux;            // LHC permutation tables of the regression modes
  int lhcKeyDivQ[4] = {0, 0, 0, 0};   // (nRays, sumLo, rows, mode) the tables were built for
  int lhcKeyFlux[4] = {0, 0, 0, 0};
  DevBuf outVR[2];                    // host-API radiometer output staging
  DevBuf regionBuf[MAXL];             // Data Onion: blocked region boxes of the levels between 0 and the fine one
  DevBuf sortScratch;                 // divq_sorted_kernel: tile counter + per-block res / ids / keys
  DevBuf replSend, replGathered;      // rmcrt_replicate_level: this device's packed slab / the gathered slabs of all devices
  std::string err;
  void* persistedStreamFor = nullptr;
  void* persistedPtr = nullptr;
};

namespace {

std::string g_createErr;
std::mutex g_mu;

int fail(rmcrt_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else { std::lock_guard<std::mutex> l(g_mu); g_createErr = buf; }
  return code;
}

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(ctx, RMCRT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

Box mkbox(const int32_t lo[3], const int32_t hi[3]) {
  Box b;
  for (int d = 0; d < 3; ++d) { b.lo[d] = lo[d]; b.n[d] = hi[d] - lo[d]; }
  return b;
}

int grid_for(long long n, int block, int sm) {
  long long g = (n + block - 1) / block;
  long long cap = (long long)sm * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

int check_level(rmcrt_ctx* ctx, int L) {
  if (!ctx) return RMCRT_ERR_INVALID;
  if (L < 0 || L >= RMCRT_MAX_LEVELS || !ctx->lv[L].defined) return fail(ctx, RMCRT_ERR_INVALID, "level %d is not defined", L);
  return 0;
}

int validate_params(rmcrt_ctx* ctx, const rmcrt_params* p) {
  if (p->nDivQRays < 1 || p->nFluxRays < 1) return fail(ctx, RMCRT_ERR_INVALID, "nDivQRays and nFluxRays must be >= 1");
  if (p->nFluxRays >= (1 << 28)) return fail(ctx, RMCRT_ERR_INVALID, "nFluxRays too large");
  if (!(p->threshold >= 0)) return fail(ctx, RMCRT_ERR_INVALID, "Threshold must be >= 0");
  if (p->marchPrecision != 0 && p->marchPrecision != 1) return fail(ctx, RMCRT_ERR_INVALID, "marchPrecision must be 0 (fp32) or 1 (fp64)");
  if (p->rayDirSampleAlgo != 0 && p->rayDirSampleAlgo != 1) return fail(ctx, RMCRT_ERR_INVALID, "rayDirSampleAlgo must be 0 (naive) or 1 (LatinHyperCube)");
  if (p->maxRayLength != p->maxRayLength) return fail(ctx, RMCRT_ERR_INVALID, "maxRayLength is NaN");
  if (!(p->sigmaScat >= 0)) return fail(ctx, RMCRT_ERR_INVALID, "sigmaScat must be >= 0");
  return 0;
}

void fill_level_geom(const rmcrt_ctx* ctx, int L, LevelGeom& g) {
  const LevelState& s = ctx->lv[L];
  g.rec = s.rec;
  g.recFull = s.recFull;
  for (int d = 0; d < 3; ++d) {
    g.lo[d] = s.d.lo[d]; g.n[d] = s.d.hi[d] - s.d.lo[d];
    g.intLo[d] = s.d.interiorLo[d]; g.intHi[d] = s.d.interiorHi[d];
    g.rr[d] = s.d.refineRatio[d] > 0 ? s.d.refineRatio[d] : 1;
    g.regLo[d] = s.d.interiorLo[d]; g.regHi[d] = s.d.interiorHi[d];
    g.dx[d] = s.d.dx[d];
    g.anchor[d] = s.d.anchor[d];
    g.off[d] = 0.0;
  }
}

// Level::mapCellToCoarser (Level.cc:1285-1307)
int map_coarser_host(int i, int rr) {
  int q = i / rr;
  if (i < 0 && rr > 1) q += (int)fmod((double)i, (double)rr);
  return q;
}

// Ray::computeExtents, patch_based and fixed ROI (Ray.cc:1303-1357)
void compute_extents(const rmcrt_ctx* ctx, int maxLevels, const int32_t pLo[3], const int32_t pHi[3], int32_t roiLo[3],
                     int32_t roiHi[3], int32_t* regLo, int32_t* regHi) {
  const int F = maxLevels - 1;
  const rmcrt_level_desc& f = ctx->lv[F].d;
  for (int d = 0; d < 3; ++d) {
    const int lo = ctx->p.roiFixed ? ctx->p.roiFixedLo[d] : pLo[d] - ctx->p.haloCells[d];
    const int hi = ctx->p.roiFixed ? ctx->p.roiFixedHi[d] : pHi[d] + ctx->p.haloCells[d];
    roiLo[d] = std::max(lo, f.lo[d]);
    roiHi[d] = std::min(hi, f.hi[d]);
    const int ec = f.interiorLo[d] - f.lo[d];
    regLo[3 * F + d] = roiLo[d] + ec;
    regHi[3 * F + d] = roiHi[d] - ec;
    regLo[d] = ctx->lv[0].d.interiorLo[d];
    regHi[d] = ctx->lv[0].d.interiorHi[d];
  }
  for (int L = maxLevels - 2; L > 0; --L) {
    const rmcrt_level_desc& l = ctx->lv[L].d;
    for (int d = 0; d < 3; ++d) {
      const int rr = l.refineRatio[d] > 0 ? l.refineRatio[d] : 1;
      regLo[3 * L + d] = std::max(map_coarser_host(regLo[3 * (L + 1) + d], rr) - ctx->p.haloCells[d], l.interiorLo[d]);
      regHi[3 * L + d] = std::min(map_coarser_host(regHi[3 * (L + 1) + d], rr) + ctx->p.haloCells[d], l.interiorHi[d]);
    }
  }
}

unsigned* lstat_of(const rmcrt_ctx* ctx, int L) { return ctx->d_lstat + (size_t)L * LSTAT_WORDS; }

// every modification of a level clears its device-side "valid" word (stream-ordered)
int invalidate_level(rmcrt_ctx* ctx, int L, cudaStream_t st) {
  ctx->lv[L].finalized = false;
  ctx->lv[L].blkValid = false;
  CK(cudaMemsetAsync(lstat_of(ctx, L) + 3, 0, sizeof(unsigned), st));
  return 0;
}

int pick_warps(const rmcrt_ctx* ctx, long long nTiles) {
  if (ctx->p.warpsPerTile > 0) {
    int w = ctx->p.warpsPerTile;
    return w > 8 ? 8 : w;
  }
  // 128-thread blocks (4 warps share a tile's rays) reach the 64 resident warps per SM that 32 registers allow; small
  // patches get more warps per tile so the machine is still filled ~4x over
  const long long want = (long long)ctx->smCount * 64 * 4;
  int w = 4;
  while (w < 8 && nTiles * w < want) w *= 2;
  return w;
}

void apply_l2_window_raw(rmcrt_ctx* ctx, void* base, size_t bytes, cudaStream_t st) {
  if (!ctx->p.l2Persist) return;
  int maxWin = 0, maxPersist = 0;
  cudaDeviceGetAttribute(&maxWin, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
  cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
  if (maxWin <= 0 || maxPersist <= 0) return;
  const size_t win = std::min(bytes, (size_t)maxWin);
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min(win, (size_t)maxPersist));
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof attr);
  attr.accessPolicyWindow.base_ptr = base;
  attr.accessPolicyWindow.num_bytes = win;
  attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)maxPersist / (double)win);
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
  cudaGetLastError();
}
// the persisting window follows the buffer the march reads: the level's records, or its blocked copy (one window per stream)
void persist_in_l2(rmcrt_ctx* ctx, void* base, size_t bytes, cudaStream_t st) {
  if (!ctx->p.l2Persist || (ctx->persistedStreamFor == (void*)st && ctx->persistedPtr == base)) return;
  apply_l2_window_raw(ctx, base, bytes, st);
  ctx->persistedStreamFor = (void*)st;
  ctx->persistedPtr = base;
}

template <class T, class RecT>
int launch_pack(rmcrt_ctx* ctx, const void* a, const void* s, const int32_t* ct, Box src, Box region, void* dst, Box dstBox,
                unsigned* status, void* full, Box fullBox, cudaStream_t st) {
  const long long n = box_cells(region);
  if (n <= 0) return 0;
  pack_kernel<T, RecT><<<grid_for(n, 256, ctx->smCount), 256, 0, st>>>(static_cast<const T*>(a), static_cast<const T*>(s), ct, src,
                                                                     region, static_cast<RecT*>(dst), dstBox, ctx->p.flowCell,
                                                                     status, static_cast<Rec8*>(full), fullBox);
  CK(cudaGetLastError());
  return 0;
}

// full / fullBox: float copy of the level (recordKind 2 only), written alongside the half-packed records
int do_pack(rmcrt_ctx* ctx, int recordKind, int isFloat, const void* a, const void* s, const int32_t* ct, Box src, Box region,
            void* dst, Box dstBox, unsigned* status, void* full, Box fullBox, cudaStream_t st) {
  if (recordKind == 0) return isFloat ? launch_pack<float, Rec8>(ctx, a, s, ct, src, region, dst, dstBox, status, nullptr, fullBox, st)
                                      : launch_pack<double, Rec8>(ctx, a, s, ct, src, region, dst, dstBox, status, nullptr, fullBox, st);
  if (recordKind == 2) return isFloat ? launch_pack<float, Rec4>(ctx, a, s, ct, src, region, dst, dstBox, status, full, fullBox, st)
                                      : launch_pack<double, Rec4>(ctx, a, s, ct, src, region, dst, dstBox, status, full, fullBox, st);
  return isFloat ? launch_pack<float, Rec16>(ctx, a, s, ct, src, region, dst, dstBox, status, nullptr, fullBox, st)
                 : launch_pack<double, Rec16>(ctx, a, s, ct, src, region, dst, dstBox, status, nullptr, fullBox, st);
}

// MT19937 exactly as MTRand (Core/Math/MersenneTwister.h:217-258,339-368): the regression modes of the LHC sampler shuffle
// with randInt(n), which draws an unbounded number of words, so the stateless mt19937_first<> does not cover it.
struct HostMT {
  uint32_t st[624];
  int next = 624;
  void seed(uint32_t s) {
    st[0] = s;
    for (int i = 1; i < 624; ++i) st[i] = 1812433253u * (st[i - 1] ^ (st[i - 1] >> 30)) + (uint32_t)i;
    reload();
  }
  static uint32_t tw(uint32_t m, uint32_t s0, uint32_t s1) {
    return m ^ (((s0 & 0x80000000u) | (s1 & 0x7fffffffu)) >> 1) ^ ((s1 & 1u) ? 0x9908b0dfu : 0u);
  }
  void reload() {
    int k = 0;
    for (; k < 624 - 397; ++k) st[k] = tw(st[k + 397], st[k], st[k + 1]);
    for (; k < 623; ++k) st[k] = tw(st[k + 397 - 624], st[k], st[k + 1]);
    st[623] = tw(st[396], st[623], st[0]);
    next = 0;
  }
  uint32_t u32() {
    if (next >= 624) reload();
    uint32_t y = st[next++];
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u;
    return y ^ (y >> 18);
  }
  uint32_t randInt(uint32_t n) {     // MersenneTwister.h:233-250
    uint32_t used = n;
    used |= used >> 1; used |= used >> 2; used |= used >> 4; used |= used >> 8; used |= used >> 16;
    uint32_t i;
    do i = u32() & used; while (i > n);
    return i;
  }
};

// RMCRTCommon::randVector (RMCRTCommon.cc:998-1024) for every value of i+j+k on the level (regression modes only)
int ensure_lhc_table(rmcrt_ctx* ctx, DevBuf& buf, int key[4], int nRays, const rmcrt_level_desc& f, cudaStream_t st) {
  const int mode = ctx->p.fixedRandom ? 2 : 1;
  const int sumLo = f.lo[0] + f.lo[1] + f.lo[2];
  const int rows = ctx->p.fixedRandom ? 1 : (f.hi[0] + f.hi[1] + f.hi[2] - 3) - sumLo + 1;
  if (buf.p && key[0] == nRays && key[1] == sumLo && key[2] == rows && key[3] == mode) return 0;
  std::vector<int> tab((size_t)rows * nRays);
  HostMT mt;
  for (int r = 0; r < rows; ++r) {
    int* a = tab.data() + (size_t)r * nRays;
    for (int i = 0; i < nRays; ++i) a[i] = i;
    if (mode == 1) mt.seed((uint32_t)(sumLo + r));                   // mTwister.seed(cell.x()+cell.y()+cell.z())
    for (int i = nRays - 1; i > 0; --i) {
      const int ri = mode == 2 ? (int)(0.3 * i) : (int)mt.randInt((uint32_t)i);
      std::swap(a[i], a[ri]);
    }
  }
  if (buf.ensure(tab.size() * sizeof(int))) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc LHC table");
  CK(cudaMemcpyAsync(buf.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));      // `tab` is pageable and goes out of scope
  key[0] = nRays; key[1] = sumLo; key[2] = rows; key[3] = mode;
  return 0;
}

int build_args(rmcrt_ctx* ctx, int maxLevels, bool multiLevel, const int32_t pLo[3], const int32_t pHi[3], const int32_t oLo[3],
               const int32_t oHi[3], uint64_t timestep, TraceArgs& a) {
  memset(&a, 0, sizeof a);
  const int F = maxLevels - 1;
  if (maxLevels < 1 || maxLevels > MAXL) return fail(ctx, RMCRT_ERR_INVALID, "maxLevels %d not in [1,%d]", maxLevels, MAXL);
  for (int L = multiLevel ? 0 : F; L <= F; ++L) {
    int rc = check_level(ctx, L);
    if (rc) return rc;
    if (ctx->lv[L].recordKind != ctx->lv[F].recordKind) return fail(ctx, RMCRT_ERR_INVALID, "all levels must use one recordKind");
  }
  if ((ctx->p.marchPrecision == 1) != (ctx->lv[F].recordKind == 1))
    return fail(ctx, RMCRT_ERR_INVALID, "marchPrecision fp64 needs recordKind 1 (double records) and fp32 needs recordKind 0 or 2");
  a.G.maxLevels = maxLevels;
  for (int L = 0; L <= F; ++L) if (ctx->lv[L].defined) fill_level_geom(ctx, L, a.G.lev[L]);
  const rmcrt_level_desc& f = ctx->lv[F].d;
  for (int d = 0; d < 3; ++d) {
    if (pLo[d] < f.interiorLo[d] || pHi[d] > f.interiorHi[d] || pHi[d] <= pLo[d])
      return fail(ctx, RMCRT_ERR_INVALID, "patch [%d,%d) outside the interior of level %d on axis %d", pLo[d], pHi[d], F, d);
    if (oLo[d] > pLo[d] || oHi[d] < pHi[d]) return fail(ctx, RMCRT_ERR_INVALID, "output window does not contain the patch");
    a.patchLo[d] = pLo[d]; a.patchN[d] = pHi[d] - pLo[d];
    a.outLo[d] = oLo[d]; a.outN[d] = oHi[d] - oLo[d];
  }
  a.multiLevel = multiLevel ? 1 : 0;
  if (multiLevel) {
    int32_t roiLo[3], roiHi[3], regLo[3 * RMCRT_MAX_LEVELS], regHi[3 * RMCRT_MAX_LEVELS];
    compute_extents(ctx, maxLevels, pLo, pHi, roiLo, roiHi, regLo, regHi);
    for (int d = 0; d < 3; ++d) { a.G.roiLo[d] = roiLo[d]; a.G.roiHi[d] = roiHi[d]; }
    for (int L = 0; L <= F; ++L) for (int d = 0; d < 3; ++d) { a.G.lev[L].regLo[d] = regLo[3 * L + d]; a.G.lev[L].regHi[d] = regHi[3 * L + d]; }
    for (int L = 0; L <= F; ++L) for (int d = 0; d < 3; ++d) a.G.lev[L].off[d] = ctx->lv[L].d.anchor[d] - f.anchor[d];
  }
  a.nCells = (long long)a.patchN[0] * a.patchN[1] * a.patchN[2];
  if (a.nCells * 6 >= (1LL << 31)) return fail(ctx, RMCRT_ERR_INVALID, "patch too large (%lld cells); split it", a.nCells);
  a.nTiles = (a.nCells + 31) / 32;
  a.threshold = ctx->p.threshold;
  a.allowReflect = ctx->p.allowReflect;
  a.skipZero = ctx->p.skipZeroWeightLegs ? 1 : 0;   // the kernels AND this with the device-side level status
  a.lstat = ctx->d_lstat;
  a.firstLevel = multiLevel ? 0 : F;
  a.recBytes = (int)ctx->lv[F].recBytes;
  a.fluxFlowGuard = multiLevel ? 0 : 1;
  a.rng.key0 = (uint32_t)ctx->p.seed;
  a.rng.key1 = (uint32_t)(ctx->p.seed >> 32) ^ (uint32_t)(timestep >> 32);
  a.rng.timestep = (uint32_t)timestep;
  a.rng.isSeedRandom = ctx->p.isSeedRandom;
  a.rng.fixedRandom = ctx->p.fixedRandom;
  a.rng.ccRays = ctx->p.CCRays;
  a.rng.coherent = ctx->p.rayCoherence;
  a.rng.lhc = ctx->p.rayDirSampleAlgo == 1 ? 1 : 0;
  a.counters = ctx->d_counters;
  const bool bounded = ctx->p.maxRayLength > 0 && ctx->p.maxRayLength < 1e300;
  if (bounded && multiLevel)
    return fail(ctx, RMCRT_ERR_INVALID, "maxRayLength (boundedRayLength ROI) applies to single-level solves only (Ray.cc:2349-2627 has no length test)");
  a.maxRayLength = bounded ? ctx->p.maxRayLength : (double)INFINITY;
  a.floatProps = ctx->lv[F].propsFloat ? 1 : 0;
  // isotropic scattering lives in RMCRTCommon::updateSumI only (RMCRTCommon.cc:670-676, 790-836): the Data-Onion march
  // (Ray::updateSumI_ML) has no scattering code, so multi-level solves ignore <sigmaScat> exactly like the reference
  a.sigmaScat = (!multiLevel && ctx->p.sigmaScat > 0) ? ctx->p.sigmaScat : 0.0;
  return 0;
}

// does this solve run on the blocked copy of its level (blockedMarch 1..3: divq_indep / divq_blk / divq_sorted_kernel)?  Production
// FP32 / Philox single-level divQ rays with the default sampling; everything else marches the level's own records.
bool blocked_default(const rmcrt_ctx* ctx, const TraceArgs& a, bool ML, bool divQ = true) {
  if (ML || (divQ && !ctx->p.solveDivQ) || ctx->p.marchPrecision != 0 || ctx->p.rayCoherence != 0 || ctx->p.blockedMarch <= 0 || ctx->p.genericMarch) return false;
  if (!ctx->p.isSeedRandom || ctx->p.fixedRandom || ctx->p.rayDirSampleAlgo != 0) return false;       // f32gen
  if (a.maxRayLength < 1e300 || a.sigmaScat > 0) return false;                                        // the generic march
  const rmcrt_level_desc& d = ctx->lv[a.G.maxLevels - 1].d;
  long long nb = 1;
  for (int k = 0; k < 3; ++k) nb *= (d.hi[k] - d.lo[k] + BLK_PAD + 2 + 1) / 2;
  return nb < (1LL << 28);      // 32-bit word index
}

// ... and the 2-level Data Onion: the same kernel on blocked copies of the patch's ROI and of level 0 (divq_sorted_kernel<ONION>)
// Boxes of the levels above 0 that the Data-Onion march copies verbatim: the fine level's ROI, a coarser level's region, each with the
// level's extra cells where it ends at the domain boundary.
void onion_copy_box(const rmcrt_ctx* ctx, const TraceArgs& a, int L, int lo[3], int hi[3]) {
  const int F = a.G.maxLevels - 1;
  const rmcrt_level_desc& d = ctx->lv[L].d;
  for (int k = 0; k < 3; ++k) {
    // (a cell outside the box but not inside the domain is a wall of THIS level, not a jump: `inDomain`, Ray.cc:2486-2493)
    const int rlo = L == F ? a.G.roiLo[k] : a.G.lev[L].regLo[k], rhi = L == F ? a.G.roiHi[k] : a.G.lev[L].regHi[k];
    lo[k] = rlo <= d.interiorLo[k] ? d.lo[k] : rlo;
    hi[k] = rhi >= d.interiorHi[k] ? d.hi[k] : rhi;
  }
}
bool onion_blocked_default(const rmcrt_ctx* ctx, const TraceArgs& a, bool ML) {
  if (!ML || a.G.maxLevels < 2 || a.G.maxLevels > ONION_MAXL || !ctx->p.solveDivQ || ctx->p.marchPrecision != 0 || ctx->p.rayCoherence != 0 ||
      ctx->p.blockedMarch != 3 || ctx->p.genericMarch) return false;
  if (!ctx->p.isSeedRandom || ctx->p.fixedRandom || ctx->p.rayDirSampleAlgo != 0) return false;
  if (ctx->p.nDivQRays < 12) return false;      // too few rays per tile to sort one tile at a time (no tile groups here): the one-cell kernels
  for (int d = 0; d < 3; ++d)      // the patch lies inside the ROI (a fixed / dynamic ROI need not contain it)
    if (!(a.G.roiLo[d] <= a.patchLo[d] && a.patchLo[d] + a.patchN[d] <= a.G.roiHi[d])) return false;
  // a ray that leaves the box of level L through an interior cell must land INSIDE the box of level L - 1 (the march tests "marker cell",
  // the reference the stepped axis only: the same thing as long as every jump ends inside the next box)
  for (int L = a.G.maxLevels - 1; L >= 2; --L) {
    int lo[3], hi[3], clo[3], chi[3];
    onion_copy_box(ctx, a, L, lo, hi);
    onion_copy_box(ctx, a, L - 1, clo, chi);
    const rmcrt_level_desc& d = ctx->lv[L].d;
    for (int k = 0; k < 3; ++k) {
      const int rr = d.refineRatio[k] > 0 ? d.refineRatio[k] : 1;
      if (lo[k] - 1 >= d.interiorLo[k] && map_coarser_host(lo[k] - 1, rr) < clo[k]) return false;
      if (hi[k] < d.interiorHi[k] && map_coarser_host(hi[k], rr) >= chi[k]) return false;
    }
  }
  return true;
}

// Origin tiles of divq_sorted_kernel: 4 x 4 x 2 cells instead of 32 consecutive cells of a row.  The rays of a tile share sectors while
// they are near their origin, and a compact tile has half the neighbourhood of a row of 32: 128^3 30.6 -> 27.8 ms, 256^3 536 -> 499 ms
// (8 x 2 x 2: 28.0, 4 x 2 x 4: 27.8, 2 x 4 x 4: 27.9, 8 x 4 x 1: 28.4, 16 x 2 x 1: 29.1; profiles/r02v48).  Edge tiles of patches whose
// sizes are not multiples of the shape have idle lanes.
void compact_tiles(TraceArgs& a) {
  a.tsx = 4; a.tsy = 4; a.tsz = 2;
  a.tcx = (a.patchN[0] + a.tsx - 1) / a.tsx; a.tcy = (a.patchN[1] + a.tsy - 1) / a.tsy; a.tcz = (a.patchN[2] + a.tsz - 1) / a.tsz;
  a.nTiles = (long long)a.tcx * a.tcy * a.tcz;
}

// blocked half-packed copy of level L (indep.cuh), rebuilt when the level changed since the last build (stream-ordered)
template <class RecT>
int ensure_blocked_level(rmcrt_ctx* ctx, int L, BlkGeom& bg, cudaStream_t st, bool persist = true) {
  LevelState& s = ctx->lv[L];
  int nb[3];
  for (int d = 0; d < 3; ++d) nb[d] = (s.d.hi[d] - s.d.lo[d] + BLK_PAD + 2 + 1) / 2;        // cells [lo - BLK_PAD, hi + 2) in 2-cell blocks
  const size_t bytes = (size_t)nb[0] * nb[1] * nb[2] * 32;
  if ((long long)nb[0] * nb[1] * nb[2] >= (1LL << 28)) return fail(ctx, RMCRT_ERR_INVALID, "level too large for the blocked march");   // 32-bit word index
  if (bytes > s.blkCap) {
    if (s.blk) cudaFree(s.blk);
    s.blk = nullptr; s.blkCap = 0;
    CK(cudaMalloc(&s.blk, bytes));
    s.blkCap = bytes;
    s.blkValid = false;
  }
  if (!s.blkValid || s.blkDims[0] != nb[0] || s.blkDims[1] != nb[1] || s.blkDims[2] != nb[2]) {
    if constexpr (!std::is_same<RecT, Rec16>::value) {
      blk_build_kernel<RecT><<<grid_for((long long)nb[0] * nb[1] * nb[2], 256, ctx->smCount), 256, 0, st>>>(
          static_cast<const RecT*>(s.rec), mkbox(s.d.lo, s.d.hi), static_cast<uint32_t*>(s.blk), nb[0], nb[1], nb[2]);
      CK(cudaGetLastError());
    }
    s.blkDims[0] = nb[0]; s.blkDims[1] = nb[1]; s.blkDims[2] = nb[2];
    s.blkValid = true;
  }
  if (persist) persist_in_l2(ctx, s.blk, bytes, st);
  bg.blk = static_cast<const uint32_t*>(s.blk);
  bg.bx = nb[0]; bg.bxy = nb[0] * nb[1];
  for (int d = 0; d < 3; ++d) bg.lo[d] = s.d.lo[d] - BLK_PAD;
  return 0;
}

// The same single-level production solve when reflective walls keep the rays going (launch_gate 2): translates separate at the first
// wall they do not reach at the same step and stay apart, so the four-cell kernel loses its point.  Pairs stacked in z reach every wall
// normal to x or y together and go on jointly after the reflection (REJOIN); a black or z wall ends the joint march.
template <class RecT>
void launch_reflective_pairs(rmcrt_ctx* ctx, TraceArgs a2, long long cap, cudaStream_t st) {
  if constexpr (!std::is_same<RecT, Rec16>::value) {
    const int nxc = (a2.patchN[0] + 31) / 32;
    long long g2 = (long long)nxc * a2.patchN[1] * ((a2.patchN[2] + 1) / 2);
    if (g2 > cap) g2 = cap;
    int W2 = ctx->p.warpsPerTile > 0 ? std::min(ctx->p.warpsPerTile, 8) : 4;
    if (W2 < 2) W2 = 2;        // cells A and B are written by warps 0 and 1
    a2.gate = 2;
    divq_dual_kernel<RecT, true><<<(int)g2, 32 * W2, W2 * 64 * sizeof(double), st>>>(a2, nxc, a2.patchN[1], 1);
  }
}

template <class Real, class RecT>
int launch_trace(rmcrt_ctx* ctx, TraceArgs& a, bool ML, cudaStream_t st) {
  // boundedRayLength ROI and scattering: the generic march carries the length test / the scattering events
  const bool bounded = a.maxRayLength < 1e300 || a.sigmaScat > 0;
  const bool lhcTable = a.rng.lhc && (!a.rng.isSeedRandom || a.rng.fixedRandom);
  // (the scattering march needs the absolute ray origin, which only the FP64 ray set-up provides)
  const bool f32gen = std::is_same<Real, float>::value && a.rng.isSeedRandom && !a.rng.fixedRandom && !a.rng.lhc && !(a.sigmaScat > 0);
  // ---- boundary flux first, like the task body (Ray.cc:645 before 742)
  if (ctx->p.solveBoundaryFlux) {
    if (!a.bflux) return fail(ctx, RMCRT_ERR_INVALID, "solveBoundaryFlux needs a boundFlux buffer");
    if (ctx->fluxFlags.ensure((size_t)a.nCells * 6)) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc flux flags");
    if (ctx->fluxItems.ensure((size_t)a.nCells * 6 * sizeof(int))) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc flux items");
    unsigned char* flags = static_cast<unsigned char*>(ctx->fluxFlags.p);
    flux_mask_kernel<RecT><<<grid_for(a.nCells, 256, ctx->smCount), 256, 0, st>>>(a, flags);
    CK(cudaGetLastError());
    const int nFlag = (int)(a.nCells * 6);
    const int nChunks = (nFlag + COMPACT_CHUNK - 1) / COMPACT_CHUNK;
    if (ctx->compactTmp.ensure((size_t)nChunks * sizeof(int))) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc compaction counts");
    int* chunk = static_cast<int*>(ctx->compactTmp.p);
    compact_count_kernel<<<nChunks, 256, 0, st>>>(flags, nFlag, chunk);
    compact_scan_kernel<<<1, 1024, 0, st>>>(chunk, nChunks, ctx->d_nItems);
    compact_scatter_kernel<<<nChunks, 256, 0, st>>>(flags, nFlag, chunk, static_cast<int*>(ctx->fluxItems.p));
    CK(cudaGetLastError());
    TraceArgs f = a;
    if (!bounded && blocked_default(ctx, a, ML, false)) {      // boundary-flux rays of the production solve: the blocked copy too
      if constexpr (std::is_same<Real, float>::value && !std::is_same<RecT, Rec16>::value) {
        BlkGeom bgF;
        int rc = ensure_blocked_level<RecT>(ctx, a.G.maxLevels - 1, bgF, st);
        if (rc) return rc;
        f.fluxBlk = bgF.blk; f.fluxBx = bgF.bx; f.fluxBxy = bgF.bxy;
        for (int d = 0; d < 3; ++d) f.fluxLo[d] = bgF.lo[d];
      }
    }
    f.items = static_cast<const int*>(ctx->fluxItems.p);
    f.nItems = ctx->d_nItems;
    f.nRays = ctx->p.nFluxRays;
    if (lhcTable) {
      int rc = ensure_lhc_table(ctx, ctx->lhcFlux, ctx->lhcKeyFlux, f.nRays, ctx->lv[a.G.maxLevels - 1].d, st);
      if (rc) return rc;
      f.rng.lhcPerm = static_cast<const int*>(ctx->lhcFlux.p);
      f.rng.lhcSumLo = ctx->lhcKeyFlux[1]; f.rng.lhcRows = ctx->lhcKeyFlux[2];
    }
    const int W = ctx->p.warpsPerTile > 0 ? std::min(ctx->p.warpsPerTile, 8) : (ctx->p.nFluxRays >= 8 ? 8 : 1);
    const int grid = ctx->smCount * 8;
    if (ML) flux_trace_kernel<Real, RecT, true><<<grid, 32 * W, W * 64 * sizeof(double), st>>>(f);
    else if (bounded) flux_trace_kernel<Real, RecT, false, false><<<grid, 32 * W, W * 64 * sizeof(double), st>>>(f);
    else    flux_trace_kernel<Real, RecT, false><<<grid, 32 * W, W * 64 * sizeof(double), st>>>(f);
    CK(cudaGetLastError());
  }
  if (ctx->p.solveDivQ) {
    a.nRays = ctx->p.nDivQRays;
    if (lhcTable) {
      int rc = ensure_lhc_table(ctx, ctx->lhcDivQ, ctx->lhcKeyDivQ, a.nRays, ctx->lv[a.G.maxLevels - 1].d, st);
      if (rc) return rc;
      a.rng.lhcPerm = static_cast<const int*>(ctx->lhcDivQ.p);
      a.rng.lhcSumLo = ctx->lhcKeyDivQ[1]; a.rng.lhcRows = ctx->lhcKeyDivQ[2];
    }
    int W = pick_warps(ctx, a.nTiles);
    if (W > a.nRays) W = 1;
    long long grid = a.nTiles;
    // one tile per block up to a large cap: the hardware block scheduler then balances the uneven ray lengths (a
    // grid-stride loop over ~7 tiles per block left a ~10 % tail at 128^3)
    const long long cap = ctx->p.maxBlocksPerSM > 0 ? (long long)ctx->smCount * ctx->p.maxBlocksPerSM : (1LL << 22);
    if (grid > cap) grid = cap;
    bool patchInRoi = true;       // a fixed ROI need not contain the patch; the ROI-copy fast path starts every ray inside the copy
    for (int d = 0; d < 3 && ML; ++d) patchInRoi = patchInRoi && a.G.roiLo[d] <= a.patchLo[d] && a.patchLo[d] + a.patchN[d] <= a.G.roiHi[d];
    if (onion_blocked_default(ctx, a, ML)) {
      // Data Onion, default sampling: length-sorted rays on blocked copies -- level 0 whole, every level above it the box the rays can
      // visit (ROI / region) wrapped in marker cells
      if constexpr (std::is_same<Real, float>::value && !std::is_same<RecT, Rec16>::value) {
        const int F = a.G.maxLevels - 1;
        compact_tiles(a);
        OnionBlk ob;
        memset(&ob, 0, sizeof ob);
        size_t fineBytes = 0;
        for (int L = F; L >= 1; --L) {
          const LevelState& ls = ctx->lv[L];
          int lo[3], hi[3], nbk[3];
          onion_copy_box(ctx, a, L, lo, hi);
          Box box;
          long long nBlocks = 1;
          for (int d = 0; d < 3; ++d) {
            box.lo[d] = lo[d]; box.n[d] = hi[d] - lo[d];
            ob.lo[L][d] = lo[d] - 1 - 2;                         // marker shell, then one guard block
            nbk[d] = (box.n[d] + 2 + 4 + 1) / 2;
            nBlocks *= nbk[d];
          }
          if (nBlocks >= (1LL << 28)) return fail(ctx, RMCRT_ERR_INVALID, "ROI too large");
          DevBuf& buf = L == F ? ctx->roiBuf : ctx->regionBuf[L];
          if (buf.ensure((size_t)nBlocks * 32)) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc ROI buffer");
          roi_blk_build_kernel<RecT><<<grid_for(nBlocks, 256, ctx->smCount), 256, 0, st>>>(
              static_cast<const RecT*>(ls.rec), mkbox(ls.d.lo, ls.d.hi), static_cast<uint32_t*>(buf.p), nbk[0], nbk[1], nbk[2], ob.lo[L][0],
              ob.lo[L][1], ob.lo[L][2], box);
          CK(cudaGetLastError());
          ob.blk[L] = static_cast<const uint32_t*>(buf.p); ob.bx[L] = nbk[0]; ob.bxy[L] = nbk[0] * nbk[1];
          if (L == F) fineBytes = (size_t)nBlocks * 32;
        }
        BlkGeom bgC;
        int rc = ensure_blocked_level<RecT>(ctx, 0, bgC, st, false);
        if (rc) return rc;
        ob.blk[0] = bgC.blk; ob.bx[0] = bgC.bx; ob.bxy[0] = bgC.bxy;
        for (int d = 0; d < 3; ++d) ob.lo[0][d] = bgC.lo[d];
        persist_in_l2(ctx, ctx->roiBuf.p, fineBytes, st);
        BlkGeom bg;      // the box of the origin cells
        bg.blk = ob.blk[F]; bg.bx = ob.bx[F]; bg.bxy = ob.bxy[F];
        for (int d = 0; d < 3; ++d) bg.lo[d] = ob.lo[F][d];
        const int RP = std::min(a.nRays, ctx->p.refillLanes > 0 ? std::min(ctx->p.refillLanes, 2048) : 256);      // (ray ids are 16 bits: RP * 32 <= 65536)
        const int nb = (int)std::min<long long>(a.nTiles, (long long)ctx->smCount * (ctx->p.maxBlocksPerSM > 0 ? ctx->p.maxBlocksPerSM : RMCRT_SORTED_BLOCKS));
        if (ctx->sortScratch.ensure((size_t)nb * sorted_scratch_bytes(RP, false) + 256)) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc sort scratch");
        unsigned char* base = static_cast<unsigned char*>(ctx->sortScratch.p);
        int* tileCounter = reinterpret_cast<int*>(base);
        CK(cudaMemsetAsync(tileCounter, 0, sizeof(int), st));
        divq_sorted_kernel<RecT, true, false><<<nb, 256, sorted_cell_smem(1), st>>>(a, bg, ob, RP, 1, base + 256, tileCounter);
      }
    } else if (ML && std::is_same<Real, float>::value && !ctx->p.genericMarch && a.G.maxLevels == 2 && patchInRoi) {
      // 2-level Data Onion fast path: per-patch ROI copy with a marker shell, then the tuned leg loop on both levels
      const LevelState& fl = ctx->lv[1];
      Box roiBox, shell;
      long long cellsB = 1;
      int cbLo[3], cbHi[3];
      onion_copy_box(ctx, a, 1, cbLo, cbHi);      // the ROI plus the level's extra cells where it ends at the domain boundary
      for (int d = 0; d < 3; ++d) {
        roiBox.lo[d] = cbLo[d]; roiBox.n[d] = cbHi[d] - cbLo[d];
        shell.lo[d] = roiBox.lo[d] - 1; shell.n[d] = roiBox.n[d] + 2;
        cellsB *= shell.n[d];
      }
      const long long guard = (long long)shell.n[0] * shell.n[1] + shell.n[0] + 1;
      if (cellsB + 2 * guard >= (1LL << 31)) return fail(ctx, RMCRT_ERR_INVALID, "ROI too large");
      if (ctx->roiBuf.ensure((size_t)(cellsB + 2 * guard) * sizeof(RecT))) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc ROI buffer");
      roi_copy_kernel<RecT><<<grid_for(cellsB + 2 * guard, 256, ctx->smCount), 256, 0, st>>>(
          static_cast<const RecT*>(fl.rec), mkbox(fl.d.lo, fl.d.hi), static_cast<RecT*>(ctx->roiBuf.p), guard, shell, roiBox);
      CK(cudaGetLastError());
      a.roi.rec = static_cast<RecT*>(ctx->roiBuf.p) + guard;
      for (int d = 0; d < 3; ++d) { a.roi.lo[d] = shell.lo[d]; a.roi.n[d] = shell.n[d]; }
      if (f32gen && ctx->p.rayCoherence >= 1 && ctx->p.cellsPerLane == 4) {
        // four cells per lane inside the ROI (divq_quad_kernel<ONION>); with reflective walls the one-cell kernel (launch_gate)
        if constexpr (std::is_same<Real, float>::value && !std::is_same<RecT, Rec16>::value) {
          const int tx = ctx->p.tileX == 8 || ctx->p.tileX == 16 || ctx->p.tileX == 32 ? ctx->p.tileX : RMCRT_QUAD_TILE_X;
          const int lxBits = tx == 8 ? 3 : (tx == 16 ? 4 : 5), ry = 2 * (32 / tx);
          const int nxc = (a.patchN[0] + tx - 1) / tx, njp = (a.patchN[1] + ry - 1) / ry, nkp = (a.patchN[2] + 1) / 2;
          long long g4 = (long long)nxc * njp * nkp;
          if (g4 > cap) g4 = cap;
          const int W4 = ctx->p.warpsPerTile > 0 ? std::min(ctx->p.warpsPerTile, 4) : 4;
          TraceArgs a2 = a;
          a2.gate = 1;
          divq_quad_kernel<RecT, true><<<(int)g4, 32 * W4, W4 * 128 * sizeof(double), st>>>(a2, nxc, njp, nkp, lxBits);
          a2.gate = 2;
          divq_kernel<Real, RecT, true, true, true><<<(int)grid, 32 * W, W * 32 * sizeof(double), st>>>(a2);
        }
      } else if (f32gen) divq_kernel<Real, RecT, true, true, true><<<(int)grid, 32 * W, W * 32 * sizeof(double), st>>>(a);
      else        divq_kernel<Real, RecT, true, true, false><<<(int)grid, 32 * W, W * 32 * sizeof(double), st>>>(a);
    } else if (ML) {
      if (f32gen) divq_kernel<Real, RecT, true, false, true><<<(int)grid, 32 * W, W * 32 * sizeof(double), st>>>(a);
      else        divq_kernel<Real, RecT, true, false, false><<<(int)grid, 32 * W, W * 32 * sizeof(double), st>>>(a);
    } else if (std::is_same<Real, float>::value && !bounded && f32gen && ctx->p.rayCoherence >= 1 && ctx->p.cellsPerLane == 4 && !ctx->p.genericMarch) {
      if constexpr (std::is_same<Real, float>::value && !std::is_same<RecT, Rec16>::value) {
        const int tx = ctx->p.tileX == 8 || ctx->p.tileX == 16 || ctx->p.tileX == 32 ? ctx->p.tileX : RMCRT_QUAD_TILE_X;
        const int lxBits = tx == 8 ? 3 : (tx == 16 ? 4 : 5), ry = 2 * (32 / tx);
        const int nxc = (a.patchN[0] + tx - 1) / tx, njp = (a.patchN[1] + ry - 1) / ry, nkp = (a.patchN[2] + 1) / 2;
        long long g4 = (long long)nxc * njp * nkp;
        if (g4 > cap) g4 = cap;
        int W4 = ctx->p.warpsPerTile > 0 ? std::min(ctx->p.warpsPerTile, 4) : 4;
        TraceArgs a2 = a;
        a2.gate = 1;
        divq_quad_kernel<RecT, false><<<(int)g4, 32 * W4, W4 * 128 * sizeof(double), st>>>(a2, nxc, njp, nkp, lxBits);
        launch_reflective_pairs<RecT>(ctx, a2, cap, st);
      }
    } else if (std::is_same<Real, float>::value && !bounded && f32gen && ctx->p.rayCoherence >= 1 && ctx->p.cellsPerLane == 2 && !ctx->p.genericMarch) {
      if constexpr (std::is_same<Real, float>::value && !std::is_same<RecT, Rec16>::value) {
        const int nxc = (a.patchN[0] + 31) / 32, njp = (a.patchN[1] + 1) / 2;
        long long g2 = (long long)nxc * njp * a.patchN[2];
        if (g2 > cap) g2 = cap;
        int W2 = ctx->p.warpsPerTile > 0 ? std::min(ctx->p.warpsPerTile, 8) : 4;
        if (W2 < 2) W2 = 2;        // rows A and B are written by warps 0 and 1
        TraceArgs a2 = a;
        a2.gate = 1;
        divq_dual_kernel<RecT, false><<<(int)g2, 32 * W2, W2 * 64 * sizeof(double), st>>>(a2, nxc, njp, 0);
        launch_reflective_pairs<RecT>(ctx, a2, cap, st);
      }
    } else if (ctx->p.blockedMarch == 3 && blocked_default(ctx, a, ML) && a.nRays >= RMCRT_SORT_MIN_RAYS) {
      // independent rays per (cell, ray, timestep), sorted by expected length inside a tile, on the blocked copy (divq_sorted_kernel)
      if constexpr (std::is_same<Real, float>::value && !std::is_same<RecT, Rec16>::value) {
        BlkGeom bg;
        int rc = ensure_blocked_level<RecT>(ctx, a.G.maxLevels - 1, bg, st);
        if (rc) return rc;
        compact_tiles(a);
        // few rays per cell: G tiles of 32 cells are sorted and marched together, so that a block still has dozens of 32-ray chunks for
        // its eight warps.  Measured at 128^3 (profiles/r02v40_few_rays.txt): it pays up to 16 rays per cell (3 / 5 / 10 rays: 4.1 / 4.3 /
        // 4.5e11 steps/s against 1.3 / 2.1 / 3.6e11 with G = 1 and 3.2 / 3.5 / 3.6e11 unsorted); from 25 rays on one tile sorts better alone
        int G = 1;
        if (a.nRays <= 16) while (G < 8 && 2 * G * a.nRays <= 96) G *= 2;
        // (rays, tile) pairs per lane and pass (A/B: refillLanes = rays of a cell per pass)
        const int RP = G * std::min(a.nRays, ctx->p.refillLanes > 0 ? std::min(ctx->p.refillLanes, 2048 / G) : 256 / G);      // (16-bit ray ids)
        // persistent blocks (their scratch is per block), tile groups handed out through a counter
        const int nb = (int)std::min<long long>((a.nTiles + G - 1) / G,
                                                (long long)ctx->smCount * (ctx->p.maxBlocksPerSM > 0 ? ctx->p.maxBlocksPerSM : RMCRT_SORTED_BLOCKS));
        const bool mayGoOn = a.allowReflect != 0;      // ... whether rays DO go on after a wall is known on the device only (level status)
        if (ctx->sortScratch.ensure((size_t)nb * sorted_scratch_bytes(RP, mayGoOn) + 256)) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc sort scratch");
        unsigned char* base = static_cast<unsigned char*>(ctx->sortScratch.p);
        int* tileCounter = reinterpret_cast<int*>(base);
        TraceArgs a2 = a;
        a2.gate = mayGoOn ? 1 : 0;
        CK(cudaMemsetAsync(tileCounter, 0, sizeof(int), st));
        OnionBlk none;
        memset(&none, 0, sizeof none);
        if (G > 1) divq_sorted_kernel<RecT, false, false, true><<<nb, 256, sorted_cell_smem(G), st>>>(a2, bg, none, RP, G, base + 256, tileCounter);
        else       divq_sorted_kernel<RecT, false, false, false><<<nb, 256, sorted_cell_smem(1), st>>>(a2, bg, none, RP, 1, base + 256, tileCounter);
        if (mayGoOn) {
          a2.gate = 2;
          CK(cudaMemsetAsync(tileCounter, 0, sizeof(int), st));
          if (G > 1) divq_sorted_kernel<RecT, false, true, true><<<nb, 256, sorted_cell_smem(G), st>>>(a2, bg, none, RP, G, base + 256, tileCounter);
          else       divq_sorted_kernel<RecT, false, true, false><<<nb, 256, sorted_cell_smem(1), st>>>(a2, bg, none, RP, 1, base + 256, tileCounter);
        }
      }
    } else if ((ctx->p.blockedMarch == 2 || ctx->p.blockedMarch == 3) && blocked_default(ctx, a, ML)) {
      // (also the default's choice below RMCRT_SORT_MIN_RAYS rays per cell)
      // independent rays per (cell, ray, timestep) on the blocked half-packed copy of the level, plain loads (divq_blk_kernel)
      if constexpr (std::is_same<Real, float>::value && !std::is_same<RecT, Rec16>::value) {
        BlkGeom bg;
        int rc = ensure_blocked_level<RecT>(ctx, a.G.maxLevels - 1, bg, st);
        if (rc) return rc;
        divq_blk_kernel<RecT><<<(int)grid, 32 * W, W * 32 * sizeof(double), st>>>(a, bg);
      }
    } else if (ctx->p.blockedMarch == 1 && blocked_default(ctx, a, ML)) {
      // opt-in: independent rays per (cell, ray, timestep) on the blocked half-packed copy of the level (indep.cuh)
      if constexpr (std::is_same<Real, float>::value && !std::is_same<RecT, Rec16>::value) {
        BlkGeom bg;
        int rc = ensure_blocked_level<RecT>(ctx, a.G.maxLevels - 1, bg, st);
        if (rc) return rc;
        const int nxc = (a.patchN[0] + 31) / 32, njp = (a.patchN[1] + 1) / 2, nkp = (a.patchN[2] + 1) / 2;
        const long long gi = (long long)nxc * njp * nkp;
        if (gi >= (1LL << 31)) return fail(ctx, RMCRT_ERR_INVALID, "patch too large");
        const int refill = ctx->p.refillLanes > 0 ? std::min(ctx->p.refillLanes, 32) : RMCRT_INDEP_REFILL;
        divq_indep_kernel<RecT><<<(int)gi, 128, 0, st>>>(a, bg, nxc, njp, nkp, refill);
      }
    } else if (std::is_same<Real, float>::value && !bounded && (!ctx->p.genericMarch || std::is_same<RecT, Rec4>::value)) {
      if (f32gen) divq_kernel<Real, RecT, false, true, true><<<(int)grid, 32 * W, W * 32 * sizeof(double), st>>>(a);
      else        divq_kernel<Real, RecT, false, true, false><<<(int)grid, 32 * W, W * 32 * sizeof(double), st>>>(a);
    } else if (std::is_same<Real, float>::value && std::is_same<RecT, Rec8>::value && !bounded && f32gen && ctx->p.rayCoherence == 0 &&
               ctx->p.originBrickHalo > 0 && ctx->p.originBrickHalo <= 16) {
      // A/B: the generic march with the origin tile's neighbourhood staged in shared memory (divq_brick_kernel)
      if constexpr (std::is_same<Real, float>::value && std::is_same<RecT, Rec8>::value) {
        const int h = ctx->p.originBrickHalo;
        // an A/B kernel: tiles must not wrap over patch rows, so that a tile's brick is (32 + 2h) x (2h + 1) x (2h + 1) at most
        if (a.patchN[0] % 32 != 0) return fail(ctx, RMCRT_ERR_INVALID, "originBrickHalo needs patch rows that are a multiple of 32 cells");
        const long long bx = std::min<long long>(32 + 2 * h, a.G.lev[a.G.maxLevels - 1].n[0]);
        const long long by = 2 * h + 1, bz = 2 * h + 1;
        const size_t smem = (size_t)W * 32 * sizeof(double) + (size_t)(bx * by * bz) * sizeof(Rec8);
        if (smem > 200 * 1024) return fail(ctx, RMCRT_ERR_INVALID, "originBrickHalo %d needs %zu bytes of shared memory", h, smem);
        CK(cudaFuncSetAttribute(divq_brick_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        divq_brick_kernel<<<(int)grid, 32 * W, smem, st>>>(a, h);
      }
    } else {
      if (f32gen) divq_kernel<Real, RecT, false, false, true><<<(int)grid, 32 * W, W * 32 * sizeof(double), st>>>(a);
      else        divq_kernel<Real, RecT, false, false, false><<<(int)grid, 32 * W, W * 32 * sizeof(double), st>>>(a);
    }
    CK(cudaGetLastError());
  }
  return 0;
}

// zero the cells of the output window that lie outside the traced patch (the trace kernels write every cell inside it)
__global__ void zero_border_kernel(double* __restrict__ q0, double* __restrict__ q1, int ox, int oy, int oz, int lx, int ly, int lz,
                                   int nx, int ny, int nz_) {
  const long long n = (long long)ox * oy * oz;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t % ox);
    const long long r = t / ox;
    const int y = (int)(r % oy), z = (int)(r / oy);
    const bool inside = x >= lx && x < lx + nx && y >= ly && y < ly + ny && z >= lz && z < lz + nz_;
    if (!inside) { q0[t] = 0.0; q1[t] = 0.0; }
  }
}

int trace_common(rmcrt_ctx* ctx, int maxLevels, bool ML, const int32_t pLo[3], const int32_t pHi[3], double* dDivQ, double* dVolq,
                 double* dFlux, const int32_t oLo[3], const int32_t oHi[3], int zeroInit, uint64_t timestep, cudaStream_t st) {
  TraceArgs a;
  int rc = build_args(ctx, maxLevels, ML, pLo, pHi, oLo, oHi, timestep, a);
  if (rc) return rc;
  if (!dDivQ || !dVolq) return fail(ctx, RMCRT_ERR_INVALID, "divQ and radiationVolq buffers are required");
  a.divQ = dDivQ; a.volq = dVolq; a.bflux = dFlux;
  a.zeroNonFlow = zeroInit ? 1 : 0;
  const size_t nOut = (size_t)a.outN[0] * a.outN[1] * a.outN[2];
  CK(cudaSetDevice(ctx->device));
  if (zeroInit == 2 && ctx->p.solveDivQ) {
    // outputs written straight into mapped host memory (trace_host): every cell of the patch is written by the divQ kernel, so only the
    // rest of the window is zeroed -- nothing crosses the bus twice
    zero_border_kernel<<<grid_for((long long)nOut, 256, ctx->smCount), 256, 0, st>>>(dDivQ, dVolq, a.outN[0], a.outN[1], a.outN[2],
        a.patchLo[0] - a.outLo[0], a.patchLo[1] - a.outLo[1], a.patchLo[2] - a.outLo[2], a.patchN[0], a.patchN[1], a.patchN[2]);
    CK(cudaGetLastError());
    if (dFlux) CK(cudaMemsetAsync(dFlux, 0, nOut * 7 * sizeof(double), st));
  } else if (zeroInit) {   // Ray.cc:577-587
    CK(cudaMemsetAsync(dDivQ, 0, nOut * sizeof(double), st));
    CK(cudaMemsetAsync(dVolq, 0, nOut * sizeof(double), st));
    if (dFlux) CK(cudaMemsetAsync(dFlux, 0, nOut * 7 * sizeof(double), st));
  }
  if (!blocked_default(ctx, a, ML) && !onion_blocked_default(ctx, a, ML)) {
    const LevelState& ls = ctx->lv[maxLevels - 1];
    persist_in_l2(ctx, ls.rec, (size_t)ls.nCells * ls.recBytes, st);
  }
  if (ctx->p.marchPrecision == 1) return launch_trace<double, Rec16>(ctx, a, ML, st);
  if (ctx->lv[maxLevels - 1].recordKind == 2) return launch_trace<float, Rec4>(ctx, a, ML, st);
  return launch_trace<float, Rec8>(ctx, a, ML, st);
}

int trace_host(rmcrt_ctx* ctx, int maxLevels, bool ML, const int32_t pLo[3], const int32_t pHi[3], double* divQ, double* volq,
               double* flux, const int32_t oLo[3], const int32_t oHi[3], int zeroInit, uint64_t timestep, cudaStream_t st) {
  if (!ctx) return RMCRT_ERR_INVALID;
  if (!divQ || !volq) return fail(ctx, RMCRT_ERR_INVALID, "divQ and radiationVolq buffers are required");
  size_t nOut = 1;
  for (int d = 0; d < 3; ++d) {
    if (oHi[d] <= oLo[d]) return fail(ctx, RMCRT_ERR_INVALID, "empty output window");
    nOut *= (size_t)(oHi[d] - oLo[d]);
  }
  CK(cudaSetDevice(ctx->device));
  double* dFlux = nullptr;
  // Pinned (page-locked, device-mapped) divQ / radiationVolq buffers are written by the kernels directly: the two f64 rows a tile
  // produces go over the bus while the other tiles are still marching, instead of a device buffer + a D2H copy after the solve.
  // Pageable buffers (a CCVariable of the stand-in DataWarehouse) take the staged path below.
  if (zeroInit && ctx->p.solveDivQ) {
    cudaPointerAttributes aq, av;
    const bool mq = cudaPointerGetAttributes(&aq, divQ) == cudaSuccess && aq.type == cudaMemoryTypeHost && aq.devicePointer;
    const bool mv = cudaPointerGetAttributes(&av, volq) == cudaSuccess && av.type == cudaMemoryTypeHost && av.devicePointer;
    cudaGetLastError();      // an unregistered host pointer is not an error here
    if (mq && mv) {
      if (flux) {
        if (ctx->outFlux.ensure(nOut * 56)) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc boundFlux");
        dFlux = static_cast<double*>(ctx->outFlux.p);
      }
      int rc = trace_common(ctx, maxLevels, ML, pLo, pHi, static_cast<double*>(aq.devicePointer), static_cast<double*>(av.devicePointer),
                            dFlux, oLo, oHi, 2, timestep, st);
      if (rc) return rc;
      if (flux) CK(cudaMemcpyAsync(flux, dFlux, nOut * 56, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      return 0;
    }
  }
  if (ctx->outDivQ.ensure(nOut * 8) || ctx->outVolq.ensure(nOut * 8)) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc outputs");
  if (flux) {
    if (ctx->outFlux.ensure(nOut * 56)) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc boundFlux");
    dFlux = static_cast<double*>(ctx->outFlux.p);
  }
  if (!zeroInit) {   // modifies_divQ: start from the caller's values
    CK(cudaMemcpyAsync(ctx->outDivQ.p, divQ, nOut * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->outVolq.p, volq, nOut * 8, cudaMemcpyHostToDevice, st));
    if (flux) CK(cudaMemcpyAsync(dFlux, flux, nOut * 56, cudaMemcpyHostToDevice, st));
  }
  int rc = trace_common(ctx, maxLevels, ML, pLo, pHi, static_cast<double*>(ctx->outDivQ.p), static_cast<double*>(ctx->outVolq.p),
                        dFlux, oLo, oHi, zeroInit, timestep, st);
  if (rc) return rc;
  CK(cudaMemcpyAsync(divQ, ctx->outDivQ.p, nOut * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(volq, ctx->outVolq.p, nOut * 8, cudaMemcpyDeviceToHost, st));
  if (flux) CK(cudaMemcpyAsync(flux, dFlux, nOut * 56, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

template <class Real, class RecT>
int launch_radiometer(rmcrt_ctx* ctx, const TraceArgs& a, const RadArgs& ra, cudaStream_t st) {
  const long long nCells = (long long)ra.n[0] * ra.n[1] * ra.n[2];
  const int grid = (int)std::min<long long>(nCells, (long long)ctx->smCount * 8);
  radiometer_kernel<Real, RecT><<<grid, 256, 0, st>>>(a, ra);
  CK(cudaGetLastError());
  return 0;
}

int radiometer_common(rmcrt_ctx* ctx, int L, const rmcrt_radiometer* rads, int nRads, const int32_t pLo[3], const int32_t pHi[3],
                      void* dFlux, void* dInt, int isFloat, const int32_t oLo[3], const int32_t oHi[3], int zeroInit,
                      uint64_t timestep, cudaStream_t st) {
  if (!rads && nRads > 0) return fail(ctx, RMCRT_ERR_INVALID, "null radiometer array");
  if (!dFlux || !dInt) return fail(ctx, RMCRT_ERR_INVALID, "VRFlux and VRIntensity buffers are required");
  TraceArgs a;
  int rc = build_args(ctx, L + 1, false, pLo, pHi, oLo, oHi, timestep, a);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  const size_t nOut = (size_t)a.outN[0] * a.outN[1] * a.outN[2];
  const size_t esz = isFloat ? 4 : 8;
  if (zeroInit) {     // Radiometer.cc:443-446
    CK(cudaMemsetAsync(dFlux, 0, nOut * esz, st));
    CK(cudaMemsetAsync(dInt, 0, nOut * esz, st));
  }
  for (int r = 0; r < nRads; ++r) {
    RadArgs ra;
    memset(&ra, 0, sizeof ra);
    ra.rad = rads[r];
    if (ra.rad.nRays < 1) return fail(ctx, RMCRT_ERR_INVALID, "radiometer %d: nRays must be >= 1", r);
    bool hit = true;
    for (int d = 0; d < 3; ++d) {     // doesIntersect + Max/Min (Radiometer.cc:451-454)
      const int lo = std::max(pLo[d], ra.rad.cellLo[d]), hi = std::min(pHi[d], ra.rad.cellHi[d]);
      ra.lo[d] = lo; ra.n[d] = hi - lo;
      hit = hit && hi > lo;
    }
    if (!hit) continue;
    ra.flux = dFlux; ra.intensity = dInt; ra.isFloat = isFloat;
    if (ctx->p.marchPrecision == 1) rc = launch_radiometer<double, Rec16>(ctx, a, ra, st);
    else if (ctx->lv[L].recordKind == 2) rc = launch_radiometer<float, Rec4>(ctx, a, ra, st);
    else rc = launch_radiometer<float, Rec8>(ctx, a, ra, st);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace

//======================================================================================================
// C ABI
//======================================================================================================
extern "C" {

int rmcrt_abi_version(void) { return RMCRT_ABI_VERSION; }

size_t rmcrt_struct_size(int which) {
  switch (which) {
    case 0: return sizeof(rmcrt_params);
    case 1: return sizeof(rmcrt_level_desc);
    case 2: return sizeof(rmcrt_counters);
    case 3: return sizeof(rmcrt_radiometer);
    default: return 0;
  }
}

void rmcrt_default_params(rmcrt_params* p) {
  memset(p, 0, sizeof *p);
  p->nDivQRays = 10; p->nFluxRays = 1; p->threshold = 0.01; p->sigma = 5.67051e-8;   // Ray.cc:162-173
  p->allowReflect = 1; p->CCRays = 0; p->solveDivQ = 1; p->solveBoundaryFlux = 0; p->isSeedRandom = 1;
  p->flowCell = -1;
  p->haloCells[0] = p->haloCells[1] = p->haloCells[2] = 10;   // Ray.cc:254
  p->cellTypeCoarsenLogic = 0;
  p->seed = 0x9E3779B97F4A7C15ull;
  p->marchPrecision = 0; p->rayCoherence = 0; p->skipZeroWeightLegs = 1; p->fixedRandom = 0; p->l2Persist = 1;
  p->warpsPerTile = 0;
  p->rayDirSampleAlgo = 0;     // "naive", Ray.cc:174
  p->maxRayLength = 0.0;       // unbounded (DBL_MAX, RMCRTCommon.cc:117)
  p->sigmaScat = 0.0;          // Ray.cc:169
  p->tileX = 0; p->nonFlowCellType = 8;
  p->blockedMarch = 3; p->refillLanes = 0;      // production: divq_sorted_kernel on the blocked half-packed copy
  p->coarsenExtraCells = 0; p->roiFixed = 0; p->originBrickHalo = 0; p->reserved2 = 0;
  for (int d = 0; d < 3; ++d) { p->roiFixedLo[d] = 0; p->roiFixedHi[d] = 0; }
  p->cellsPerLane = 4;         // production: the 2 x 2 cells of a lane share one DDA (divq_quad_kernel); 2 = a pair (divq_dual_kernel); 1 = one cell per lane
}

const char* rmcrt_last_error(const rmcrt_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  std::lock_guard<std::mutex> l(g_mu);
  return g_createErr.c_str();
}

int rmcrt_create(int device, const rmcrt_params* p, rmcrt_ctx** out) {
  rmcrt_ctx* ctx = nullptr;
  if (!out || !p) return fail(nullptr, RMCRT_ERR_INVALID, "null argument");
  *out = nullptr;
  int nDev = 0;
  cudaError_t e = cudaGetDeviceCount(&nDev);
  if (e != cudaSuccess || nDev <= 0)
    return fail(nullptr, RMCRT_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= nDev) return fail(nullptr, RMCRT_ERR_INVALID, "device %d out of range [0,%d)", device, nDev);
  int rc = validate_params(nullptr, p);
  if (rc) return rc;
  ctx = new rmcrt_ctx();
  ctx->device = device;
  ctx->p = *p;
  auto bail = [&](const char* what) { std::string m = what; rmcrt_destroy(ctx); return fail(nullptr, RMCRT_ERR_CUDA, "%s", m.c_str()); };
  if (cudaSetDevice(device) != cudaSuccess) return bail("cudaSetDevice failed");
  cudaDeviceGetAttribute(&ctx->smCount, cudaDevAttrMultiProcessorCount, device);
  if (cudaMalloc(&ctx->d_counters, NCOUNTERS * sizeof(unsigned long long)) != cudaSuccess) return bail("cudaMalloc counters");
  if (cudaMalloc(&ctx->d_lstat, RMCRT_MAX_LEVELS * LSTAT_WORDS * sizeof(unsigned)) != cudaSuccess) return bail("cudaMalloc status");
  if (cudaMalloc(&ctx->d_nItems, sizeof(int)) != cudaSuccess) return bail("cudaMalloc nItems");
  cudaMemset(ctx->d_counters, 0, NCOUNTERS * sizeof(unsigned long long));
  cudaMemset(ctx->d_lstat, 0, RMCRT_MAX_LEVELS * LSTAT_WORDS * sizeof(unsigned));
  cudaMemset(ctx->d_nItems, 0, sizeof(int));
  *out = ctx;
  return 0;
}

int rmcrt_set_params(rmcrt_ctx* ctx, const rmcrt_params* p) {
  if (!ctx || !p) return RMCRT_ERR_INVALID;
  int rc = validate_params(ctx, p);
  if (rc) return rc;
  ctx->p = *p;
  ctx->persistedStreamFor = nullptr;
  return 0;
}

void rmcrt_destroy(rmcrt_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (auto& l : ctx->lv) { if (l.base) cudaFree(l.base); if (l.recFull) cudaFree(l.recFull); if (l.blk) cudaFree(l.blk); }
  if (ctx->d_counters) cudaFree(ctx->d_counters);
  if (ctx->d_lstat) cudaFree(ctx->d_lstat);
  if (ctx->d_nItems) cudaFree(ctx->d_nItems);
  for (auto& b : ctx->stage) b.release();
  ctx->outDivQ.release(); ctx->outVolq.release(); ctx->outFlux.release();
  ctx->fluxFlags.release(); ctx->fluxItems.release(); ctx->compactTmp.release(); ctx->misc.release(); ctx->roiBuf.release();
  ctx->lhcDivQ.release(); ctx->lhcFlux.release(); ctx->outVR[0].release(); ctx->outVR[1].release();
  for (auto& b : ctx->dbg) b.release();
  ctx->replSend.release(); ctx->replGathered.release(); ctx->sortScratch.release();
  for (auto& b : ctx->regionBuf) b.release();
  delete ctx;
}

int rmcrt_define_level(rmcrt_ctx* ctx, int L, const rmcrt_level_desc* d, int recordKind) {
  if (!ctx || !d) return RMCRT_ERR_INVALID;
  if (L < 0 || L >= RMCRT_MAX_LEVELS) return fail(ctx, RMCRT_ERR_INVALID, "level index %d out of range", L);
  if (recordKind < 0 || recordKind > 2) return fail(ctx, RMCRT_ERR_INVALID, "recordKind must be 0 (float), 1 (double) or 2 (half-packed)");
  long long n = 1;
  for (int k = 0; k < 3; ++k) {
    if (d->hi[k] - d->lo[k] < 3) return fail(ctx, RMCRT_ERR_INVALID, "level needs >= 3 cells per axis incl. extra cells");
    if (d->interiorLo[k] <= d->lo[k] || d->interiorHi[k] >= d->hi[k] || d->interiorHi[k] <= d->interiorLo[k])
      return fail(ctx, RMCRT_ERR_INVALID, "level interior must be strictly inside [lo,hi): extraCells >= 1 (RMCRT_Test.cc:238-251)");
    if (!(d->dx[k] > 0)) return fail(ctx, RMCRT_ERR_INVALID, "dx must be positive");
    n *= (d->hi[k] - d->lo[k]);
  }
  if (n >= (1LL << 31)) return fail(ctx, RMCRT_ERR_INVALID, "level too large for 32-bit record strides");
  CK(cudaSetDevice(ctx->device));
  LevelState& s = ctx->lv[L];
  if (s.defined && s.base && s.recordKind == recordKind && s.nCells == n && memcmp(s.d.lo, d->lo, sizeof d->lo) == 0 &&
      memcmp(s.d.hi, d->hi, sizeof d->hi) == 0) {
    s.d = *d;              // same shape: keep the allocation (a new radiation solve re-packs into it)
    s.finalized = false;
    s.blkValid = false;
    CK(cudaMemset(lstat_of(ctx, L), 0, LSTAT_WORDS * sizeof(unsigned)));
    return 0;
  }
  if (s.base) { cudaFree(s.base); s.base = nullptr; s.rec = nullptr; }
  if (s.recFull) { cudaFree(s.recFull); s.recFull = nullptr; }
  s.blkValid = false;
  s.d = *d;
  s.recordKind = recordKind;
  s.recBytes = recordKind == 0 ? sizeof(Rec8) : (recordKind == 1 ? sizeof(Rec16) : sizeof(Rec4));
  s.nCells = n;
  // guard records on both ends (flagged "not a flow cell") so neighbour / speculative loads issued from any cell of
  // the level stay inside the allocation (one speculative step past the cell whose wall test is pending; sized for two)
  const long long guard = 2 * ((long long)(d->hi[0] - d->lo[0]) * (d->hi[1] - d->lo[1]) + (d->hi[0] - d->lo[0]) + 1);
  void* base = nullptr;
  CK(cudaMalloc(&base, (size_t)(n + 2 * guard) * s.recBytes));
  CK(cudaMemset(base, 0xFF, (size_t)(n + 2 * guard) * s.recBytes));   // 0xFF.. = negative NaN: wall flag set, caught by finalize if left
  s.base = base;
  s.rec = static_cast<char*>(base) + (size_t)guard * s.recBytes;
  if (recordKind == 2) {
    CK(cudaMalloc(&s.recFull, (size_t)n * sizeof(Rec8)));
    CK(cudaMemset(s.recFull, 0xFF, (size_t)n * sizeof(Rec8)));
  }
  s.defined = true;
  s.finalized = false;
  CK(cudaMemset(lstat_of(ctx, L), 0, LSTAT_WORDS * sizeof(unsigned)));
  ctx->persistedStreamFor = nullptr;
  return 0;
}

int rmcrt_pack_level_device(rmcrt_ctx* ctx, int L, const void* a, const void* s, const int32_t* ct, int isFloat,
                            const int32_t srcLo[3], const int32_t srcHi[3], void* stream) {
  int rc = check_level(ctx, L);
  if (rc) return rc;
  if (!a || !s || !ct) return fail(ctx, RMCRT_ERR_INVALID, "null property pointer");
  CK(cudaSetDevice(ctx->device));
  LevelState& lv = ctx->lv[L];
  Box src = mkbox(srcLo, srcHi), dst = mkbox(lv.d.lo, lv.d.hi);
  Box region;
  for (int d = 0; d < 3; ++d) {
    const int lo = std::max(srcLo[d], lv.d.lo[d]), hi = std::min(srcHi[d], lv.d.hi[d]);
    if (hi <= lo) return fail(ctx, RMCRT_ERR_INVALID, "source window does not overlap level %d", L);
    region.lo[d] = lo; region.n[d] = hi - lo;
  }
  int rci = invalidate_level(ctx, L, (cudaStream_t)stream);
  if (rci) return rci;
  lv.propsFloat = isFloat != 0;
  return do_pack(ctx, lv.recordKind, isFloat, a, s, ct, src, region, lv.rec, dst, lstat_of(ctx, L), lv.recFull, dst, (cudaStream_t)stream);
}

static int wait_for_copies(rmcrt_ctx* ctx, cudaStream_t st) {
  cudaEvent_t ev;
  CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaError_t e = cudaEventRecord(ev, st);
  if (e == cudaSuccess) e = cudaEventSynchronize(ev);
  cudaEventDestroy(ev);
  if (e != cudaSuccess) return fail(ctx, RMCRT_ERR_CUDA, "waiting for the host-to-device copies failed: %s", cudaGetErrorString(e));
  return 0;
}

int rmcrt_upload_level(rmcrt_ctx* ctx, int L, const void* a, const void* s, const int32_t* ct, int isFloat,
                       const int32_t srcLo[3], const int32_t srcHi[3], void* stream) {
  int rc = check_level(ctx, L);
  if (rc) return rc;
  if (!a || !s || !ct) return fail(ctx, RMCRT_ERR_INVALID, "null property pointer");
  CK(cudaSetDevice(ctx->device));
  size_t n = 1;
  for (int d = 0; d < 3; ++d) {
    if (srcHi[d] <= srcLo[d]) return fail(ctx, RMCRT_ERR_INVALID, "empty source window");
    n *= (size_t)(srcHi[d] - srcLo[d]);
  }
  const size_t eb = isFloat ? 4 : 8;
  if (ctx->stage[0].ensure(n * eb) || ctx->stage[1].ensure(n * eb) || ctx->stage[2].ensure(n * 4))
    return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc staging failed");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(ctx->stage[0].p, a, n * eb, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->stage[1].p, s, n * eb, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->stage[2].p, ct, n * 4, cudaMemcpyHostToDevice, st));
  // The caller's arrays are not read after this call returns (the DataWarehouse frees or recycles getLevel temporaries at once): from
  // page-locked memory the copies above are truly asynchronous, so wait for THEM -- not for the pack kernel queued behind them
  rc = wait_for_copies(ctx, st);
  if (rc) return rc;
  return rmcrt_pack_level_device(ctx, L, ctx->stage[0].p, ctx->stage[1].p, static_cast<const int32_t*>(ctx->stage[2].p), isFloat,
                                 srcLo, srcHi, stream);
}

static int finalize_launch(rmcrt_ctx* ctx, int L, cudaStream_t st) {
  LevelState& lv = ctx->lv[L];
  unsigned* ls = lstat_of(ctx, L);
  CK(cudaMemsetAsync(ls, 0, 5 * sizeof(unsigned), st));      // words 0..4; word 5 (pack-time input check) survives
  CK(cudaMemsetAsync(ls + 6, 0, 2 * sizeof(unsigned), st));
  Box b = mkbox(lv.d.lo, lv.d.hi);
  if (lv.recordKind == 0)      finalize_kernel<Rec8><<<grid_for(lv.nCells, 256, ctx->smCount), 256, 0, st>>>(static_cast<const Rec8*>(lv.rec), b, ls);
  else if (lv.recordKind == 2) finalize_kernel<Rec4><<<grid_for(lv.nCells, 256, ctx->smCount), 256, 0, st>>>(static_cast<const Rec4*>(lv.rec), b, ls);
  else                         finalize_kernel<Rec16><<<grid_for(lv.nCells, 256, ctx->smCount), 256, 0, st>>>(static_cast<const Rec16*>(lv.rec), b, ls);
  CK(cudaGetLastError());
  return 0;
}

int rmcrt_finalize_level_async(rmcrt_ctx* ctx, int L, void* stream) {
  int rc = check_level(ctx, L);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  return finalize_launch(ctx, L, (cudaStream_t)stream);
}

int rmcrt_finalize_level(rmcrt_ctx* ctx, int L, void* stream) {
  int rc = check_level(ctx, L);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  LevelState& lv = ctx->lv[L];
  rc = finalize_launch(ctx, L, st);
  if (rc) return rc;
  unsigned h[LSTAT_WORDS];
  CK(cudaMemcpyAsync(h, lstat_of(ctx, L), sizeof h, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (h[5]) CK(cudaMemsetAsync(lstat_of(ctx, L) + 5, 0, sizeof(unsigned), st));
  if (h[0] || h[5])
    return fail(ctx, RMCRT_ERR_NONPHYSICAL, "level %d: %u non-finite abskg/sigmaT4 values (or cells never packed)", L, h[0] + h[5]);
  if (h[1]) return fail(ctx, RMCRT_ERR_OPEN_DOMAIN, "level %d: %u flow cells on the outer (extra-cell) layer; rays could leave the level", L, h[1]);
  lv.wallsEmit = (h[2] != 0);
  lv.finalized = true;
  return 0;
}

int rmcrt_level_records(rmcrt_ctx* ctx, int L, void** d_records, size_t* bytes, size_t* recordBytes) {
  int rc = check_level(ctx, L);
  if (rc) return rc;
  if (d_records) *d_records = ctx->lv[L].rec;
  if (bytes) *bytes = (size_t)ctx->lv[L].nCells * ctx->lv[L].recBytes;
  if (recordBytes) *recordBytes = ctx->lv[L].recBytes;
  return 0;
}

int rmcrt_trace_divq_device(rmcrt_ctx* ctx, int L, const int32_t pLo[3], const int32_t pHi[3], double* dq, double* vq, double* bf,
                            const int32_t oLo[3], const int32_t oHi[3], int zeroInit, uint64_t timestep, void* stream) {
  int rc = check_level(ctx, L);
  if (rc) return rc;
  if (L + 1 > MAXL) return fail(ctx, RMCRT_ERR_INVALID, "level index too large");
  return trace_common(ctx, L + 1, false, pLo, pHi, dq, vq, bf, oLo, oHi, zeroInit, timestep, (cudaStream_t)stream);
}

int rmcrt_trace_divq(rmcrt_ctx* ctx, int L, const int32_t pLo[3], const int32_t pHi[3], double* dq, double* vq, double* bf,
                     const int32_t oLo[3], const int32_t oHi[3], int zeroInit, uint64_t timestep, void* stream) {
  int rc = check_level(ctx, L);
  if (rc) return rc;
  return trace_host(ctx, L + 1, false, pLo, pHi, dq, vq, bf, oLo, oHi, zeroInit, timestep, (cudaStream_t)stream);
}

int rmcrt_trace_dataonion_device(rmcrt_ctx* ctx, int maxLevels, const int32_t pLo[3], const int32_t pHi[3], double* dq, double* vq,
                                 double* bf, const int32_t oLo[3], const int32_t oHi[3], int zeroInit, uint64_t timestep, void* stream) {
  if (!ctx) return RMCRT_ERR_INVALID;
  return trace_common(ctx, maxLevels, true, pLo, pHi, dq, vq, bf, oLo, oHi, zeroInit, timestep, (cudaStream_t)stream);
}

int rmcrt_trace_dataonion(rmcrt_ctx* ctx, int maxLevels, const int32_t pLo[3], const int32_t pHi[3], double* dq, double* vq,
                          double* bf, const int32_t oLo[3], const int32_t oHi[3], int zeroInit, uint64_t timestep, void* stream) {
  if (!ctx) return RMCRT_ERR_INVALID;
  return trace_host(ctx, maxLevels, true, pLo, pHi, dq, vq, bf, oLo, oHi, zeroInit, timestep, (cudaStream_t)stream);
}

// Radiometer::problemSetup (Radiometer.cc:118-120, 203-260)
int rmcrt_radiometer_setup(double viewAngle, const double vec[3], int isNormalVector, rmcrt_radiometer* r) {
  if (!vec || !r) return RMCRT_ERR_INVALID;
  if (viewAngle > 360) return fail(nullptr, RMCRT_ERR_INVALID, "VRViewAngle (%g) exceeds the maximum acceptable value of 360 degrees", viewAngle);
  double n[3] = {vec[0], vec[1], vec[2]};
  if (isNormalVector) {
    const double l = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    for (int d = 0; d < 3; ++d) n[d] = n[d] / l;
  }
  for (int d = 0; d < 3; ++d) if (n[d] == 0) n[d] = 1e-16;       // 216-220 (the reference's divide-by-zero guard)
  const double len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  r->phi_rotate = 0;                                              // 238: never a rotation about x
  r->theta_rotate = acos(n[2] / len);                             // 240
  double xi = acos(n[0] / sqrt(n[0] * n[0] + n[1] * n[1]));       // 242
  if (n[0] < 0 && n[1] < 0) xi = (M_PI / 2 + xi);                 // 246-248 (quadrant 3, as coded in the reference)
  if (n[0] > 0 && n[1] < 0) xi = (2 * M_PI - xi);                 // 249-251 (quadrant 4)
  r->xi_rotate = xi;
  r->theta_viewAngle = viewAngle / 360 * M_PI;                    // 258
  r->range = 1 - cos(r->theta_viewAngle);                         // 259
  r->solidAngle = 2 * M_PI * r->range;                            // 260
  return 0;
}

int rmcrt_trace_radiometer_device(rmcrt_ctx* ctx, int L, const rmcrt_radiometer* rads, int nRads, const int32_t pLo[3],
                                  const int32_t pHi[3], void* dFlux, void* dInt, int isFloat, const int32_t oLo[3],
                                  const int32_t oHi[3], int zeroInit, uint64_t timestep, void* stream) {
  if (!ctx) return RMCRT_ERR_INVALID;
  return radiometer_common(ctx, L, rads, nRads, pLo, pHi, dFlux, dInt, isFloat, oLo, oHi, zeroInit, timestep, (cudaStream_t)stream);
}

int rmcrt_trace_radiometer(rmcrt_ctx* ctx, int L, const rmcrt_radiometer* rads, int nRads, const int32_t pLo[3], const int32_t pHi[3],
                           void* flux, void* inten, int isFloat, const int32_t oLo[3], const int32_t oHi[3], int zeroInit,
                           uint64_t timestep, void* stream) {
  if (!ctx) return RMCRT_ERR_INVALID;
  if (!flux || !inten) return fail(ctx, RMCRT_ERR_INVALID, "VRFlux and VRIntensity buffers are required");
  cudaStream_t st = (cudaStream_t)stream;
  size_t nOut = 1;
  for (int d = 0; d < 3; ++d) {
    if (oHi[d] <= oLo[d]) return fail(ctx, RMCRT_ERR_INVALID, "empty output window");
    nOut *= (size_t)(oHi[d] - oLo[d]);
  }
  const size_t bytes = nOut * (isFloat ? 4 : 8);
  CK(cudaSetDevice(ctx->device));
  if (ctx->outVR[0].ensure(bytes) || ctx->outVR[1].ensure(bytes)) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc radiometer outputs");
  if (!zeroInit) {     // modifiesFlux: start from the caller's values
    CK(cudaMemcpyAsync(ctx->outVR[0].p, flux, bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->outVR[1].p, inten, bytes, cudaMemcpyHostToDevice, st));
  }
  int rc = radiometer_common(ctx, L, rads, nRads, pLo, pHi, ctx->outVR[0].p, ctx->outVR[1].p, isFloat, oLo, oHi, zeroInit, timestep, st);
  if (rc) return rc;
  CK(cudaMemcpyAsync(flux, ctx->outVR[0].p, bytes, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(inten, ctx->outVR[1].p, bytes, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int rmcrt_roi_extents(rmcrt_ctx* ctx, int L, const int32_t pLo[3], const int32_t pHi[3], double abskgThreshold, double sigmaT4Threshold,
                      void* magGradAbskg, void* magGradSigmaT4, int32_t* flaggedCells, int isFloat, const int32_t oLo[3], const int32_t oHi[3],
                      int32_t* anyFlagged, void* stream) {
  int rc = check_level(ctx, L);
  if (rc) return rc;
  if (!anyFlagged) return fail(ctx, RMCRT_ERR_INVALID, "anyFlagged is required");
  const bool outs = magGradAbskg || magGradSigmaT4 || flaggedCells;
  if (outs && !(magGradAbskg && magGradSigmaT4 && flaggedCells && oLo && oHi))
    return fail(ctx, RMCRT_ERR_INVALID, "mag_grad_abskg, mag_grad_sigmaT4 and flaggedCells are given together or not at all");
  const LevelState& lv = ctx->lv[L];
  if ((lv.recordKind == 1) == (isFloat != 0) && outs) return fail(ctx, RMCRT_ERR_INVALID, "isFloat must match the level's records (double records <=> double outputs)");
  for (int d = 0; d < 3; ++d)
    if (pLo[d] < lv.d.interiorLo[d] || pHi[d] > lv.d.interiorHi[d] || pHi[d] <= pLo[d]) return fail(ctx, RMCRT_ERR_INVALID, "patch outside the level interior");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  Box lb = mkbox(lv.d.lo, lv.d.hi), pb = mkbox(pLo, pHi), ob = outs ? mkbox(oLo, oHi) : pb;
  const size_t nOut = outs ? (size_t)box_cells(ob) : 0;
  const size_t eb = isFloat ? 4 : 8;
  if (ctx->misc.ensure(64)) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc");
  unsigned* dAny = static_cast<unsigned*>(ctx->misc.p);
  CK(cudaMemsetAsync(dAny, 0, sizeof(unsigned), st));
  void *dA = nullptr, *dS = nullptr; int* dF = nullptr;
  if (outs) {
    if (ctx->stage[0].ensure(nOut * eb) || ctx->stage[1].ensure(nOut * eb) || ctx->stage[2].ensure(nOut * 4)) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc staging failed");
    dA = ctx->stage[0].p; dS = ctx->stage[1].p; dF = static_cast<int*>(ctx->stage[2].p);
    CK(cudaMemsetAsync(dA, 0, nOut * eb, st)); CK(cudaMemsetAsync(dS, 0, nOut * eb, st)); CK(cudaMemsetAsync(dF, 0, nOut * 4, st));   // Ray.cc:2010-2012
  }
  const int grid = grid_for(box_cells(pb), 256, ctx->smCount);
  const double* dx = lv.d.dx;
  if (lv.recordKind == 1)
    roi_extents_kernel<Rec16, double><<<grid, 256, 0, st>>>(static_cast<const Rec16*>(lv.rec), lb, pb, dx[0], dx[1], dx[2], abskgThreshold, sigmaT4Threshold,
                                                           static_cast<double*>(dA), static_cast<double*>(dS), dF, ob, dAny);
  else if (lv.recordKind == 2)
    roi_extents_kernel<Rec4, float><<<grid, 256, 0, st>>>(static_cast<const Rec4*>(lv.rec), lb, pb, dx[0], dx[1], dx[2], abskgThreshold, sigmaT4Threshold,
                                                         static_cast<float*>(dA), static_cast<float*>(dS), dF, ob, dAny);
  else
    roi_extents_kernel<Rec8, float><<<grid, 256, 0, st>>>(static_cast<const Rec8*>(lv.rec), lb, pb, dx[0], dx[1], dx[2], abskgThreshold, sigmaT4Threshold,
                                                         static_cast<float*>(dA), static_cast<float*>(dS), dF, ob, dAny);
  CK(cudaGetLastError());
  unsigned h = 0;
  CK(cudaMemcpyAsync(&h, dAny, sizeof h, cudaMemcpyDeviceToHost, st));
  if (outs) {
    CK(cudaMemcpyAsync(magGradAbskg, dA, nOut * eb, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(magGradSigmaT4, dS, nOut * eb, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(flaggedCells, dF, nOut * 4, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  *anyFlagged = h ? 1 : 0;
  return 0;
}

int rmcrt_compute_extents(rmcrt_ctx* ctx, int maxLevels, const int32_t pLo[3], const int32_t pHi[3], int32_t roiLo[3],
                          int32_t roiHi[3], int32_t* regionLo, int32_t* regionHi) {
  if (!ctx) return RMCRT_ERR_INVALID;
  if (maxLevels < 1 || maxLevels > RMCRT_MAX_LEVELS) return fail(ctx, RMCRT_ERR_INVALID, "bad maxLevels");
  for (int L = 0; L < maxLevels; ++L) { int rc = check_level(ctx, L); if (rc) return rc; }
  compute_extents(ctx, maxLevels, pLo, pHi, roiLo, roiHi, regionLo, regionHi);
  return 0;
}

int rmcrt_reset_counters(rmcrt_ctx* ctx, void* stream) {
  if (!ctx) return RMCRT_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemsetAsync(ctx->d_counters, 0, NCOUNTERS * sizeof(unsigned long long), (cudaStream_t)stream));
  return 0;
}

int rmcrt_get_counters(rmcrt_ctx* ctx, rmcrt_counters* out, void* stream) {
  if (!ctx || !out) return RMCRT_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  unsigned long long h[NCOUNTERS];
  CK(cudaMemcpyAsync(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  out->nRaySteps = h[0]; out->nFluxRaySteps = h[1]; out->nDivQCells = h[2]; out->nFluxCells = h[3]; out->nRadiometerSteps = h[4];
  return 0;
}

int rmcrt_sigmaT4_device(rmcrt_ctx* ctx, const double* d_temp, void* d_out, int isFloat, int64_t n, void* stream) {
  if (!ctx || !d_temp || !d_out || n < 0) return RMCRT_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return 0;
  const double sop = ctx->p.sigma / M_PI;
  if (isFloat) sigmaT4_kernel<float><<<grid_for(n, 256, ctx->smCount), 256, 0, (cudaStream_t)stream>>>(d_temp, static_cast<float*>(d_out), n, sop);
  else         sigmaT4_kernel<double><<<grid_for(n, 256, ctx->smCount), 256, 0, (cudaStream_t)stream>>>(d_temp, static_cast<double*>(d_out), n, sop);
  CK(cudaGetLastError());
  return 0;
}

int rmcrt_sigmaT4(rmcrt_ctx* ctx, const double* temp, void* out, int isFloat, int64_t n, void* stream) {
  if (!ctx || !temp || !out || n < 0) return RMCRT_ERR_INVALID;
  if (n == 0) return 0;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t ob = (size_t)n * (isFloat ? 4 : 8);
  if (ctx->stage[0].ensure((size_t)n * 8) || ctx->stage[1].ensure(ob)) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc staging failed");
  CK(cudaMemcpyAsync(ctx->stage[0].p, temp, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  int rc = rmcrt_sigmaT4_device(ctx, static_cast<const double*>(ctx->stage[0].p), ctx->stage[1].p, isFloat, n, stream);
  if (rc) return rc;
  CK(cudaMemcpyAsync(out, ctx->stage[1].p, ob, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int rmcrt_coarsen_level_device(rmcrt_ctx* ctx, int fineL, int coarseL, void* stream) {
  int rc = check_level(ctx, fineL);
  if (rc) return rc;
  rc = check_level(ctx, coarseL);
  if (rc) return rc;
  LevelState& f = ctx->lv[fineL];
  LevelState& c = ctx->lv[coarseL];
  if (f.recordKind != c.recordKind) return fail(ctx, RMCRT_ERR_INVALID, "recordKind mismatch between levels");
  c.propsFloat = f.propsFloat;
  const int* rr = f.d.refineRatio;
  for (int d = 0; d < 3; ++d) {
    if (rr[d] < 1) return fail(ctx, RMCRT_ERR_INVALID, "fine level refineRatio must be >= 1");
    if ((c.d.interiorHi[d] - c.d.interiorLo[d]) * rr[d] != (f.d.interiorHi[d] - f.d.interiorLo[d]) ||
        c.d.interiorLo[d] * rr[d] != f.d.interiorLo[d])
      return fail(ctx, RMCRT_ERR_INVALID, "coarse interior x refineRatio must equal the fine interior on axis %d", d);
  }
  CK(cudaSetDevice(ctx->device));
  Box fb = mkbox(f.d.lo, f.d.hi), cb = mkbox(c.d.lo, c.d.hi), ci = mkbox(c.d.interiorLo, c.d.interiorHi);
  const long long n = box_cells(ci);
  cudaStream_t st = (cudaStream_t)stream;
  if (f.recordKind == 0)      coarsen_kernel<Rec8><<<grid_for(n, 128, ctx->smCount), 128, 0, st>>>(static_cast<const Rec8*>(f.rec), fb, static_cast<Rec8*>(c.rec), cb, ci, rr[0], rr[1], rr[2], ctx->p.cellTypeCoarsenLogic, ctx->p.nonFlowCellType, nullptr);
  else if (f.recordKind == 2) coarsen_kernel<Rec4><<<grid_for(n, 128, ctx->smCount), 128, 0, st>>>(static_cast<const Rec4*>(f.rec), fb, static_cast<Rec4*>(c.rec), cb, ci, rr[0], rr[1], rr[2], ctx->p.cellTypeCoarsenLogic, ctx->p.nonFlowCellType, static_cast<Rec8*>(c.recFull));
  else                        coarsen_kernel<Rec16><<<grid_for(n, 128, ctx->smCount), 128, 0, st>>>(static_cast<const Rec16*>(f.rec), fb, static_cast<Rec16*>(c.rec), cb, ci, rr[0], rr[1], rr[2], ctx->p.cellTypeCoarsenLogic, ctx->p.nonFlowCellType, nullptr);
  CK(cudaGetLastError());
  if (ctx->p.coarsenExtraCells) {
    for (int d = 0; d < 3; ++d)
      if (f.d.interiorLo[d] - f.d.lo[d] != 1 || f.d.hi[d] - f.d.interiorHi[d] != 1 || c.d.interiorLo[d] - c.d.lo[d] != 1 || c.d.hi[d] - c.d.interiorHi[d] != 1)
        return fail(ctx, RMCRT_ERR_INVALID, "coarsenExtraCells needs one layer of extra cells on both levels");
    Box fi = mkbox(f.d.interiorLo, f.d.interiorHi);
    const long long nc = box_cells(cb);
    if (f.recordKind == 0)      coarsen_extra_kernel<Rec8><<<grid_for(nc, 128, ctx->smCount), 128, 0, st>>>(static_cast<const Rec8*>(f.rec), fb, fi, static_cast<Rec8*>(c.rec), cb, ci, rr[0], rr[1], rr[2], nullptr);
    else if (f.recordKind == 2) coarsen_extra_kernel<Rec4><<<grid_for(nc, 128, ctx->smCount), 128, 0, st>>>(static_cast<const Rec4*>(f.rec), fb, fi, static_cast<Rec4*>(c.rec), cb, ci, rr[0], rr[1], rr[2], static_cast<Rec8*>(c.recFull));
    else                        coarsen_extra_kernel<Rec16><<<grid_for(nc, 128, ctx->smCount), 128, 0, st>>>(static_cast<const Rec16*>(f.rec), fb, fi, static_cast<Rec16*>(c.rec), cb, ci, rr[0], rr[1], rr[2], nullptr);
    CK(cudaGetLastError());
  }
  return invalidate_level(ctx, coarseL, st);
}

int rmcrt_set_level_extra_cells(rmcrt_ctx* ctx, int L, const double abskg[6], const double sig[6], const int32_t types[6], void* stream) {
  int rc = check_level(ctx, L);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  LevelState& lv = ctx->lv[L];
  if (ctx->misc.ensure(12 * sizeof(double) + 6 * sizeof(int))) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc");
  double h[12];
  for (int f = 0; f < 6; ++f) { h[f] = abskg[f]; h[6 + f] = sig[f]; }
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(ctx->misc.p, h, sizeof h, cudaMemcpyHostToDevice, st));
  int* dT = reinterpret_cast<int*>(static_cast<char*>(ctx->misc.p) + sizeof h);
  CK(cudaMemcpyAsync(dT, types, 6 * sizeof(int), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));   // h[] is a stack buffer
  Box b = mkbox(lv.d.lo, lv.d.hi), in = mkbox(lv.d.interiorLo, lv.d.interiorHi);
  if (lv.recordKind == 0)      extra_cells_kernel<Rec8><<<grid_for(lv.nCells, 256, ctx->smCount), 256, 0, st>>>(static_cast<Rec8*>(lv.rec), b, in, static_cast<const double*>(ctx->misc.p), dT, ctx->p.flowCell, nullptr);
  else if (lv.recordKind == 2) extra_cells_kernel<Rec4><<<grid_for(lv.nCells, 256, ctx->smCount), 256, 0, st>>>(static_cast<Rec4*>(lv.rec), b, in, static_cast<const double*>(ctx->misc.p), dT, ctx->p.flowCell, static_cast<Rec8*>(lv.recFull));
  else                         extra_cells_kernel<Rec16><<<grid_for(lv.nCells, 256, ctx->smCount), 256, 0, st>>>(static_cast<Rec16*>(lv.rec), b, in, static_cast<const double*>(ctx->misc.p), dT, ctx->p.flowCell, nullptr);
  CK(cudaGetLastError());
  return invalidate_level(ctx, L, st);
}

//------------------------------------------------------------------ callers either side of the path
static int refine_args(rmcrt_ctx* ctx, const int32_t cLo[3], const int32_t cHi[3], const int32_t gLo[3], const int32_t gHi[3],
                       const int32_t rr[3], int order, const int32_t fl[3], const int32_t fh[3], const int32_t oLo[3],
                       const int32_t oHi[3], RefineArgs& a) {
  if (order < 0 || order > 2)      // selectInterpolator's InternalError (AMR.h:735-737)
    return fail(ctx, RMCRT_ERR_INVALID, "ERROR:AMR: You're trying to use an interpolator that doesn't exist.  <orderOfInterpolation> must be 0,1,2");
  a.cw = mkbox(cLo, cHi); a.fr = mkbox(fl, fh); a.ow = mkbox(oLo, oHi);
  a.order = order;
  const int reach = order == 0 ? 0 : 1;
  for (int d = 0; d < 3; ++d) {
    if (rr[d] < 1) return fail(ctx, RMCRT_ERR_INVALID, "refinement ratio must be >= 1");
    if (fh[d] < fl[d] || oHi[d] <= oLo[d] || cHi[d] <= cLo[d]) return fail(ctx, RMCRT_ERR_INVALID, "empty window");
    if (fl[d] < oLo[d] || fh[d] > oHi[d]) return fail(ctx, RMCRT_ERR_INVALID, "fine region outside of the fine window on axis %d", d);
    a.rr[d] = rr[d]; a.gridLo[d] = gLo[d]; a.gridHiM1[d] = gHi[d] - 1;
    if (fh[d] > fl[d]) {
      const int c0 = map_coarser_host(fl[d], rr[d]) - reach, c1 = map_coarser_host(fh[d] - 1, rr[d]) + reach;
      if (c0 < cLo[d] || c1 >= cHi[d])
        return fail(ctx, RMCRT_ERR_INVALID, "coarse window [%d,%d) does not cover the cells [%d,%d] the interpolation reads on axis %d", cLo[d], cHi[d], c0, c1, d);
    }
  }
  return 0;
}

int rmcrt_refine_Q_device(rmcrt_ctx* ctx, const double* d_coarseQ, const int32_t cLo[3], const int32_t cHi[3], const int32_t gLo[3],
                          const int32_t gHi[3], const int32_t rr[3], int order, const int32_t fl[3], const int32_t fh[3],
                          double* d_fineQ, const int32_t oLo[3], const int32_t oHi[3], int zeroInit, void* stream) {
  if (!ctx) return RMCRT_ERR_INVALID;
  if (!d_coarseQ || !d_fineQ) return fail(ctx, RMCRT_ERR_INVALID, "null field pointer");
  RefineArgs a;
  int rc = refine_args(ctx, cLo, cHi, gLo, gHi, rr, order, fl, fh, oLo, oHi, a);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (zeroInit) CK(cudaMemsetAsync(d_fineQ, 0, (size_t)box_cells(a.ow) * sizeof(double), st));   // Ray.cc:1899-1900
  const long long n = box_cells(a.fr);
  if (n == 0) return 0;
  refine_kernel<<<grid_for(n, 256, ctx->smCount), 256, 0, st>>>(d_coarseQ, d_fineQ, a);
  CK(cudaGetLastError());
  return 0;
}

int rmcrt_refine_Q(rmcrt_ctx* ctx, const double* coarseQ, const int32_t cLo[3], const int32_t cHi[3], const int32_t gLo[3],
                   const int32_t gHi[3], const int32_t rr[3], int order, const int32_t fl[3], const int32_t fh[3], double* fineQ,
                   const int32_t oLo[3], const int32_t oHi[3], void* stream) {
  if (!ctx) return RMCRT_ERR_INVALID;
  if (!coarseQ || !fineQ) return fail(ctx, RMCRT_ERR_INVALID, "null field pointer");
  RefineArgs a;
  int rc = refine_args(ctx, cLo, cHi, gLo, gHi, rr, order, fl, fh, oLo, oHi, a);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t cb = (size_t)box_cells(a.cw) * sizeof(double), fb = (size_t)box_cells(a.ow) * sizeof(double);
  if (ctx->stage[0].ensure(cb) || ctx->stage[1].ensure(fb)) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc staging failed");
  CK(cudaMemcpyAsync(ctx->stage[0].p, coarseQ, cb, cudaMemcpyHostToDevice, st));
  rc = rmcrt_refine_Q_device(ctx, static_cast<const double*>(ctx->stage[0].p), cLo, cHi, gLo, gHi, rr, order, fl, fh,
                             static_cast<double*>(ctx->stage[1].p), oLo, oHi, 1, stream);
  if (rc) return rc;
  CK(cudaMemcpyAsync(fineQ, ctx->stage[1].p, fb, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

static int partgas_ptrs(rmcrt_ctx* ctx, const double* const* absk, const double* const* temp, int nPartGas, PartGasPtrs& pg) {
  if (nPartGas < 1 || nPartGas > RMCRT_MAX_PARTGAS) return fail(ctx, RMCRT_ERR_INVALID, "nPartGas must be in [1,%d]", RMCRT_MAX_PARTGAS);
  pg.n = nPartGas;
  for (int i = 0; i < RMCRT_MAX_PARTGAS; ++i) {
    pg.absk[i] = (absk && i < nPartGas) ? absk[i] : nullptr;
    pg.temp[i] = (temp && i < nPartGas) ? temp[i] : nullptr;
    if (i < nPartGas && ((absk && !absk[i]) || (temp && !temp[i]))) return fail(ctx, RMCRT_ERR_INVALID, "null gas/particle field %d", i);
  }
  return 0;
}

int rmcrt_arches_sigmaT4_device(rmcrt_ctx* ctx, const double* d_gasVolFrac, const double* const* d_absk, const double* const* d_temp,
                                int nPartGas, int doPartRadiation, void* d_out, int isFloat, int64_t n, void* stream) {
  if (!ctx) return RMCRT_ERR_INVALID;
  if (!d_temp || !d_out || n < 0 || (doPartRadiation && (!d_gasVolFrac || !d_absk))) return fail(ctx, RMCRT_ERR_INVALID, "null field pointer");
  PartGasPtrs pg;
  int rc = partgas_ptrs(ctx, doPartRadiation ? d_absk : nullptr, d_temp, nPartGas, pg);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return 0;
  const double sop = ctx->p.sigma / M_PI;
  cudaStream_t st = (cudaStream_t)stream;
  if (isFloat) arches_sigmaT4_kernel<float><<<grid_for(n, 256, ctx->smCount), 256, 0, st>>>(d_gasVolFrac, pg, doPartRadiation, static_cast<float*>(d_out), n, sop);
  else         arches_sigmaT4_kernel<double><<<grid_for(n, 256, ctx->smCount), 256, 0, st>>>(d_gasVolFrac, pg, doPartRadiation, static_cast<double*>(d_out), n, sop);
  CK(cudaGetLastError());
  return 0;
}

int rmcrt_arches_sumAbsk_device(rmcrt_ctx* ctx, const double* d_gasVolFrac, const double* const* d_absk, int nPartGas,
                                const int32_t lo[3], const int32_t hi[3], const int32_t intLo[3], const int32_t intHi[3],
                                void* d_out, int isFloat, void* stream) {
  if (!ctx) return RMCRT_ERR_INVALID;
  if (!d_gasVolFrac || !d_absk || !d_out) return fail(ctx, RMCRT_ERR_INVALID, "null field pointer");
  PartGasPtrs pg;
  int rc = partgas_ptrs(ctx, d_absk, nullptr, nPartGas, pg);
  if (rc) return rc;
  for (int d = 0; d < 3; ++d) if (hi[d] <= lo[d]) return fail(ctx, RMCRT_ERR_INVALID, "empty window");
  CK(cudaSetDevice(ctx->device));
  Box w = mkbox(lo, hi), in = mkbox(intLo, intHi);
  const long long n = box_cells(w);
  cudaStream_t st = (cudaStream_t)stream;
  if (isFloat) arches_sumAbsk_kernel<float><<<grid_for(n, 256, ctx->smCount), 256, 0, st>>>(d_gasVolFrac, pg, w, in, static_cast<float*>(d_out));
  else         arches_sumAbsk_kernel<double><<<grid_for(n, 256, ctx->smCount), 256, 0, st>>>(d_gasVolFrac, pg, w, in, static_cast<double*>(d_out));
  CK(cudaGetLastError());
  return 0;
}

int rmcrt_stencil_to_dbls_device(rmcrt_ctx* ctx, const double* d_boundFlux, int64_t n, double* d_E, double* d_W, double* d_N,
                                 double* d_S, double* d_T, double* d_B, void* stream) {
  if (!ctx) return RMCRT_ERR_INVALID;
  if (!d_boundFlux || !d_E || !d_W || !d_N || !d_S || !d_T || !d_B || n < 0) return fail(ctx, RMCRT_ERR_INVALID, "null field pointer");
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return 0;
  SixPtrs o{{d_E, d_W, d_N, d_S, d_T, d_B}};
  stencil_to_dbls_kernel<<<grid_for(n, 256, ctx->smCount), 256, 0, (cudaStream_t)stream>>>(d_boundFlux, n, o);
  CK(cudaGetLastError());
  return 0;
}

// host-buffer forms of the three Arches tasks (H2D, kernel, D2H; blocking) -- what the C++ RMCRT_Radiation mirror calls
int rmcrt_arches_sigmaT4(rmcrt_ctx* ctx, const double* gasVolFrac, const double* const* absk, const double* const* temp, int nPartGas,
                         int doPartRadiation, void* out, int isFloat, int64_t n, void* stream) {
  if (!ctx) return RMCRT_ERR_INVALID;
  if (!temp || !out || n < 0 || (doPartRadiation && (!gasVolFrac || !absk))) return fail(ctx, RMCRT_ERR_INVALID, "null field pointer");
  if (nPartGas < 1 || nPartGas > RMCRT_MAX_PARTGAS) return fail(ctx, RMCRT_ERR_INVALID, "nPartGas must be in [1,%d]", RMCRT_MAX_PARTGAS);
  if (n == 0) return 0;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t fb = (size_t)n * sizeof(double);
  const int nIn = 1 + 2 * nPartGas;
  if (ctx->stage[0].ensure(fb * nIn) || ctx->stage[1].ensure((size_t)n * (isFloat ? 4 : 8))) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc staging failed");
  double* base = static_cast<double*>(ctx->stage[0].p);
  const double* dA[RMCRT_MAX_PARTGAS]; const double* dT[RMCRT_MAX_PARTGAS];
  if (gasVolFrac) CK(cudaMemcpyAsync(base, gasVolFrac, fb, cudaMemcpyHostToDevice, st));
  for (int i = 0; i < nPartGas; ++i) {
    dA[i] = base + (size_t)n * (1 + i); dT[i] = base + (size_t)n * (1 + nPartGas + i);
    if (!temp[i] || (doPartRadiation && !absk[i])) return fail(ctx, RMCRT_ERR_INVALID, "null gas/particle field %d", i);
    if (doPartRadiation) CK(cudaMemcpyAsync(const_cast<double*>(dA[i]), absk[i], fb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(const_cast<double*>(dT[i]), temp[i], fb, cudaMemcpyHostToDevice, st));
  }
  int rc = rmcrt_arches_sigmaT4_device(ctx, base, dA, dT, nPartGas, doPartRadiation, ctx->stage[1].p, isFloat, n, stream);
  if (rc) return rc;
  CK(cudaMemcpyAsync(out, ctx->stage[1].p, (size_t)n * (isFloat ? 4 : 8), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int rmcrt_arches_sumAbsk(rmcrt_ctx* ctx, const double* gasVolFrac, const double* const* absk, int nPartGas, const int32_t lo[3],
                         const int32_t hi[3], const int32_t intLo[3], const int32_t intHi[3], void* out, int isFloat, void* stream) {
  if (!ctx) return RMCRT_ERR_INVALID;
  if (!gasVolFrac || !absk || !out) return fail(ctx, RMCRT_ERR_INVALID, "null field pointer");
  if (nPartGas < 1 || nPartGas > RMCRT_MAX_PARTGAS) return fail(ctx, RMCRT_ERR_INVALID, "nPartGas must be in [1,%d]", RMCRT_MAX_PARTGAS);
  size_t n = 1;
  for (int d = 0; d < 3; ++d) { if (hi[d] <= lo[d]) return fail(ctx, RMCRT_ERR_INVALID, "empty window"); n *= (size_t)(hi[d] - lo[d]); }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t fb = n * sizeof(double);
  if (ctx->stage[0].ensure(fb * (1 + nPartGas)) || ctx->stage[1].ensure(n * (isFloat ? 4 : 8))) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc staging failed");
  double* base = static_cast<double*>(ctx->stage[0].p);
  const double* dA[RMCRT_MAX_PARTGAS];
  CK(cudaMemcpyAsync(base, gasVolFrac, fb, cudaMemcpyHostToDevice, st));
  for (int i = 0; i < nPartGas; ++i) {
    if (!absk[i]) return fail(ctx, RMCRT_ERR_INVALID, "null gas/particle field %d", i);
    dA[i] = base + n * (1 + i);
    CK(cudaMemcpyAsync(const_cast<double*>(dA[i]), absk[i], fb, cudaMemcpyHostToDevice, st));
  }
  int rc = rmcrt_arches_sumAbsk_device(ctx, base, dA, nPartGas, lo, hi, intLo, intHi, ctx->stage[1].p, isFloat, stream);
  if (rc) return rc;
  CK(cudaMemcpyAsync(out, ctx->stage[1].p, n * (isFloat ? 4 : 8), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int rmcrt_stencil_to_dbls(rmcrt_ctx* ctx, const double* boundFlux, int64_t n, double* E, double* W, double* N, double* S, double* T,
                          double* B, void* stream) {
  if (!ctx) return RMCRT_ERR_INVALID;
  if (!boundFlux || !E || !W || !N || !S || !T || !B || n < 0) return fail(ctx, RMCRT_ERR_INVALID, "null field pointer");
  if (n == 0) return 0;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t fb = (size_t)n * sizeof(double);
  if (ctx->stage[0].ensure(fb * 7) || ctx->stage[1].ensure(fb * 6)) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc staging failed");
  CK(cudaMemcpyAsync(ctx->stage[0].p, boundFlux, fb * 7, cudaMemcpyHostToDevice, st));
  double* o = static_cast<double*>(ctx->stage[1].p);
  int rc = rmcrt_stencil_to_dbls_device(ctx, static_cast<const double*>(ctx->stage[0].p), n, o, o + n, o + 2 * n, o + 3 * n, o + 4 * n, o + 5 * n, stream);
  if (rc) return rc;
  double* host[6] = {E, W, N, S, T, B};
  for (int i = 0; i < 6; ++i) CK(cudaMemcpyAsync(host[i], o + (size_t)i * n, fb, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int rmcrt_pack_box_device(rmcrt_ctx* ctx, int L, const void* a, const void* s, const int32_t* ct, int isFloat,
                          const int32_t srcLo[3], const int32_t srcHi[3], const int32_t boxLo[3], const int32_t boxHi[3],
                          void* d_send, void* stream) {
  int rc = check_level(ctx, L);
  if (rc) return rc;
  if (!a || !s || !ct || !d_send) return fail(ctx, RMCRT_ERR_INVALID, "null pointer");
  for (int d = 0; d < 3; ++d)
    if (boxLo[d] < srcLo[d] || boxHi[d] > srcHi[d] || boxHi[d] <= boxLo[d]) return fail(ctx, RMCRT_ERR_INVALID, "box must lie inside the source window");
  CK(cudaSetDevice(ctx->device));
  ctx->lv[L].propsFloat = isFloat != 0;
  Box src = mkbox(srcLo, srcHi), box = mkbox(boxLo, boxHi);
  // half-packed levels: the float copy of the rank's own cells goes straight into the level (only origin cells need it)
  return do_pack(ctx, ctx->lv[L].recordKind, isFloat, a, s, ct, src, box, d_send, box, lstat_of(ctx, L), ctx->lv[L].recFull,
                 mkbox(ctx->lv[L].d.lo, ctx->lv[L].d.hi), (cudaStream_t)stream);
}

int rmcrt_assemble_level(rmcrt_ctx* ctx, int L, const void* gathered, size_t slotBytes, int nRanks, const int32_t* boxLo,
                         const int32_t* boxHi, void* stream) {
  int rc = check_level(ctx, L);
  if (rc) return rc;
  if (!gathered || nRanks < 1) return fail(ctx, RMCRT_ERR_INVALID, "bad gather buffer");
  CK(cudaSetDevice(ctx->device));
  LevelState& lv = ctx->lv[L];
  Box dst = mkbox(lv.d.lo, lv.d.hi);
  for (int r = 0; r < nRanks; ++r) {
    Box b = mkbox(boxLo + 3 * r, boxHi + 3 * r);
    const long long n = box_cells(b);
    if (n <= 0) continue;
    if ((size_t)n * lv.recBytes > slotBytes) return fail(ctx, RMCRT_ERR_INVALID, "rank %d box does not fit its slot", r);
    for (int d = 0; d < 3; ++d)
      if (b.lo[d] < lv.d.lo[d] || b.lo[d] + b.n[d] > lv.d.hi[d]) return fail(ctx, RMCRT_ERR_INVALID, "rank %d box outside level", r);
    if (lv.recordKind == 0)      assemble_kernel<Rec8><<<grid_for(n, 256, ctx->smCount), 256, 0, (cudaStream_t)stream>>>(static_cast<const unsigned char*>(gathered), slotBytes, r, b, static_cast<Rec8*>(lv.rec), dst);
    else if (lv.recordKind == 2) assemble_kernel<Rec4><<<grid_for(n, 256, ctx->smCount), 256, 0, (cudaStream_t)stream>>>(static_cast<const unsigned char*>(gathered), slotBytes, r, b, static_cast<Rec4*>(lv.rec), dst);
    else                         assemble_kernel<Rec16><<<grid_for(n, 256, ctx->smCount), 256, 0, (cudaStream_t)stream>>>(static_cast<const unsigned char*>(gathered), slotBytes, r, b, static_cast<Rec16*>(lv.rec), dst);
    CK(cudaGetLastError());
  }
  return invalidate_level(ctx, L, (cudaStream_t)stream);
}

int rmcrt_allgather_records(rmcrt_ctx* ctx, const void* d_send, void* d_gathered, size_t slotBytes, void* comm, void* stream) {
  if (!ctx) return RMCRT_ERR_INVALID;
  typedef int (*allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
  static allgather_fn fn = nullptr;
  if (!fn) {
    fn = (allgather_fn)dlsym(RTLD_DEFAULT, "ncclAllGather");
    if (!fn) {
      void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (h) fn = (allgather_fn)dlsym(h, "ncclAllGather");
    }
    if (!fn) return fail(ctx, RMCRT_ERR_NCCL, "ncclAllGather not found (libnccl.so.2 not loadable)");
  }
  const int ncclInt8 = 0;   // ncclInt8 / ncclChar
  const int r = fn(d_send, d_gathered, slotBytes, ncclInt8, comm, (cudaStream_t)stream);
  if (r != 0) return fail(ctx, RMCRT_ERR_NCCL, "ncclAllGather failed with code %d", r);
  return 0;
}

//------------------------------------------------------------------------------------------------------ one process, several devices
// Replication of a level over the devices of ONE process (the C++ host path: one rmcrt_ctx per visible device, patches dealt round robin,
// GPUGridVariableInfo.cc:610-632).  Instead of uploading the whole level to every device (n x the level over PCIe), device d uploads and
// packs only its z-slab (1/n of the level) and the packed records are replicated with ONE ncclAllGather over NVLink, the collective
// SURVEY 8e prescribes for the "all-to-all" ghost requirement of Ray.cc:479-481; then every device assembles the level.
namespace {
struct NcclApi {
  int (*commInitAll)(void**, int, const int*) = nullptr;
  int (*groupStart)() = nullptr;
  int (*groupEnd)() = nullptr;
  int (*allGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*commDestroy)(void*) = nullptr;
  bool ok = false;
  bool tried = false;
} g_nccl;
std::mutex g_ncclMu;
std::map<std::vector<int>, std::vector<void*>> g_ncclComms;      // device list -> communicators

bool load_nccl() {
  if (g_nccl.tried) return g_nccl.ok;
  g_nccl.tried = true;
  void* h = nullptr;
  auto sym = [&](const char* n) -> void* {
    void* p = dlsym(RTLD_DEFAULT, n);
    if (!p) {
      if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (h) p = dlsym(h, n);
    }
    return p;
  };
  g_nccl.commInitAll = (int (*)(void**, int, const int*))sym("ncclCommInitAll");
  g_nccl.groupStart = (int (*)())sym("ncclGroupStart");
  g_nccl.groupEnd = (int (*)())sym("ncclGroupEnd");
  g_nccl.allGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))sym("ncclAllGather");
  g_nccl.commDestroy = (int (*)(void*))sym("ncclCommDestroy");
  g_nccl.ok = g_nccl.commInitAll && g_nccl.groupStart && g_nccl.groupEnd && g_nccl.allGather;
  return g_nccl.ok;
}
}  // namespace

int rmcrt_replicate_level(rmcrt_ctx** ctxs, int n, int L, const void* a, const void* s, const int32_t* ct, int isFloat,
                          const int32_t srcLo[3], const int32_t srcHi[3]) {
  if (!ctxs || n < 1 || !ctxs[0]) return RMCRT_ERR_INVALID;
  rmcrt_ctx* ctx = ctxs[0];
  if (!a || !s || !ct) return fail(ctx, RMCRT_ERR_INVALID, "null property pointer");
  for (int d = 0; d < n; ++d) { int rc = check_level(ctxs[d], L); if (rc) return rc; }
  const LevelState& l0 = ctx->lv[L];
  for (int d = 0; d < 3; ++d)
    if (srcLo[d] != l0.d.lo[d] || srcHi[d] != l0.d.hi[d]) return fail(ctx, RMCRT_ERR_INVALID, "rmcrt_replicate_level needs arrays over the whole level window");
  const int nz = srcHi[2] - srcLo[2];
  // half-packed levels keep a float copy of the cells a device ORIGINATES rays from, which pack_box only writes for the packed box:
  // a device traces patches anywhere in the level here, so those levels (and single devices, short levels) take the plain upload
  bool plain = n == 1 || nz < n || l0.recordKind == 2;
  std::lock_guard<std::mutex> lock(g_ncclMu);
  if (!plain && !load_nccl()) plain = true;
  std::vector<void*>* comms = nullptr;
  if (!plain) {
    std::vector<int> devs(n);
    for (int d = 0; d < n; ++d) devs[d] = ctxs[d]->device;
    auto it = g_ncclComms.find(devs);
    if (it == g_ncclComms.end()) {
      std::vector<void*> c(n, nullptr);
      if (g_nccl.commInitAll(c.data(), n, devs.data()) != 0) plain = true;
      else it = g_ncclComms.emplace(devs, c).first;
    }
    if (!plain) comms = &it->second;
  }
  if (plain) {
    for (int d = 0; d < n; ++d) {
      int rc = rmcrt_upload_level(ctxs[d], L, a, s, ct, isFloat, srcLo, srcHi, nullptr);
      if (rc) return rc;
    }
    return 0;
  }
  // z-slabs: contiguous chunks of the x-fastest arrays
  const size_t plane = (size_t)(srcHi[0] - srcLo[0]) * (srcHi[1] - srcLo[1]);
  const size_t eb = isFloat ? 4 : 8;
  std::vector<int32_t> blo(3 * n), bhi(3 * n);
  size_t slot = 0;
  for (int d = 0; d < n; ++d) {
    const int z0 = srcLo[2] + (int)((long long)nz * d / n), z1 = srcLo[2] + (int)((long long)nz * (d + 1) / n);
    blo[3 * d] = srcLo[0]; blo[3 * d + 1] = srcLo[1]; blo[3 * d + 2] = z0;
    bhi[3 * d] = srcHi[0]; bhi[3 * d + 1] = srcHi[1]; bhi[3 * d + 2] = z1;
    slot = std::max(slot, plane * (size_t)(z1 - z0) * l0.recBytes);
  }
  slot = (slot + 255) / 256 * 256;
  std::vector<cudaEvent_t> copyEvents(n, nullptr);
  for (int d = 0; d < n; ++d) {
    rmcrt_ctx* c = ctxs[d];
    CK(cudaSetDevice(c->device));
    const size_t cells = plane * (size_t)(bhi[3 * d + 2] - blo[3 * d + 2]);
    if (c->stage[0].ensure(cells * eb) || c->stage[1].ensure(cells * eb) || c->stage[2].ensure(cells * 4) || c->replSend.ensure(slot) ||
        c->replGathered.ensure(slot * n))
      return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc replication buffers failed");
    const size_t off = plane * (size_t)(blo[3 * d + 2] - srcLo[2]);
    CK(cudaMemcpyAsync(c->stage[0].p, static_cast<const char*>(a) + off * eb, cells * eb, cudaMemcpyHostToDevice, 0));
    CK(cudaMemcpyAsync(c->stage[1].p, static_cast<const char*>(s) + off * eb, cells * eb, cudaMemcpyHostToDevice, 0));
    CK(cudaMemcpyAsync(c->stage[2].p, ct + off, cells * 4, cudaMemcpyHostToDevice, 0));
    copyEvents[d] = nullptr;
    if (cudaEventCreateWithFlags(&copyEvents[d], cudaEventDisableTiming) == cudaSuccess) cudaEventRecord(copyEvents[d], 0);
    int rc = rmcrt_pack_box_device(c, L, c->stage[0].p, c->stage[1].p, static_cast<const int32_t*>(c->stage[2].p), isFloat, &blo[3 * d], &bhi[3 * d],
                                   &blo[3 * d], &bhi[3 * d], c->replSend.p, nullptr);
    if (rc) { if (c != ctx) ctx->err = c->err; return rc; }
  }
  // the caller's arrays are not read after return: wait for every device's slab copy (they ran concurrently), not for the kernels
  for (int d = 0; d < n; ++d)
    if (copyEvents[d]) { cudaSetDevice(ctxs[d]->device); cudaEventSynchronize(copyEvents[d]); cudaEventDestroy(copyEvents[d]); }
  g_nccl.groupStart();
  int nrc = 0;
  for (int d = 0; d < n; ++d) {
    cudaSetDevice(ctxs[d]->device);
    nrc |= g_nccl.allGather(ctxs[d]->replSend.p, ctxs[d]->replGathered.p, slot, 0 /* ncclInt8 */, (*comms)[d], 0);
  }
  nrc |= g_nccl.groupEnd();
  if (nrc != 0) return fail(ctx, RMCRT_ERR_NCCL, "ncclAllGather (group of %d devices) failed with code %d", n, nrc);
  for (int d = 0; d < n; ++d) {
    int rc = rmcrt_assemble_level(ctxs[d], L, ctxs[d]->replGathered.p, slot, n, blo.data(), bhi.data(), nullptr);
    if (rc) { if (ctxs[d] != ctx) ctx->err = ctxs[d]->err; return rc; }
  }
  return 0;
}

//------------------------------------------------------------------------------------------------------ pinned host pool
namespace {
struct HostPool {
  std::mutex mu;
  std::map<void*, size_t> live;                    // block -> capacity
  std::multimap<size_t, void*> freeList;           // capacity -> block
} g_hostPool;
size_t pool_bucket(size_t bytes) {                 // 64-KiB granularity: windows of neighbouring patches differ by a few cells
  const size_t g = 64u << 10;
  return ((bytes ? bytes : 1) + g - 1) / g * g;
}
}  // namespace

void* rmcrt_host_alloc(size_t bytes) {
  const size_t cap = pool_bucket(bytes);
  {
    std::lock_guard<std::mutex> l(g_hostPool.mu);
    auto it = g_hostPool.freeList.find(cap);
    if (it != g_hostPool.freeList.end()) {
      void* p = it->second;
      g_hostPool.freeList.erase(it);
      g_hostPool.live[p] = cap;
      return p;
    }
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, cap, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  std::lock_guard<std::mutex> l(g_hostPool.mu);
  g_hostPool.live[p] = cap;
  return p;
}

void rmcrt_host_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> l(g_hostPool.mu);
  auto it = g_hostPool.live.find(p);
  if (it == g_hostPool.live.end()) return;         // not ours
  g_hostPool.freeList.emplace(it->second, p);
  g_hostPool.live.erase(it);
}

void rmcrt_host_pool_trim(void) {
  std::lock_guard<std::mutex> l(g_hostPool.mu);
  for (auto& kv : g_hostPool.freeList) cudaFreeHost(kv.second);
  g_hostPool.freeList.clear();
}

int rmcrt_measure_read_bandwidth(rmcrt_ctx* ctx, size_t bytes, int repeats, int pattern, double* gbps, double* gaccess) {
  if (!ctx || bytes < 4096 || repeats < 1) return RMCRT_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  void* buf = nullptr;
  unsigned* sink = nullptr;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(buf, 0x11, bytes));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int grid = ctx->smCount * 16, block = 256;
  const long long n16 = (long long)(bytes / 16);
  const unsigned nRec = (unsigned)std::min<size_t>(bytes / 8, 0xFFFFFFFFu);
  const int perThread = 256;
  // small (L2-resident) buffers: sweep them many times inside one launch so launch latency does not dominate
  const int inner = bytes <= (64u << 20) ? 64 : 1;
  auto pass = [&]() {
    if (pattern == 0) bw_stream_kernel<<<grid, block>>>(static_cast<const uint4*>(buf), n16, inner, sink);
    else if (pattern == 2) bw_stream_cg_kernel<<<ctx->smCount * 4, 512>>>(static_cast<const uint4*>(buf), n16, inner, sink);
    else bw_gather_kernel<<<grid, block>>>(static_cast<const float2*>(buf), nRec, perThread, sink);
  };
  for (int i = 0; i < 3; ++i) pass();     // warm-up (and L2 fill for small buffers)
  CK(cudaEventRecord(e0));
  for (int i = 0; i < repeats; ++i) pass();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  const double sec = ms * 1e-3;
  if (pattern == 2) {      // every thread issues whole groups of eight loads (indices wrap around the buffer)
    const long long per = 8LL * ctx->smCount * 4 * 512;
    const double loads = (double)((n16 + per - 1) / per * per);
    if (gbps) *gbps = loads * 16.0 * repeats * inner / sec / 1e9;
    if (gaccess) *gaccess = loads * repeats * inner / sec / 1e9;
  } else if (pattern == 0) {
    if (gbps) *gbps = (double)n16 * 16.0 * repeats * inner / sec / 1e9;
    if (gaccess) *gaccess = (double)n16 * repeats * inner / sec / 1e9;
  } else {
    const double acc = (double)grid * block * perThread * repeats;
    if (gbps) *gbps = acc * 32.0 / sec / 1e9;
    if (gaccess) *gaccess = acc / sec / 1e9;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf); cudaFree(sink);
  return 0;
}

int rmcrt_debug_trace_rays(rmcrt_ctx* ctx, int L, int maxLevels, const int32_t pLo[3], const int32_t pHi[3], int n,
                           const int32_t* cells, const double* dir, const double* org, double* sumI, uint32_t* steps,
                           int32_t* path, int maxPath) {
  if (!ctx) return RMCRT_ERR_INVALID;
  if (n <= 0) return 0;
  const bool ML = maxLevels > 1;
  const int nl = ML ? maxLevels : L + 1;
  TraceArgs a;
  const rmcrt_level_desc* fd = nullptr;
  int rc = check_level(ctx, nl - 1);
  if (rc) return rc;
  fd = &ctx->lv[nl - 1].d;
  int32_t lo[3], hi[3];
  for (int d = 0; d < 3; ++d) { lo[d] = pLo ? pLo[d] : fd->interiorLo[d]; hi[d] = pHi ? pHi[d] : fd->interiorHi[d]; }
  rc = build_args(ctx, nl, ML, lo, hi, lo, hi, 0, a);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  if (ctx->dbg[0].ensure((size_t)n * 12) || ctx->dbg[1].ensure((size_t)n * 24) || ctx->dbg[2].ensure((size_t)n * 24) ||
      ctx->dbg[3].ensure((size_t)n * 8) || ctx->dbg[4].ensure((size_t)n * 4))
    return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc debug buffers");
  int* dPath = nullptr;
  if (path && maxPath > 0) {
    if (ctx->dbg[5].ensure((size_t)n * maxPath * 16)) return fail(ctx, RMCRT_ERR_CUDA, "cudaMalloc path buffer");
    dPath = static_cast<int*>(ctx->dbg[5].p);
    CK(cudaMemset(dPath, 0, (size_t)n * maxPath * 16));
  }
  CK(cudaMemcpy(ctx->dbg[0].p, cells, (size_t)n * 12, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->dbg[1].p, dir, (size_t)n * 24, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->dbg[2].p, org, (size_t)n * 24, cudaMemcpyHostToDevice));
  const int grid = (n + 63) / 64;
#define DBG_LAUNCH(REAL, REC, MLF)                                                                                         \
  debug_rays_kernel<REAL, REC, MLF><<<grid, 64>>>(a, n, static_cast<const int*>(ctx->dbg[0].p),                            \
      static_cast<const double*>(ctx->dbg[1].p), static_cast<const double*>(ctx->dbg[2].p), static_cast<double*>(ctx->dbg[3].p), \
      static_cast<unsigned*>(ctx->dbg[4].p), dPath, maxPath)
  if (ctx->p.marchPrecision == 1) { if (ML) DBG_LAUNCH(double, Rec16, true); else DBG_LAUNCH(double, Rec16, false); }
  else if (ctx->lv[nl - 1].recordKind == 2) { if (ML) DBG_LAUNCH(float, Rec4, true); else DBG_LAUNCH(float, Rec4, false); }
  else                            { if (ML) DBG_LAUNCH(float, Rec8, true);   else DBG_LAUNCH(float, Rec8, false); }
#undef DBG_LAUNCH
  CK(cudaGetLastError());
  CK(cudaMemcpy(sumI, ctx->dbg[3].p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(steps, ctx->dbg[4].p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (dPath) CK(cudaMemcpy(path, dPath, (size_t)n * maxPath * 16, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
