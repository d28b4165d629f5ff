/*
 * rmcrt_oracle.h -- CPU restatement of the reference RMCRT hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (uintah_b200/, include/)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and there only as the
 * checker / the CPU baseline.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src; RMCRT/ = CCA/Components/Models/Radiation/RMCRT/).
 *
 * Pinning (see DESIGN.md "Oracle"):
 *   - MT19937 stream + rand()/randDblExc() conversions: checked against the
 *     reference's own Core/Math/MersenneTwister.h compiled from where it lies
 *     (oracle/_ref, built by oracle/Makefile) and against tests/golden/mt19937.json
 *     generated from it.
 *   - Burns & Christon exact centre line (parametricStudies/postProcessTools/RMCRT.m:22-38).
 *   - Modest parallel-plate closed form (RMCRT.m:44-94).
 *   - The reference's OWN per-ray routines (raySignStep, findRayDirection, ray_Origin, reflect, updateSumI<T>,
 *     containsCell, updateSumI_ML<T>), cut out of RMCRTCommon.cc / Ray.cc at build time (oracle/extract_ref.py) and
 *     compiled into oracle/_ref/libref_rmcrt.so (oracle/ref_rmcrt_shim.cpp): sumI and step counts of random rays agree
 *     bit for bit (tests/test_oracle_ref.py).
 *   - The exact-solution tables of parametricStudies/postProcessTools (tests/golden/rmcrt_benchmarks.json).
 *   sus itself cannot be built here and the gold-standard UDAs live outside the repo, so whole-deck outputs
 *   of the reference executable are not available; the task-level loops around the per-ray routines are restatements.
 */
#ifndef RMCRT_ORACLE_H
#define RMCRT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mirrors the statics/members read by the trace: Ray.h:162-181, RMCRTCommon.h:287-293 */
typedef struct oracle_params {
  int    nDivQRays;          /* Ray.cc:162 default 10   */
  int    nFluxRays;          /* Ray.cc:168 default 1    */
  double threshold;          /* Ray.cc:163 default 0.01 */
  int    allowReflect;       /* Ray.cc:170 default true */
  int    CCRays;             /* Ray.cc:167 default false*/
  int    isSeedRandom;       /* Ray.cc:164 default true */
  int    solveDivQ;          /* Ray.cc:171 default true */
  int    solveBoundaryFlux;  /* Ray.cc:166 default false*/
  int    fixedRandom;        /* MersenneTwister.h:93 FIXED_RANDOM_NUM: every draw == 0.3 */
  int    flowCell;           /* RMCRTCommon.cc:116, hard-wired -1 */
  int    dataOnionFluxFlowGuard; /* 0 = reference (Ray.cc:1098-1107 has no flow-cell guard) */
  int    rayDirSampleAlgo;   /* Ray.cc:174-179: 0 NAIVE, 1 LATIN_HYPER_CUBE */
  double maxRayLength;       /* RMCRTCommon.cc:117 DBL_MAX; <ROI_extents type="boundedRayLength"><length> (Ray.cc:243) */
  double sigmaScat;          /* <sigmaScat> (Ray.cc:169): > 0 = a RAY_SCATTER build with isotropic scattering (RMCRTCommon.cc:670-676,
                                790-836); 0 = the stock build (no scattering code, no extra random draws) */
  int    roiFixed;           /* <ROI_extents type="fixed"> (Ray.cc:1303-1314): the Data-Onion ROI is the cell box below for every patch */
  int    roiFixedLo[3], roiFixedHi[3];   /* fineLevel->getCellIndex(d_ROI_minPt / d_ROI_maxPt) */
} oracle_params;

/* one <Radiometer> block after Radiometer::problemSetup (Radiometer.cc:96-271, Radiometer.h:117-127) */
typedef struct oracle_radiometer {
  int    cellLo[3], cellHi[3];   /* level->getCellIndex(locationsMin/Max) (Radiometer.cc:448-449) */
  int    nRays;
  double solidAngle, theta_viewAngle, range;       /* 258-260 */
  double phi_rotate, theta_rotate, xi_rotate;      /* 240-254 */
} oracle_radiometer;

/* one level's property arrays: x-fastest Array3 window [lo,hi) (Array3Data.h:93,212-220) */
typedef struct oracle_level {
  const void*    abskg;      /* float or double, see isFloat */
  const void*    sigmaT4;    /* sigma T^4 / pi */
  const int32_t* cellType;
  int    isFloat;
  int    lo[3], hi[3];       /* array window, hi exclusive (includes extra cells) */
  double dx[3];
  double anchor[3];          /* Level::m_anchor (Level.cc:473-477) */
  int    refineRatio[3];     /* ratio to the next COARSER level (Level.cc:1285-1307) */
  int    interiorLo[3], interiorHi[3]; /* findInteriorCellIndexRange */
} oracle_level;

void oracle_default_params(oracle_params* p);

/* ---- MT19937 (MersenneTwister.h:172-258,339-370) ---- */
void oracle_mt_u32(uint32_t seed, int n, uint32_t* out);
void oracle_mt_rand(uint32_t seed, int n, double* out);        /* rand()       [0,1] */
void oracle_mt_randDblExc(uint32_t seed, int n, double* out);  /* randDblExc() (0,1) */

/* ---- one ray, single level: RMCRTCommon::updateSumI (RMCRTCommon.cc:607-904) ----
 * path: optional [maxPath][3] buffer receiving every `cur` visited; nPath receives the count. */
void oracle_updateSumI(const oracle_params* p, const oracle_level* lv,
                       const double dir[3], const double rayOrigin[3], const int originCell[3],
                       double* sumI, uint64_t* nRaySteps,
                       int* path, int maxPath, int* nPath);

/* ---- one ray, multi level: Ray::updateSumI_ML (Ray.cc:2349-2627) ---- */
void oracle_updateSumI_ML(const oracle_params* p, const oracle_level* levels, int maxLevels,
                          const int roiLo[3], const int roiHi[3],
                          const int* regionLo /*[maxLevels][3]*/, const int* regionHi,
                          const double dir[3], const double rayOrigin[3], const int originCell[3],
                          double* sumI, uint64_t* nRaySteps,
                          int* path /*[maxPath][4] = i,j,k,L*/, int maxPath, int* nPath);

/* ---- Ray::rayTrace<T> for one patch [patchLo,patchHi) (Ray.cc:513-818) ----
 * Outputs are x-fastest arrays over window [outLo,outHi); they must be zero-initialised
 * by the caller (Ray.cc:581-587).  boundFlux = Stencil7 AoS, 7 doubles/cell (Stencil7.h:38-76).
 * mtSeed: the per-task MTRand seed when isSeedRandom (replaces hash(time,clock), Ray.cc:539). */
void oracle_rayTrace(const oracle_params* p, const oracle_level* lv,
                     const int patchLo[3], const int patchHi[3], uint32_t mtSeed,
                     double* divQ, double* radiationVolq, double* boundFlux,
                     const int outLo[3], const int outHi[3],
                     uint64_t* nRaySteps, uint64_t* nFluxRaySteps);

/* ---- Radiometer::problemSetup angle derivation (Radiometer.cc:118-120,214-260) from viewAngle [deg] and <unitNormalVector>
 * (isNormalVector = 0) or <normalVector> (1, normalised first); fills the angle members of *r only */
void oracle_radiometer_setup(double viewAngle, const double vec[3], int isNormalVector, oracle_radiometer* r);
/* ---- Radiometer::radiometerFlux<T> for one patch (Radiometer.cc:428-500) + rayDirection_VR (505-560).
 * VRFlux / VRIntensity: arrays of T (lv->isFloat) over [outLo,outHi); only cells of (radiometer box ∩ patch) are written. */
void oracle_radiometerFlux(const oracle_params* p, const oracle_level* lv, const oracle_radiometer* rads, int nRads,
                           const int patchLo[3], const int patchHi[3], uint32_t mtSeed,
                           void* VRFlux, void* VRIntensity, const int outLo[3], const int outHi[3], uint64_t* nRaySteps);
/* ---- RMCRTCommon::randVector (RMCRTCommon.cc:998-1024): Fisher-Yates shuffle of 0..n-1 */
void oracle_randVector(const oracle_params* p, uint32_t mtSeed, const int cell[3], int n, int* out);

/* ---- Ray::computeExtents, patch_based ROI (Ray.cc:1279-1358) ---- */
void oracle_computeExtents(const oracle_level* levels, int maxLevels,
                           const int patchLo[3], const int patchHi[3], const int haloCells[3],
                           int roiLo[3], int roiHi[3], int* regionLo, int* regionHi);
/* the same with a fixed ROI box instead of patch +- halo (Ray.cc:1303-1314); fixedLo == NULL: patch_based */
void oracle_computeExtents_roi(const oracle_level* levels, int maxLevels,
                               const int patchLo[3], const int patchHi[3], const int haloCells[3],
                               const int fixedLo[3], const int fixedHi[3],
                               int roiLo[3], int roiHi[3], int* regionLo, int* regionHi);

/* ---- Ray::rayTrace_dataOnion<T> for one fine patch (Ray.cc:943-1274) ----
 * levels[maxLevels-1] is the fine level; its arrays must cover at least the ROI. */
void oracle_rayTrace_dataOnion(const oracle_params* p, const oracle_level* levels, int maxLevels,
                               const int patchLo[3], const int patchHi[3], const int haloCells[3],
                               uint32_t mtSeed,
                               double* divQ, double* radiationVolq, double* boundFlux,
                               const int outLo[3], const int outHi[3],
                               uint64_t* nRaySteps, uint64_t* nFluxRaySteps);
/* the same task body with the ROI / region extents of the PATCH but only the origin cells of [subLo,subHi) traced
 * (sampling a production-size patch for the at-size parity fixtures) */
void oracle_rayTrace_dataOnion_sub(const oracle_params* p, const oracle_level* levels, int maxLevels,
                                   const int patchLo[3], const int patchHi[3], const int haloCells[3],
                                   const int subLo[3], const int subHi[3], uint32_t mtSeed,
                                   double* divQ, double* radiationVolq, double* boundFlux,
                                   const int outLo[3], const int outHi[3],
                                   uint64_t* nRaySteps, uint64_t* nFluxRaySteps);

/* ---- property preparation ---- */
/* RMCRTCommon::sigmaT4<T> (RMCRTCommon.cc:304-343): out = (sigma/pi) T^2 T^2 */
void oracle_sigmaT4(const double* temp, void* out, int isFloat, int64_t n, double sigma);
/* RMCRT_Test::initialize benchmark 1 (RMCRT_Test.cc:653-680) + Dirichlet walls of RMCRT_bm1_1L.ups:28-93.
 * Fills (N+2)^3 arrays with window lo=-1: abskg, temperature ("color"), cellType. */
void oracle_burns_christon(int N, double* abskg, double* temperature, int32_t* cellType,
                           double T_gas, double abskg_wall, double T_wall, int wallType);
/* Ray::coarsen_Q<T> interior average (Ray.cc:2103-2145 via fineToCoarseOperator, computesAve) */
void oracle_coarsen_interior(const void* fine, int isFloat, const int fLo[3], const int fHi[3],
                             void* coarse, const int cLo[3], const int cHi[3],
                             const int cIntLo[3], const int cIntHi[3], const int rr[3]);
/* Ray::computeCellType (Ray.cc:2272-2345): logic 0 = ROUNDUP, 1 = ROUNDDOWN */
void oracle_computeCellType(const int32_t* fine, const int fLo[3], const int fHi[3],
                            int32_t* coarse, const int cLo[3], const int cHi[3],
                            const int cIntLo[3], const int cIntHi[3], const int rr[3], int logic);

/* Level::mapCellToCoarser including its negative-index behaviour (Level.cc:1285-1307) */
void oracle_mapCellToCoarser(const int idx[3], const int rr[3], int out[3]);

/* ---- Ray::refine_Q (Ray.cc:1872-1945) = selectInterpolator (Core/Grid/AMR.h:708-745), order 0|1|2 ---- */
void oracle_refine_Q(const double* qC, const int cLo[3], const int cHi[3], const int gridLo[3], const int gridHi[3],
                     const int rr[3], int order, const int fl[3], const int fh[3], double* qF, const int oLo[3], const int oHi[3]);

/* ---- Arches RMCRT_Radiation property tasks (Arches/SourceTerms/RMCRT.cc) ---- */
void oracle_arches_sigmaT4(const double* gasVolFrac, const double* const* partGas_absk, const double* const* partGas_temp,
                           int nPartGas, int doPartRadiation, void* out, int isFloat, int64_t n, double sigma);
void oracle_arches_sumAbsk(const double* gasVolFrac, const double* const* partGas_absk, int nPartGas, const int lo[3],
                           const int hi[3], const int intLo[3], const int intHi[3], void* out, int isFloat);
void oracle_arches_stencilToDBLs(const double* boundFlux7, int64_t n, double* E, double* W, double* N, double* S, double* T, double* B);

/* ---- ray generation of the divQ loop (Ray.cc:752-760): findRayDirection + ray_Origin for rays 0..nRays-1 of one cell ---- */
void oracle_generate_rays(const oracle_params* p, const oracle_level* lv, const int cell[3], int nRays, uint32_t mtSeed,
                          double* dirs /*[nRays][3]*/, double* origins);

/* one ray with the generator seeded with mtSeed (the RAY_SCATTER code of updateSumI draws from it inside the march) */
void oracle_updateSumI_seeded(const oracle_params* p, const oracle_level* lv, const double dir[3], const double rayOrigin[3],
                              const int originCell[3], uint32_t mtSeed, double* sumI, uint64_t* nRaySteps);

#ifdef __cplusplus
}
#endif
#endif
