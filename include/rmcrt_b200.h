/*
 * rmcrt_b200.h -- C ABI of the B200-native RMCRT divQ / boundary-flux solve.
 *
 * This is the drop-in boundary for ONE hot path of Uintah: the body of the tasks
 * `Ray::rayTrace<T>` and `Ray::rayTrace_dataOnion<T>` registered by
 * `Ray::sched_rayTrace` / `Ray::sched_rayTrace_dataOnion`.  The host-side `Ray` /
 * `RMCRTCommon` classes (uintah_b200/host/) keep the reference's plugin surface and
 * call these entry points with the raw x-fastest Array3 buffers they obtain from the
 * DataWarehouse (getLevel / getRegion / allocateAndPut).
 *
 * Reference paths: /root/reference/src; RMCRT/ = CCA/Components/Models/Radiation/RMCRT/.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types; every function returns 0 on
 *     success or a negative rmcrt_status, with text in rmcrt_last_error().
 *   - cell boxes are [lo,hi) IntVector triples (x,y,z); arrays are x-fastest over their
 *     window, exactly the memory layout of Uintah's Array3 (Array3Data.h:93,212-220).
 *   - "host" entry points take host pointers and do the H2D/D2H copies themselves;
 *     "_device" entry points take device pointers of the context's device.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  Calls are
 *     asynchronous with respect to the host unless stated otherwise -- but a HOST pointer
 *     passed to an entry point is never read or written after that call has returned
 *     (the DataWarehouse may free or recycle the buffer at once).
 *   - the library never falls back to a CPU path: without a CUDA device every call fails.
 */
#ifndef RMCRT_B200_H
#define RMCRT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMCRT_ABI_VERSION 5
#define RMCRT_MAX_LEVELS 8

typedef struct rmcrt_ctx rmcrt_ctx;

typedef enum rmcrt_status {
  RMCRT_OK = 0,
  RMCRT_ERR_INVALID = -1,      /* bad argument (ProblemSetupException on the C++ side) */
  RMCRT_ERR_CUDA = -2,         /* CUDA runtime error                                    */
  RMCRT_ERR_NONPHYSICAL = -3,  /* non-finite abskg/sigmaT4 (InternalError, RMCRTCommon.cc:267-271) */
  RMCRT_ERR_OPEN_DOMAIN = -4,  /* a flow cell lies on the outer layer of a level: rays could leave it */
  RMCRT_ERR_NCCL = -5
} rmcrt_status;

/* Mirrors the members/statics read by the trace tasks: Ray.h:162-181, RMCRTCommon.h:287-293,
 * with the .ups defaults of Ray::problemSetup (Ray.cc:162-173). */
typedef struct rmcrt_params {
  int32_t  nDivQRays;          /* <nDivQRays>          10    */
  int32_t  nFluxRays;          /* <nFluxRays>          1     */
  double   threshold;          /* <Threshold>          0.01  */
  double   sigma;              /* <StefanBoltzmann>    5.67051e-8 */
  int32_t  allowReflect;       /* <allowReflect>       true  */
  int32_t  CCRays;             /* <CCRays>             false */
  int32_t  solveDivQ;          /* <solveDivQ>          true  */
  int32_t  solveBoundaryFlux;  /* <solveBoundaryFlux>  false */
  int32_t  isSeedRandom;       /* <randomSeed> true: Philox4x32-10 keyed (cell, ray, timestep);
                                  false: the reference's per-ray MT19937 reseed (i+j+k)*iRay+1
                                  (RMCRTCommon.cc:504-506), reproducing its regression mode */
  int32_t  flowCell;           /* -1, RMCRTCommon.cc:116 */
  int32_t  haloCells[3];       /* <haloCells> [10,10,10], Data-Onion ROI = patch +- halo */
  int32_t  cellTypeCoarsenLogic; /* 0 ROUNDUP (default, Ray.h:194), 1 ROUNDDOWN */
  uint64_t seed;               /* mixed into the Philox key (randomSeed=true) */
  /* ---- implementation knobs (no counterpart in the reference) ---- */
  int32_t  marchPrecision;     /* 0 = FP32 march (production), 1 = FP64 march (parity debugging) */
  int32_t  rayCoherence;       /* ray batching of the divQ rays.
                                  0 (DEFAULT) = the reference's sampling: every (cell, ray) draws its own direction and in-cell origin,
                                      Philox4x32-10 keyed on (cell, ray, timestep) -- independent rays per cell like the task-local MTRand
                                      stream of Ray.cc:539,757-771; errors of neighbouring cells are uncorrelated.
                                  1 = the origin cells of a warp tile share one DIRECTION per ray index (own in-cell origins);
                                  2 = they share direction AND in-cell origin, so the rays of a tile are exact translates marching in
                                      lockstep (with cellsPerLane 4: the fastest kernel, 3.3x the default at 128^3).
                                  1 and 2 keep every cell's estimate unbiased with the same per-cell variance, but the errors of the
                                  cells of a tile become (almost fully) correlated: the mean over a 128-cell tile is no better than one
                                  cell (variance 100x the independent-ray value, tests/test_gpu_atsize.py::test_error_correlation).
                                  Opt-in for callers that post-filter or only need per-cell statistics; .ups: <rayCoherence>. */
  int32_t  skipZeroWeightLegs; /* 1 = end a ray once fs==0 and no non-flow cell can emit (exact);
                                  0 = march every leg the reference marches */
  int32_t  fixedRandom;        /* FIXED_RANDOM_NUM (MersenneTwister.h:93): every draw = 0.3 */
  int32_t  l2Persist;          /* 1 = pin the level records in L2 with an access-policy window */
  int32_t  warpsPerTile;       /* 0 = auto; rays of a 32-cell tile are split over this many warps */
  int32_t  genericMarch;       /* 1 = run the FP32 solve through the generic march template instead of the tuned
                                  single-level fast path (A/B measurements, debugging); results agree to FP32 round-off */
  int32_t  maxBlocksPerSM;     /* 0 = one tile per block (hardware balances the tail); > 0 caps the divQ grid at this many
                                  blocks per SM and grid-strides over the tiles (A/B measurements) */
  int32_t  rayDirSampleAlgo;   /* <rayDirSampleAlgo> 0 naive (default), 1 LatinHyperCube (Ray.cc:174-179, 1398-1460) */
  int32_t  cellsPerLane;       /* (rayCoherence >= 1 only) 4 (default): a lane owns the 2 x 2 cells (i, j..j+1, k..k+1); their rays share the
                                  direction AND the in-cell origin, so they are exact translates and ONE DDA drives all four
                                  (divq_quad_kernel: the DDA instructions are paid once per four ray-cell steps).
                                  2: the pair (i,j,k), (i,j+1,k) (divq_dual_kernel).  1: one cell per lane.
                                  Per-ray arithmetic is identical in all three (bit-identical divQ for cells that trace the same rays).
                                  Applies to the production FP32 / Philox single-level solve; all other modes use one cell per lane. */
  double   maxRayLength;       /* <ROI_extents type="boundedRayLength"><length> (Ray.cc:243, RMCRTCommon.cc:680,684);
                                  <= 0 or >= 1e300: unbounded (DBL_MAX, RMCRTCommon.cc:117).  Single-level solves only, like
                                  the reference (updateSumI_ML has no length test) */
  double   sigmaScat;          /* <sigmaScat> 0 (Ray.cc:169).  > 0: isotropic scattering, the code the reference compiles with
                                  --enable-ray-scatter (RAY_SCATTER; RMCRTCommon.cc:670-676, 790-836): exponentially distributed free
                                  paths, a new isotropic direction per event.  Single-level divQ and boundary-flux rays; runs through
                                  the generic march kernels.  Multi-level solves ignore it like the reference (no scattering code in
                                  Ray::updateSumI_ML). */
  int32_t  tileX;              /* implementation knob, cellsPerLane == 4: origin cells of a warp in x (32, 16 or 8; 0 = default 16).  The
                                  32 lanes cover tileX x (32 / tileX) blocks of 2 x 2 cells; a squarer footprint makes the lockstep rays
                                  of a warp reach a wall within fewer steps of each other (fewer idle lanes) at the same number of
                                  sectors per request.  The rays of a cell depend on the tile it falls in, not on the arithmetic. */
  int32_t  nonFlowCellType;    /* integer cell type the coarsening assumes for non-flow fine cells when it rebuilds the type sum of
                                  Ray::computeCellType (Ray.cc:2285-2342) from the packed records: 8 (default; RMCRT_Test.cc:80, Arches walls) */
  int32_t  blockedMarch;       /* kernel of the default sampling (rayCoherence 0; production FP32 / Philox single-level divQ rays).  All four
                                  trace the SAME rays (Philox keyed on (cell, ray, timestep)):
                                  3 (DEFAULT): divq_sorted_kernel -- the march runs on a blocked half-packed copy of the level (2 x 2 x 2 cells
                                     per 32-B sector, fp16 abskg / sigmaT4 like recordKind 2, rebuilt when the level changes) and the rays of
                                     a 32-cell tile (4 x 4 x 2 cells) are sorted by their expected number of steps, so the 32 lanes of a warp end together
                                     (30.6 instead of 13.4 active lanes) and the second visit to a sector is an L1 hit; per cell the rays are
                                     added in ray order in FP64, so divQ does not depend on the sort, the launch shape or the patch layout.
                                     (With 3 ... 16 rays per cell up to eight tiles are sorted together; below 3 the default runs kernel 2.)
                                  2: divq_blk_kernel -- the blocked copy with the tile / warp layout of the one-cell kernel (no sort).
                                  1: divq_indep_kernel -- the blocked copy, one 256-bit load per block entered, per-lane ray refill (A/B).
                                  0: divq_kernel on the level's own records (float or half-packed), one cell per lane.
                                  Levels too large for a 32-bit blocked index, and every other mode (FP64 march, regression generators,
                                  Latin hyper cube, scattering, bounded rays, Data Onion), march the level's own records. */
  int32_t  refillLanes;        /* implementation knob.  blockedMarch 1: lanes of a warp that must wait for a new ray before ray set-up runs
                                  (0 = default 8).  blockedMarch 3: rays of a cell sorted per pass (0 = default min(nDivQRays, 256); at most 2048). */
  int32_t  coarsenExtraCells;  /* <coarsenExtraCells> (Ray.cc:256,315; default false): rmcrt_coarsen_level_device also fills the coarse level's
                                  extra cells on the domain faces with the tangential average of the fine level's extra cells (Ray::coarsen_Q,
                                  Ray.cc:2148-2245) instead of leaving them to rmcrt_set_level_extra_cells (setBoundaryConditions) */
  int32_t  roiFixed;           /* <ROI_extents type="fixed"> (Ray.cc:265-271, 1303-1314): 1 = the Data-Onion fine-level region of interest is the
                                  cell box [roiFixedLo, roiFixedHi) = fineLevel->getCellIndex(min / max point), the same for every patch,
                                  instead of patch +- haloCells (0, patch_based).  Clamped to the fine level like the reference. */
  int32_t  roiFixedLo[3], roiFixedHi[3];
  int32_t  originBrickHalo;    /* A/B knob (0 = off): with genericMarch = 1 and rayCoherence 0 the block stages its origin tile + this many
                                  cells on every side in shared memory and the march reads records inside that brick from there
                                  (divq_brick_kernel; north_star's "shared-memory staging of the per-patch origin block") */
  int32_t  reserved2;
} rmcrt_params;

/* One <Radiometer> block after Radiometer::problemSetup (Radiometer.cc:96-271; struct radiometer, Radiometer.h:117-127) */
typedef struct rmcrt_radiometer {
  int32_t cellLo[3], cellHi[3];   /* level->getCellIndex(locationsMin / locationsMax) (Radiometer.cc:448-449) */
  int32_t nRays;                  /* <nRays> */
  int32_t pad;
  double  solidAngle, theta_viewAngle, range;    /* Radiometer.cc:258-260 */
  double  phi_rotate, theta_rotate, xi_rotate;   /* Radiometer.cc:238-254 */
} rmcrt_radiometer;

/* geometry of one level (Level.cc:473-477, 1285-1307) */
typedef struct rmcrt_level_desc {
  int32_t lo[3], hi[3];                  /* whole index range incl. extra cells (findCellIndexRange) */
  int32_t interiorLo[3], interiorHi[3];  /* findInteriorCellIndexRange */
  double  dx[3];                         /* dCell() */
  double  anchor[3];                     /* position of index (0,0,0)'s low corner */
  int32_t refineRatio[3];                /* ratio to the next coarser level; [1,1,1] on level 0 */
} rmcrt_level_desc;

typedef struct rmcrt_counters {
  uint64_t nRaySteps;        /* executed divQ ray-cell steps  ("Size", Ray.cc:807)        */
  uint64_t nFluxRaySteps;    /* executed boundary-flux steps  (Ray.cc:1262)               */
  uint64_t nDivQCells;       /* flow cells that received a divQ                           */
  uint64_t nFluxCells;       /* cells with boundFlux.p == 1                               */
  uint64_t nRadiometerSteps; /* executed virtual-radiometer ray-cell steps                */
} rmcrt_counters;

/* ------------------------------------------------------------------ life cycle */
int  rmcrt_abi_version(void);
/* sizeof of the ABI structs as compiled (0 rmcrt_params, 1 rmcrt_level_desc, 2 rmcrt_counters, 3 rmcrt_radiometer): lets a
 * foreign-language binding verify its struct layout at load time */
size_t rmcrt_struct_size(int which);
void rmcrt_default_params(rmcrt_params* p);
/* replaces the per-process Ray instance state; device = CUDA ordinal (patch->GPU map of
 * GPUGridVariableInfo.cc:610-632 is the caller's business). */
int  rmcrt_create(int device, const rmcrt_params* p, rmcrt_ctx** out);
int  rmcrt_set_params(rmcrt_ctx* ctx, const rmcrt_params* p);
void rmcrt_destroy(rmcrt_ctx* ctx);
const char* rmcrt_last_error(const rmcrt_ctx* ctx);   /* ctx may be NULL: last create() error */

/* ------------------------------------------------------------------ property levels
 * Replaces getLevel()/getRegion() + the H2D staging of DetailedTask::initiateH2DCopies
 * (DetailedTask.cc:1217, 1326-1340).  Levels are indexed as in Uintah (0 = coarsest).
 * recordKind: 0 = {float abskg|wall-sign, float sigmaT4OverPi} (8 B, production),
 *             1 = {double,double} (16 B, FP64 parity march),
 *             2 = half-packed {fp16 abskg|wall-sign, fp16 sigmaT4OverPi/32} (4 B) for levels that do not fit the L2 as
 *                 8-B records (256^3: 69 MB instead of 137 MB); needs 18 K < T < 3270 K and abskg < 65504 (else
 *                 RMCRT_ERR_NONPHYSICAL); a float copy of the level is kept for the divQ epilogue of the origin cells. */
int rmcrt_define_level(rmcrt_ctx* ctx, int L, const rmcrt_level_desc* desc, int recordKind);

/* Pack abskg / sigmaT4OverPi / cellType given on window [srcLo,srcHi) into the level's device
 * records (replaces the three getLevel gathers of Ray.cc:549-553 and the slim-variant pack task
 * RMCRTCommon.cc:378-427).  isFloat selects T = float|double of the two property arrays. */
int rmcrt_upload_level(rmcrt_ctx* ctx, int L, const void* abskg, const void* sigmaT4OverPi,
                       const int32_t* cellType, int isFloat,
                       const int32_t srcLo[3], const int32_t srcHi[3], void* stream);
int rmcrt_pack_level_device(rmcrt_ctx* ctx, int L, const void* d_abskg, const void* d_sigmaT4OverPi,
                            const int32_t* d_cellType, int isFloat,
                            const int32_t srcLo[3], const int32_t srcHi[3], void* stream);
/* Validate a fully populated level (blocking): non-finite properties, open domain; also decides
 * whether zero-weight legs may be skipped (no emitting non-flow cell). */
int rmcrt_finalize_level(rmcrt_ctx* ctx, int L, void* stream);
/* Same scan without the host round trip: the result stays in device status words that the trace kernels read (the
 * zero-weight-leg skip is enabled only for levels whose status is valid and shows no emitting wall).  Validation
 * errors are then reported by the next blocking rmcrt_finalize_level. */
int rmcrt_finalize_level_async(rmcrt_ctx* ctx, int L, void* stream);
/* Device address/size of the packed records (for the NCCL all-gather of the replicated level). */
int rmcrt_level_records(rmcrt_ctx* ctx, int L, void** d_records, size_t* bytes, size_t* recordBytes);

/* ------------------------------------------------------------------ the hot path
 * rmcrt_trace_divq[_device]: body of Ray::rayTrace<T> for one patch (Ray.cc:556-788):
 * boundary flux (if solveBoundaryFlux) then divQ (if solveDivQ) for cells [patchLo,patchHi) of level L.
 * Outputs are x-fastest over [outLo,outHi) (the patch's allocateAndPut window incl. extra cells):
 *   divQ, radiationVolq : double per cell;  boundFlux : Stencil7 {w,e,s,n,b,t,p} = 7 doubles per cell.
 * zeroInit != 0 reproduces `modifies_divQ == false` (Ray.cc:577-587): the whole window is zeroed first (1), or -- same result,
 * less traffic when the outputs live in mapped page-locked host memory -- only the part of the window outside the patch (2; the
 * kernels write every cell of the patch).  zeroInit == 0 (`modifies_divQ`): cells the trace does not compute, non-flow cells of the
 * patch included, keep the caller's values (Ray.cc:746-748).
 * boundFlux may be NULL when solveBoundaryFlux is off.  timestep enters the Philox counter. */
int rmcrt_trace_divq_device(rmcrt_ctx* ctx, int L, const int32_t patchLo[3], const int32_t patchHi[3],
                            double* d_divQ, double* d_radiationVolq, double* d_boundFlux,
                            const int32_t outLo[3], const int32_t outHi[3],
                            int zeroInit, uint64_t timestep, void* stream);
/* same with host output buffers; blocks until the results are in host memory */
int rmcrt_trace_divq(rmcrt_ctx* ctx, int L, const int32_t patchLo[3], const int32_t patchHi[3],
                     double* divQ, double* radiationVolq, double* boundFlux,
                     const int32_t outLo[3], const int32_t outHi[3],
                     int zeroInit, uint64_t timestep, void* stream);

/* Body of Ray::rayTrace_dataOnion<T> for one fine patch (Ray.cc:1036-1241) with ROI_extents
 * patch_based: ROI = patch +- haloCells clamped to the fine level (Ray::computeExtents, 1279-1358).
 * All levels 0..maxLevels-1 must have been uploaded. */
int rmcrt_trace_dataonion_device(rmcrt_ctx* ctx, int maxLevels, const int32_t patchLo[3], const int32_t patchHi[3],
                                 double* d_divQ, double* d_radiationVolq, double* d_boundFlux,
                                 const int32_t outLo[3], const int32_t outHi[3],
                                 int zeroInit, uint64_t timestep, void* stream);
int rmcrt_trace_dataonion(rmcrt_ctx* ctx, int maxLevels, const int32_t patchLo[3], const int32_t patchHi[3],
                          double* divQ, double* radiationVolq, double* boundFlux,
                          const int32_t outLo[3], const int32_t outHi[3],
                          int zeroInit, uint64_t timestep, void* stream);

/* ------------------------------------------------------------------ virtual radiometer
 * Radiometer::problemSetup's angle derivation (Radiometer.cc:118-120, 214-260): viewAngle in degrees, vec = <unitNormalVector>
 * (isNormalVector 0) or <normalVector> (1: normalised first).  Fills the six angle members of *r; cell box and nRays are the
 * caller's.  Returns RMCRT_ERR_INVALID for viewAngle > 360 (Radiometer.cc:203-207). */
int rmcrt_radiometer_setup(double viewAngle, const double vec[3], int isNormalVector, rmcrt_radiometer* r);
/* Body of Radiometer::radiometerFlux<T> for one patch (Radiometer.cc:428-500; ray directions rayDirection_VR 505-560):
 * for every cell of (radiometer box n patch) nRays cell-centred rays inside the view cone,
 *   VRFlux = (T)sumProjI * solidAngle / nRays,  VRIntensity = (T)sumProjI / nRays   (Radiometer.cc:492-493).
 * VRFlux / VRIntensity are arrays of T (isFloat) x-fastest over [outLo,outHi); cells outside the radiometer boxes are left
 * untouched (modifiesFlux) or zeroed first (zeroInit, Radiometer.cc:443-446).  Single level L. */
int rmcrt_trace_radiometer_device(rmcrt_ctx* ctx, int L, const rmcrt_radiometer* rads, int nRads,
                                  const int32_t patchLo[3], const int32_t patchHi[3],
                                  void* d_VRFlux, void* d_VRIntensity, int isFloat,
                                  const int32_t outLo[3], const int32_t outHi[3],
                                  int zeroInit, uint64_t timestep, void* stream);
/* same with host output buffers (blocking) */
int rmcrt_trace_radiometer(rmcrt_ctx* ctx, int L, const rmcrt_radiometer* rads, int nRads,
                           const int32_t patchLo[3], const int32_t patchHi[3],
                           void* VRFlux, void* VRIntensity, int isFloat,
                           const int32_t outLo[3], const int32_t outHi[3],
                           int zeroInit, uint64_t timestep, void* stream);

/* Ray::computeExtents (patch_based) on the host, exported so callers/tests see the same boxes. */
int rmcrt_compute_extents(rmcrt_ctx* ctx, int maxLevels, const int32_t patchLo[3], const int32_t patchHi[3],
                          int32_t roiLo[3], int32_t roiHi[3],
                          int32_t* regionLo /*[maxLevels][3]*/, int32_t* regionHi);

/* counters accumulate over trace calls until reset; get blocks on `stream` */
int rmcrt_reset_counters(rmcrt_ctx* ctx, void* stream);
int rmcrt_get_counters(rmcrt_ctx* ctx, rmcrt_counters* out, void* stream);

/* ------------------------------------------------------------------ property pipeline on device
 * sigmaT4OverPi = (sigma/pi) T^4 (RMCRTCommon.cc:304-343); n cells; T of the output = isFloat */
int rmcrt_sigmaT4_device(rmcrt_ctx* ctx, const double* d_temperature, void* d_sigmaT4OverPi, int isFloat,
                         int64_t n, void* stream);
/* same with host buffers (H2D, kernel, D2H; blocking): the body of the RMCRTCommon::sigmaT4<T> task for one patch */
int rmcrt_sigmaT4(rmcrt_ctx* ctx, const double* temperature, void* sigmaT4OverPi, int isFloat, int64_t n, void* stream);
/* Data-Onion coarse level from the fine level's records: Ray::coarsen_Q interior average (Ray.cc:2103-2145)
 * + Ray::computeCellType (2272-2345) for interior coarse cells; extra cells of the coarse level keep what
 * rmcrt_upload_level / rmcrt_set_level_extra_cells put there. */
int rmcrt_coarsen_level_device(rmcrt_ctx* ctx, int fineL, int coarseL, void* stream);
/* Dirichlet fill of a level's extra-cell layer (Ray::setBoundaryConditions, Ray.cc:1662-1742) per face
 * x-,x+,y-,y+,z-,z+: abskg[6], sigmaT4OverPi[6], cellType[6]. */
int rmcrt_set_level_extra_cells(rmcrt_ctx* ctx, int L, const double abskg[6], const double sigmaT4OverPi[6],
                                const int32_t cellType[6], void* stream);

/* ------------------------------------------------------------------ callers either side of the path
 * Ray::refine_Q (Ray.cc:1872-1945), the last task of <algorithm type="RMCRT_coarseLevel">: interpolate one coarse-level field
 * (divQ, radiationVolq) to the interior cells [fl,fh) of one fine patch with selectInterpolator (Core/Grid/AMR.h:708-745):
 * order 0 piecewise constant, 1 linear, 2 quadratic.  coarseQ is x-fastest over [cLo,cHi) -- the getRegion box, which must
 * cover the mapped cells +- 1 for order > 0 (getCoarseLevelRange with one ghost cell, AMR.cc:90-135); gridLo/gridHi =
 * coarseLevel->findCellIndexRange() incl. extra cells; rr = the fine level's refinement ratio.  fineQ is x-fastest over
 * [outLo,outHi) (the fine patch's allocateAndPut window); zeroInit reproduces `divQ_fine.initialize(0.0)` (Ray.cc:1899).
 * FP64 in the reference's operation order: results are bit-identical to the CPU interpolators. */
int rmcrt_refine_Q_device(rmcrt_ctx* ctx, const double* d_coarseQ, const int32_t cLo[3], const int32_t cHi[3],
                          const int32_t gridLo[3], const int32_t gridHi[3], const int32_t rr[3], int order,
                          const int32_t fl[3], const int32_t fh[3], double* d_fineQ,
                          const int32_t outLo[3], const int32_t outHi[3], int zeroInit, void* stream);
/* same with host buffers (blocking); the fine window is zero-initialised first, like the task */
int rmcrt_refine_Q(rmcrt_ctx* ctx, const double* coarseQ, const int32_t cLo[3], const int32_t cHi[3],
                   const int32_t gridLo[3], const int32_t gridHi[3], const int32_t rr[3], int order,
                   const int32_t fl[3], const int32_t fh[3], double* fineQ,
                   const int32_t outLo[3], const int32_t outHi[3], void* stream);

/* The three property tasks of the Arches caller, RMCRT_Radiation (CCA/Components/Arches/SourceTerms/RMCRT.cc), over one patch.
 * partGas arrays: [0] = gas phase (radiation_temperature / gas absk), [1..] = particle size classes; nPartGas <= 16.
 *   sigmaT4  (RMCRT.cc:939-1016): doPartRadiation == 0 -> (sigma/pi) T_gas^4; else the absk-weighted mean of T^4 over gas and
 *            particles in cells with gasVolFrac > 1e-16 and (sigma/pi) T_gas^4 in walls / intrusions.  n cells, T = isFloat.
 *   sumAbsk  (RMCRT.cc:1062-1130): sum of the gas and particle absorption coefficients in interior gas cells, 1 in solid cells
 *            and on the boundary faces (hard-wired in the reference).  Arrays x-fastest over [lo,hi), interior [intLo,intHi).
 *   stencilToDBLs (RMCRT.cc:1389-1423): RMCRTboundFlux Stencil7 {w,e,s,n,b,t,p} -> radiationFlux{E,W,N,S,T,B}. */
int rmcrt_arches_sigmaT4_device(rmcrt_ctx* ctx, const double* d_gasVolFrac, const double* const* d_partGas_absk,
                                const double* const* d_partGas_temp, int nPartGas, int doPartRadiation,
                                void* d_sigmaT4OverPi, int isFloat, int64_t n, void* stream);
int rmcrt_arches_sumAbsk_device(rmcrt_ctx* ctx, const double* d_gasVolFrac, const double* const* d_partGas_absk, int nPartGas,
                                const int32_t lo[3], const int32_t hi[3], const int32_t intLo[3], const int32_t intHi[3],
                                void* d_sumAbsk, int isFloat, void* stream);
int rmcrt_stencil_to_dbls_device(rmcrt_ctx* ctx, const double* d_boundFlux, int64_t n, double* d_E, double* d_W, double* d_N,
                                 double* d_S, double* d_T, double* d_B, void* stream);
/* host-buffer forms (H2D, kernel, D2H; blocking) */
int rmcrt_arches_sigmaT4(rmcrt_ctx* ctx, const double* gasVolFrac, const double* const* partGas_absk,
                         const double* const* partGas_temp, int nPartGas, int doPartRadiation,
                         void* sigmaT4OverPi, int isFloat, int64_t n, void* stream);
int rmcrt_arches_sumAbsk(rmcrt_ctx* ctx, const double* gasVolFrac, const double* const* partGas_absk, int nPartGas,
                         const int32_t lo[3], const int32_t hi[3], const int32_t intLo[3], const int32_t intHi[3],
                         void* sumAbsk, int isFloat, void* stream);
int rmcrt_stencil_to_dbls(rmcrt_ctx* ctx, const double* boundFlux, int64_t n, double* E, double* W, double* N, double* S,
                          double* T, double* B, void* stream);

/* ------------------------------------------------------------------ multi-GPU replication
 * Each rank packs the cells it owns, box [lo,hi) of level L, into its slot of an all-gather send buffer
 * (records in x-fastest order of the box); after the collective, rmcrt_assemble_level scatters every rank's
 * box into the level records.  The collective itself is the caller's (NCCL via torch.distributed in the
 * python harness; ncclAllGather in a C++ host) -- or rmcrt_allgather_level when an ncclComm_t is at hand. */
int rmcrt_pack_box_device(rmcrt_ctx* ctx, int L, const void* d_abskg, const void* d_sigmaT4OverPi,
                          const int32_t* d_cellType, int isFloat,
                          const int32_t srcLo[3], const int32_t srcHi[3],
                          const int32_t boxLo[3], const int32_t boxHi[3],
                          void* d_sendRecords, void* stream);
int rmcrt_assemble_level(rmcrt_ctx* ctx, int L, const void* d_gathered, size_t slotBytes, int nRanks,
                         const int32_t* boxLo /*[nRanks][3]*/, const int32_t* boxHi, void* stream);
/* ncclAllGather(sendSlot -> gathered) on `comm` (an ncclComm_t passed as void*), resolved with dlsym at
 * first use so the library has no link-time NCCL dependency. */
int rmcrt_allgather_records(rmcrt_ctx* ctx, const void* d_sendSlot, void* d_gathered, size_t slotBytes,
                            void* ncclComm, void* stream);

/* ------------------------------------------------------------------ dynamic ROI (Ray::ROI_Extents, Ray.cc:1984-2043)
 * For one fine patch [patchLo,patchHi) of level L (records already on the device): |grad abskg| and |grad sigmaT4/pi| by first-order
 * central differences (compute_Mag_gradient, Core/Grid/AMR.cc:464-483, in the precision of the level's records) and the flag
 * `mag_grad_abskg > abskgThreshold || mag_grad_sigmaT4 > sigmaT4Threshold` per cell.  *anyFlagged = 1 when a cell of the patch is
 * flagged: the caller then extends the ROI by the patch's extra-cell box (Ray.cc:2032-2038) and, after all patches, hands the box to
 * the solve as rmcrt_params.roiFixed* (Ray::computeExtents, dynamic branch, Ray.cc:1291-1300: the same box for every patch).
 * Optional host outputs (all three or none) over [outLo,outHi), zero outside the patch: the reference's mag_grad_abskg /
 * mag_grad_sigmaT4 (float or double as the records) / flaggedCells (int) visualisation variables.  Blocking. */
int rmcrt_roi_extents(rmcrt_ctx* ctx, int L, const int32_t patchLo[3], const int32_t patchHi[3], double abskgThreshold,
                      double sigmaT4Threshold, void* magGradAbskg, void* magGradSigmaT4, int32_t* flaggedCells, int isFloat,
                      const int32_t outLo[3], const int32_t outHi[3], int32_t* anyFlagged, void* stream);

/* ------------------------------------------------------------------ one process driving several devices
 * rmcrt_replicate_level: what n calls of rmcrt_upload_level (one per context / device) achieve -- level L of every context holds the
 * packed records of the host arrays -- with 1/n of the PCIe traffic: device d uploads and packs only its z-slab of the level and ONE
 * ncclAllGather (single-process communicators, created on first use) replicates the packed records over NVLink; then every device
 * assembles the level (Ray.cc:479-481: every patch needs the whole level).  Arrays must cover the whole level window.
 * Falls back to plain per-device uploads when NCCL is not loadable, for n == 1 and for half-packed levels.  Blocking on the host only
 * as far as the CUDA calls are; all work is queued on each device's default stream. */
int rmcrt_replicate_level(rmcrt_ctx** ctxs, int n, int L, const void* abskg, const void* sigmaT4, const int32_t* cellType, int isFloat,
                          const int32_t srcLo[3], const int32_t srcHi[3]);

/* ------------------------------------------------------------------ page-locked host memory for the DataWarehouse side
 * The host entry points (rmcrt_upload_level, rmcrt_trace_divq, rmcrt_trace_dataonion) are fastest when the caller's grid variables live
 * in page-locked, device-mapped memory: property arrays upload at full PCIe rate and divQ / radiationVolq are written by the kernels in
 * place (no staging buffer, no D2H copy after the solve).  These two functions are an allocator for such memory with a size-bucketed
 * free list (page-locking costs ~0.3 ms per MB, so blocks are recycled, not returned to the driver): the hook for the framework's
 * Array3Data allocation (Array3Data.h:93-125; the stand-in DataWarehouse installs it in Ray::devices()).  Thread safe.
 * rmcrt_host_alloc returns NULL without a CUDA device; rmcrt_host_free(NULL) is a no-op; rmcrt_host_pool_trim releases the free list. */
void* rmcrt_host_alloc(size_t bytes);
void  rmcrt_host_free(void* p);
void  rmcrt_host_pool_trim(void);

/* ------------------------------------------------------------------ roofline denominators
 * Measures device read bandwidth on a scratch buffer of `bytes` (L2-resident when well below the L2 size, HBM when far
 * above it), `repeats` passes, CUDA-event timed.  pattern 0: coalesced 16-B loads (streaming);
 * pattern 1: one dependent-free 8-B gather per lane at a hashed record index -- every lane touches its own 32-B sector,
 * the access shape of an incoherent ray step; *gbps counts 32 B per gather (sector traffic), *gaccess the gathers/s;
 * pattern 2: eight independent 16-B L2-only loads (ld.global.cg) in flight per thread, whole lines per warp request -- the L2 -> SM
 * bandwidth the L2 roofline is quoted against. */
int rmcrt_measure_read_bandwidth(rmcrt_ctx* ctx, size_t bytes, int repeats, int pattern, double* gbps, double* gaccess);

/* ------------------------------------------------------------------ debugging (cf. isDbgCell, RMCRTCommon.cc:38,632-635)
 * Trace n caller-specified rays through level L (single level) or the Data-Onion stack (maxLevels > 1)
 * with the production march code.  Host pointers; blocking.
 *   originCell[n][3], dir[n][3], rayOrigin[n][3] (physical position);  out: sumI[n], steps[n]
 *   path: optional [n][maxPath][4] int32 (i,j,k,level) of every cell entered. */
int rmcrt_debug_trace_rays(rmcrt_ctx* ctx, int L, int maxLevels, const int32_t patchLo[3], const int32_t patchHi[3],
                           int n, const int32_t* originCell, const double* dir, const double* rayOrigin,
                           double* sumI, uint32_t* steps, int32_t* path, int maxPath);

#ifdef __cplusplus
}
#endif
#endif /* RMCRT_B200_H */
