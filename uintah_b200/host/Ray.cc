// Ray.cc -- see Ray.h.  Reference behaviour followed: RMCRT/Ray.cc (line ranges cited per method).
#include "Ray.h"

#include <chrono>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <thread>

namespace Uintah {

//______________________________________________________________________
// Ray.cc:82-126: labels local to Ray.  The face index / sign tables of the reference (112-125) moved into the
// boundary-flux kernel (gen_flux_ray in csrc/rmcrt_b200.cu).
Ray::Ray(TypeDescription::Type FLT_DBL) : RMCRTCommon(FLT_DBL) {
  m_timeStepLabel = VarLabel::create("timeStep", TypeDescription::int_type);
  d_mag_grad_abskgLabel = VarLabel::create("mag_grad_abskg", FLT_DBL);
  d_mag_grad_sigmaT4Label = VarLabel::create("mag_grad_sigmaT4", FLT_DBL);
  d_flaggedCellsLabel = VarLabel::create("flaggedCells", TypeDescription::int_type);
  d_ROI_LoCellLabel = VarLabel::create("ROI_loCell", TypeDescription::int_type);
  d_ROI_HiCellLabel = VarLabel::create("ROI_hiCell", TypeDescription::int_type);
  d_PPTimerLabel = VarLabel::create("Ray_PPTimer", TypeDescription::double_type);
}

Ray::~Ray() {
  delete d_radiometer;   // Ray.cc:145-147
  VarLabel::destroy(m_timeStepLabel);
  VarLabel::destroy(d_mag_grad_abskgLabel);
  VarLabel::destroy(d_mag_grad_sigmaT4Label);
  VarLabel::destroy(d_flaggedCellsLabel);
  VarLabel::destroy(d_ROI_LoCellLabel);
  VarLabel::destroy(d_ROI_HiCellLabel);
  VarLabel::destroy(d_PPTimerLabel);
}

//______________________________________________________________________
// Ray.cc:154-364: same keys, same defaults, same exceptions
void Ray::problemSetup(const ProblemSpecP& prob_spec, const ProblemSpecP& rmcrtps, const GridP& grid) {
  ProblemSpecP rmcrt_ps = rmcrtps;
  std::string rayDirSampleAlgo;
  d_grid = grid;

  rmcrt_ps->getWithDefault("nDivQRays", d_nDivQRays, 10);
  rmcrt_ps->getWithDefault("Threshold", d_threshold, 0.01);
  rmcrt_ps->getWithDefault("randomSeed", d_isSeedRandom, true);
  rmcrt_ps->getWithDefault("StefanBoltzmann", d_sigma, 5.67051e-8);
  rmcrt_ps->getWithDefault("solveBoundaryFlux", d_solveBoundaryFlux, false);
  rmcrt_ps->getWithDefault("CCRays", d_CCRays, false);
  rmcrt_ps->getWithDefault("nFluxRays", d_nFluxRays, 1);
  rmcrt_ps->getWithDefault("sigmaScat", d_sigmaScat, 0.0);
  rmcrt_ps->getWithDefault("allowReflect", d_allowReflect, true);
  rmcrt_ps->getWithDefault("solveDivQ", d_solveDivQ, true);
  rmcrt_ps->getWithDefault("applyFilter", d_applyFilter, false);
  rmcrt_ps->getWithDefault("rayDirSampleAlgo", rayDirSampleAlgo, "naive");
  // back-end keys (no counterpart in the reference; absent from a stock deck => the reference's sampling):
  //   <rayCoherence> 0 (default) independent rays per (cell, ray, timestep); 1 / 2 opt into warp-batched directions (rmcrt_b200.h)
  //   <cellsPerLane> 1 | 2 | 4 cells of a lane sharing one DDA when rayCoherence >= 1
  int rayCoherence = 0, cellsPerLane = 4;
  rmcrt_ps->getWithDefault("rayCoherence", rayCoherence, 0);
  rmcrt_ps->getWithDefault("cellsPerLane", cellsPerLane, 4);
  if (rayCoherence < 0 || rayCoherence > 2)
    throw ProblemSetupException("RMCRT (B200 back end): <rayCoherence> must be 0, 1 or 2", __FILE__, __LINE__);
  if (cellsPerLane != 1 && cellsPerLane != 2 && cellsPerLane != 4)
    throw ProblemSetupException("RMCRT (B200 back end): <cellsPerLane> must be 1, 2 or 4", __FILE__, __LINE__);
  if (rayCoherence != 0) {
    d_knobs["rayCoherence"] = rayCoherence;
    std::cout << "  - B200 back end: rayCoherence " << rayCoherence << " (ray directions shared by the origin cells of a warp tile; errors of\n"
              << "    neighbouring cells are correlated -- not the reference's sampling)\n";
  }
  if (cellsPerLane != 4) d_knobs["cellsPerLane"] = cellsPerLane;

  if (rayDirSampleAlgo == "LatinHyperCube") {                       // Ray.cc:175-180
    d_rayDirSampleAlgo = LATIN_HYPER_CUBE;
    std::cout << "  - Using Latin Hyper Cube method for selecting ray directions.\n";
  } else {
    std::cout << "  - Using traditional Monte-Carlo method for selecting ray directions.\n";
  }
  if (rmcrt_ps->findBlock("Radiometer")) {                          // Ray.cc:183-188
    d_radiometer = new Radiometer(d_FLT_DBL, this);
    d_radiometer->problemSetup(prob_spec, rmcrtps, grid);
  }

  // Ray.cc:192-211.  The stock reference refuses sigmaScat > 0 unless it was configured with --enable-ray-scatter; this back end always
  // carries the scattering march (the RAY_SCATTER code of updateSumI), selected at run time, so a scattering deck just runs
  std::cout << "    - sigmaScat: " << d_sigmaScat << "\n";
  if (d_sigmaScat > 0) {
    std::cout << "  - Ray scattering is enabled.\n";
    if (d_radiometer)   // Radiometer::radiometerFlux calls updateSumI as well; its kernel has no scattering variant
      throw ProblemSetupException("RMCRT (B200 back end): <sigmaScat> > 0 together with a <Radiometer> is not supported", __FILE__, __LINE__);
  }
  if (d_nDivQRays == 1) {
    std::cout << "    WARNING: You have specified only 1 ray to compute the radiative flux divergence.\n"
              << "              For higher accuracy specify nDivQRays greater than 2.\n";
  }
  if (d_nFluxRays == 1 && d_whichAlgo != radiometerOnly)
    std::cout << "    WARNING: You have specified only 1 ray to compute radiative fluxes on the boundaries.\n";

  //__________________________________ the algorithm section (228-317)
  bool isMultilevel = false;
  Algorithm algorithm = singleLevel;
  ProblemSpecP alg_ps = rmcrt_ps->findBlock("algorithm");
  if (alg_ps) {
    std::string type = "nullptr";
    if (!alg_ps->getAttribute("type", type))
      throw ProblemSetupException("RMCRT: No type specified for algorithm.  Please choose singleLevel, dataOnion or RMCRT_coarseLevel", __FILE__, __LINE__);

    if (type == "singleLevel") {
      ProblemSpecP ROI_ps = alg_ps->findBlock("ROI_extents");
      if (ROI_ps) {
        ROI_ps->getAttribute("type", type);
        ROI_ps->get("length", d_maxRayLength);
        d_ROI_algo = boundedRayLength;
        std::cout << "  RMCRT: The bounded ray length has been set to: (" << d_maxRayLength << ")\n";
      }
    } else if (type == "dataOnion") {
      isMultilevel = true;
      algorithm = dataOnion;
      alg_ps->getWithDefault("haloCells", d_haloCells, IntVector(10, 10, 10));
      alg_ps->get("haloLength", d_haloLength);
      alg_ps->get("coarsenExtraCells", d_coarsenExtraCells);
      ProblemSpecP ROI_ps = alg_ps->findBlock("ROI_extents");
      if (!ROI_ps) throw ProblemSetupException("RMCRT: dataOnion needs an <ROI_extents> block", __FILE__, __LINE__);
      ROI_ps->getAttribute("type", type);
      if (type == "fixed") {
        d_ROI_algo = fixed;
        ROI_ps->get("min", d_ROI_minPt);
        ROI_ps->get("max", d_ROI_maxPt);
      } else if (type == "dynamic") {
        d_ROI_algo = dynamic;
        ROI_ps->getWithDefault("abskgd_threshold", d_abskg_thld, DBL_MAX);
        ROI_ps->getWithDefault("sigmaT4d_threshold", d_sigmaT4_thld, DBL_MAX);
      } else if (type == "patch_based") {
        d_ROI_algo = patch_based;
      }
    } else if (type == "RMCRT_coarseLevel") {
      isMultilevel = true;
      algorithm = coarseLevel;
      d_ROI_algo = entireDomain;
      alg_ps->require("orderOfInterpolation", d_orderOfInterpolation);
      alg_ps->get("coarsenExtraCells", d_coarsenExtraCells);
    }
  }

  //__________________________________ cellType coarsening logic (321-333)
  if (isMultilevel) {
    std::string tmp = "nullptr";
    rmcrt_ps->get("cellTypeCoarsenLogic", tmp);
    if (tmp == "ROUNDDOWN") d_cellTypeCoarsenLogic = ROUNDDOWN;
    if (tmp == "ROUNDUP") d_cellTypeCoarsenLogic = ROUNDUP;
  }

  //__________________________________ bulletproofing (Ray.cc:338-345): the reference's own device build refuses every ROI but patch_based.
  // This back end takes all three Data-Onion ROI types: the whole fine level is resident on the device, so the ROI is only a box test
  // in the march; fixed and dynamic ROIs are one box for every patch (Ray.cc:1291-1314).
  d_sigma_over_pi = d_sigma / M_PI;
}

//______________________________________________________________________
// Ray.cc:370-413
void Ray::BC_bulletproofing(const ProblemSpecP& rmcrtps, const bool chk_temp, const bool chk_absk) {
  if (d_onOff_SetBCs == false) return;
  ProblemSpecP rmcrt_ps = rmcrtps;
  bool ignore_BC_bulletproofing = false;
  rmcrt_ps->get("ignore_BC_bulletproofing", ignore_BC_bulletproofing);
  if (ignore_BC_bulletproofing) {
    std::cout << "  WARNING: bulletproofing of the boundary conditions specs is off!\n";
    return;
  }
  ProblemSpecP root_ps = rmcrt_ps->getRootNode();
  auto is_BC_specified = [&](const std::string& label) {
    BCEntry e;
    for (int f = 0; f < 6; ++f)
      if (!d_grid->getBC(f, label, e)) {
        static const char* names[6] = {"x-", "x+", "y-", "y+", "z-", "z+"};
        throw ProblemSetupException("\nERROR: The boundary condition for (" + label + ") was not specified on face (" + names[f] + ")", __FILE__, __LINE__);
      }
  };
  if (chk_temp) is_BC_specified(d_compTempLabel->getName());
  if (chk_absk) is_BC_specified(d_abskgBC_tag);
  ProblemSpecP grid_ps = root_ps->findBlock("Grid");
  ProblemSpecP level_ps = grid_ps ? grid_ps->findBlock("Level") : nullptr;
  Vector periodic(0, 0, 0);
  if (level_ps) level_ps->getWithDefault("periodic", periodic, Vector(0, 0, 0));
  if (periodic.length() != 0)
    throw ProblemSetupException("\nERROR RMCRT:\nPeriodic boundary conditions are not allowed with Reverse Monte-Carlo Ray Tracing.", __FILE__, __LINE__);
}

//______________________________________________________________________
void Ray::setKnob(const std::string& name, int value) {
  static const char* known[] = {"rayCoherence", "skipZeroWeightLegs", "l2Persist", "genericMarch", "marchPrecision", "warpsPerTile", "fixedRandom", "halfRecords", "cellsPerLane", "tileX", "blockedMarch"};
  for (auto k : known)
    if (name == k) { d_knobs[name] = value; return; }
  throw ProblemSetupException("RMCRT (B200 back end): unknown knob " + name, __FILE__, __LINE__);
}

void Ray::fillParams(rmcrt_params& p) const {
  RMCRTCommon::fillParams(p);
  p.nDivQRays = d_nDivQRays;
  p.nFluxRays = d_nFluxRays;
  p.CCRays = d_CCRays ? 1 : 0;
  p.solveDivQ = d_solveDivQ ? 1 : 0;
  p.solveBoundaryFlux = d_solveBoundaryFlux ? 1 : 0;
  for (int d = 0; d < 3; ++d) p.haloCells[d] = d_haloCells[d] > 0 ? d_haloCells[d] : 10;
  p.cellTypeCoarsenLogic = d_cellTypeCoarsenLogic == ROUNDDOWN ? 1 : 0;
  p.seed = d_seed;
  p.rayDirSampleAlgo = d_rayDirSampleAlgo == LATIN_HYPER_CUBE ? 1 : 0;
  p.maxRayLength = d_maxRayLength;
  p.sigmaScat = d_sigmaScat;
  p.roiFixed = 0;
  if (d_ROI_algo == fixed && d_grid) {                      // Ray::computeExtents, fixed branch (Ray.cc:1303-1314)
    const Level* fineLevel = d_grid->getLevel(d_grid->numLevels() - 1).get();
    const IntVector lo = fineLevel->getCellIndex(d_ROI_minPt), hi = fineLevel->getCellIndex(d_ROI_maxPt);
    if (!fineLevel->containsCell(lo) || !fineLevel->containsCell(hi)) {
      std::ostringstream warn;
      warn << "ERROR:  the fixed ROI extents " << d_ROI_minPt << " " << d_ROI_maxPt << " are not contained on the fine level.\n";
      throw ProblemSetupException(warn.str(), __FILE__, __LINE__);
    }
    p.roiFixed = 1;
    for (int d = 0; d < 3; ++d) { p.roiFixedLo[d] = lo[d]; p.roiFixedHi[d] = hi[d]; }
  }
  if (d_ROI_algo == dynamic) {                              // Ray::computeExtents, dynamic branch (Ray.cc:1291-1300): one box for every patch
    p.roiFixed = 1;
    for (int d = 0; d < 3; ++d) { p.roiFixedLo[d] = d_dynROI_lo[d]; p.roiFixedHi[d] = d_dynROI_hi[d]; }
  }
  p.coarsenExtraCells = d_coarsenExtraCells ? 1 : 0;     // Ray::coarsen_Q's domain-face branch (Ray.cc:2148-2245) runs inside rmcrt_coarsen_level_device
  for (auto& kv : d_knobs) {
    if (kv.first == "rayCoherence") p.rayCoherence = kv.second;
    else if (kv.first == "skipZeroWeightLegs") p.skipZeroWeightLegs = kv.second;
    else if (kv.first == "l2Persist") p.l2Persist = kv.second;
    else if (kv.first == "genericMarch") p.genericMarch = kv.second;
    else if (kv.first == "marchPrecision") p.marchPrecision = kv.second;
    else if (kv.first == "warpsPerTile") p.warpsPerTile = kv.second;
    else if (kv.first == "fixedRandom") p.fixedRandom = kv.second;
    else if (kv.first == "cellsPerLane") p.cellsPerLane = kv.second;
    else if (kv.first == "tileX") p.tileX = kv.second;
    else if (kv.first == "blockedMarch") p.blockedMarch = kv.second;
  }
}

//______________________________________________________________________
// Ray.cc:422-506: same task name, same dependencies, same task graph
void Ray::sched_rayTrace(const LevelP& level, SchedulerP& sched, Task::WhichDW /*notUsed*/, Task::WhichDW sigma_dw,
                         Task::WhichDW celltype_dw, bool modifies_divQ) {
  std::string taskname = "Ray::rayTrace";
  Task* tsk = nullptr;
  const int L = level->getIndex();
  Task::WhichDW abskg_dw = get_abskg_whichDW(L, d_abskgLabel);   // the argument is ignored, like the reference (436-437)

  if (d_FLT_DBL == TypeDescription::double_type)
    tsk = new Task(taskname, this, &Ray::rayTrace<double>, modifies_divQ, abskg_dw, sigma_dw, celltype_dw);
  else
    tsk = new Task(taskname, this, &Ray::rayTrace<float>, modifies_divQ, abskg_dw, sigma_dw, celltype_dw);

  // an "infinite" number of ghost cells: the whole level is needed around every patch (449-481)
  Ghost::GhostType gac = Ghost::AroundCells;
  int n_ghostCells = SHRT_MAX;
  if (d_ROI_algo == boundedRayLength) {                             // Ray.cc:455-475
    const Vector Dx = level->dCell();
    const Vector nCells = Vector(d_maxRayLength) / Dx;
    n_ghostCells = (int)std::ceil(nCells.length());
    IntVector lo, hi;
    level->findCellIndexRange(lo, hi);
    const IntVector diff = hi - lo;
    const int nCellsLevel = Max(diff.x(), diff.y(), diff.z());
    if (n_ghostCells > SHRT_MAX || n_ghostCells > nCellsLevel) {
      std::cout << "\n  WARNING  RMCRT:sched_rayTrace Clamping the number of ghostCells to SHRT_MAX, (n_ghostCells: " << n_ghostCells
                << ") max cells in any direction on a Levels: " << nCellsLevel << "\n\n";
      n_ghostCells = SHRT_MAX;
      d_ROI_algo = entireDomain;
    }
    d_haloCells = IntVector(n_ghostCells, n_ghostCells, n_ghostCells);
  }
  if (!isDeviceLevel(level.get())) {
    tsk->requiresVar(abskg_dw, d_abskgLabel, gac, n_ghostCells);
    tsk->requiresVar(sigma_dw, d_sigmaT4Label, gac, n_ghostCells);
    tsk->requiresVar(celltype_dw, d_cellTypeLabel, gac, n_ghostCells);
  }   // else (RMCRT_coarseLevel): the coarse level's records were derived on the device by Ray::coarsenAll, scheduled before this task
  if (modifies_divQ) {
    tsk->modifiesVar(d_divQLabel);
    tsk->modifiesVar(d_boundFluxLabel);
    tsk->modifiesVar(d_radiationVolqLabel);
  } else {
    tsk->computesVar(d_divQLabel);
    tsk->computesVar(d_boundFluxLabel);
    tsk->computesVar(d_radiationVolqLabel);
  }
  tsk->usesDevice(true, 4);   // RayKokkos.cc:498-501
  sched->addTask(tsk, level->eachPatch(), d_matlSet, RMCRTCommon::TG_RMCRT);
}

//______________________________________________________________________
// Ray.cc:824-938
void Ray::sched_rayTrace_dataOnion(const LevelP& level, SchedulerP& sched, Task::WhichDW /*notUsed*/, Task::WhichDW sigma_dw,
                                   Task::WhichDW celltype_dw, bool modifies_divQ) {
  const int maxLevels = level->getGrid()->numLevels() - 1;
  const int L_indx = level->getIndex();
  if (L_indx != maxLevels) return;   // only schedule on the finest level (836-841)

  std::string taskname = "Ray::rayTrace_dataOnion";
  Task::WhichDW abskg_dw = get_abskg_whichDW(L_indx, d_abskgLabel);
  Task* tsk = nullptr;
  if (d_FLT_DBL == TypeDescription::double_type)
    tsk = new Task(taskname, this, &Ray::rayTrace_dataOnion<double>, modifies_divQ, abskg_dw, sigma_dw, celltype_dw);
  else
    tsk = new Task(taskname, this, &Ray::rayTrace_dataOnion<float>, modifies_divQ, abskg_dw, sigma_dw, celltype_dw);

  Ghost::GhostType gac = Ghost::AroundCells;
  // patch_based ROI: the fine level is needed within patch +- halo (862-887)
  int maxElem = Max(d_haloCells.x(), d_haloCells.y(), d_haloCells.z());
  tsk->requiresVar(abskg_dw, d_abskgLabel, gac, maxElem);
  tsk->requiresVar(sigma_dw, d_sigmaT4Label, gac, maxElem);
  tsk->requiresVar(celltype_dw, d_cellTypeLabel, gac, maxElem);
  // every coarser level, all of it (904-914)
  Task::DomainSpec ND = Task::NormalDomain;
  for (int l = 0; l < maxLevels; ++l) {
    int offset = maxLevels - l;
    tsk->requiresVar(Task::NewDW, d_abskgLabel, nullptr, Task::CoarseLevel, offset, nullptr, ND, gac, SHRT_MAX);
    tsk->requiresVar(sigma_dw, d_sigmaT4Label, nullptr, Task::CoarseLevel, offset, nullptr, ND, gac, SHRT_MAX);
    tsk->requiresVar(celltype_dw, d_cellTypeLabel, nullptr, Task::CoarseLevel, offset, nullptr, ND, gac, SHRT_MAX);
  }
  if (modifies_divQ) {
    tsk->modifiesVar(d_divQLabel);
    tsk->modifiesVar(d_boundFluxLabel);
    tsk->modifiesVar(d_radiationVolqLabel);
  } else {
    tsk->computesVar(d_divQLabel);
    tsk->computesVar(d_boundFluxLabel);
    tsk->computesVar(d_radiationVolqLabel);
  }
  tsk->usesDevice(true, 4);
  sched->addTask(tsk, level->eachPatch(), d_matlSet, RMCRTCommon::TG_RMCRT);
}

//______________________________________________________________________
// 0 float records, 1 double records (FP64 parity march), 2 half-packed records (knob halfRecords; for levels beyond the L2)
int Ray::recordKind(const rmcrt_params& p) const {
  if (p.marchPrecision == 1) return 1;
  auto it = d_knobs.find("halfRecords");
  return (it != d_knobs.end() && it->second) ? 2 : 0;
}

void Ray::forEachPatchOnDevices(const PatchSubset* patches, const std::function<void(Device&, int)>& fn) {
  auto& devs = devices();
  const int nDev = (int)devs.size();
  std::vector<std::thread> workers;
  std::vector<std::string> errors(nDev);
  for (int d = 0; d < nDev; ++d) {
    workers.emplace_back([&, d]() {
      try {
        for (int p = d; p < patches->size(); p += nDev) fn(devs[d], p);   // GPUGridVariableInfo.cc:610-632: round robin
      } catch (const std::exception& e) { errors[d] = e.what(); }
    });
  }
  for (auto& w : workers) w.join();
  for (auto& e : errors) if (!e.empty()) throw InternalError(e, __FILE__, __LINE__);
}

//______________________________________________________________________
// the body shared by Ray::rayTrace<T> (Ray.cc:513-818) and Ray::rayTrace_dataOnion<T> (Ray.cc:943-1274)
template <class T>
void Ray::traceCommon(bool dataOnion, const PatchSubset* patches, DataWarehouse* new_dw, bool modifies_divQ) {
  const Level* level = getLevel(patches);
  const int L = level->getIndex();
  const int maxLevels = level->getGrid()->numLevels();
  auto& devs = devices();
  const auto tTask = std::chrono::steady_clock::now();

  if (dataOnion && d_ROI_algo == dynamic) {     // this solve's ROI_loCell / ROI_hiCell (Ray::ROI_Extents ran after the level was uploaded)
    if (d_dynROI_stamp != new_dw->getID())
      throw InternalError("Ray::rayTrace_dataOnion: dynamic ROI without Ray::ROI_Extents this time step (sched_ROI_Extents missing)", __FILE__, __LINE__);
    rmcrt_params p;
    fillParams(p);
    for (auto& dev : devs) check(dev.ctx, rmcrt_set_params(dev.ctx, &p), "rmcrt_set_params");
  }
  for (auto& dev : devs) {
    check(dev.ctx, rmcrt_finalize_level(dev.ctx, L, nullptr), "rmcrt_finalize_level");   // non-finite abskg etc. -> InternalError
    if (dataOnion)
      for (int l = 0; l < L; ++l) check(dev.ctx, rmcrt_finalize_level(dev.ctx, l, nullptr), "rmcrt_finalize_level");
    check(dev.ctx, rmcrt_reset_counters(dev.ctx, nullptr), "rmcrt_reset_counters");
  }

  // allocate / fetch the outputs on the host thread (the DataWarehouse is not thread safe in the shim)
  struct Out { CCVariable<double> divQ, volq; CCVariable<Stencil7> flux; };
  std::vector<Out> outs(patches->size());
  for (int p = 0; p < patches->size(); ++p) {
    const Patch* patch = patches->get(p);
    Out& o = outs[p];
    if (modifies_divQ) {
      new_dw->getModifiable(o.divQ, d_divQLabel, d_matl, patch);
      new_dw->getModifiable(o.flux, d_boundFluxLabel, d_matl, patch);
      new_dw->getModifiable(o.volq, d_radiationVolqLabel, d_matl, patch);
    } else {                                                               // Ray.cc:573-588
      new_dw->allocateAndPut(o.divQ, d_divQLabel, d_matl, patch);
      new_dw->allocateAndPut(o.flux, d_boundFluxLabel, d_matl, patch);
      new_dw->allocateAndPut(o.volq, d_radiationVolqLabel, d_matl, patch);
      o.divQ.initialize(0.0);
      o.volq.initialize(0.0);
      Stencil7 zero; zero.initialize(0.0);
      o.flux.initialize(zero);
    }
  }
  const uint64_t timestep = (uint64_t)new_dw->getTimeStep();
  // the reference starts its timer after the outputs are allocated and initialised (Ray.cc:590 `timer.start()` follows 573-588)
  const auto t0 = std::chrono::steady_clock::now();

  forEachPatchOnDevices(patches, [&](Device& dev, int p) {
    const Patch* patch = patches->get(p);
    Out& o = outs[p];
    const IntVector lo = patch->getCellLowIndex(), hi = patch->getCellHighIndex();
    const IntVector olo = o.divQ.getLowIndex(), ohi = o.divQ.getHighIndex();
    double* flux = d_solveBoundaryFlux ? reinterpret_cast<double*>(o.flux.getPointer()) : nullptr;
    const int zeroInit = modifies_divQ ? 0 : 1;
    if (dataOnion)
      check(dev.ctx, rmcrt_trace_dataonion(dev.ctx, maxLevels, lo.get_pointer(), hi.get_pointer(), o.divQ.getPointer(), o.volq.getPointer(),
                                           flux, olo.get_pointer(), ohi.get_pointer(), zeroInit, timestep, nullptr), "rmcrt_trace_dataonion");
    else
      check(dev.ctx, rmcrt_trace_divq(dev.ctx, L, lo.get_pointer(), hi.get_pointer(), o.divQ.getPointer(), o.volq.getPointer(), flux,
                                      olo.get_pointer(), ohi.get_pointer(), zeroInit, timestep, nullptr), "rmcrt_trace_divq");
  });

  d_report = Report();
  for (auto& dev : devs) {
    rmcrt_counters c;
    check(dev.ctx, rmcrt_get_counters(dev.ctx, &c, nullptr), "rmcrt_get_counters");
    d_report.nRaySteps += c.nRaySteps; d_report.nFluxRaySteps += c.nFluxRaySteps;
    d_report.nDivQCells += c.nDivQCells; d_report.nFluxCells += c.nFluxCells;
  }
  d_report.devicesUsed = (int)std::min<size_t>(devs.size(), (size_t)patches->size());
  d_report.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const double taskSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tTask).count();
  // Ray.cc:803-812 prints this for patch 0
  const unsigned long long size = d_report.nRaySteps + d_report.nFluxRaySteps;
  std::cout << "\n RMCRT REPORT: Patch " << patches->get(0)->getID() << " +" << patches->size() - 1 << " more on " << d_report.devicesUsed
            << " GPU(s)\n Used " << d_report.seconds * 1000 << " milliseconds of wall time (trace kernels + D2H of divQ / radiationVolq / boundFlux; "
            << taskSeconds * 1000 << " ms for the whole task incl. level validation and allocateAndPut of the outputs)\n"
            << " Size: " << size << " ray-cell steps\n Efficiency: " << size / d_report.seconds << " steps per sec\n" << std::endl;
}

//______________________________________________________________________
// Ray.cc:513-818
template <class T>
void Ray::rayTrace(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse* /*old_dw*/,
                   DataWarehouse* new_dw, bool modifies_divQ, Task::WhichDW which_abskg_dw, Task::WhichDW which_sigmaT4_dw,
                   Task::WhichDW which_celltype_dw) {
  const Level* level = getLevel(patches);
  DataWarehouse* abskg_dw = new_dw->getOtherDataWarehouse(which_abskg_dw);
  DataWarehouse* sigmaT4_dw = new_dw->getOtherDataWarehouse(which_sigmaT4_dw);
  DataWarehouse* celltype_dw = new_dw->getOtherDataWarehouse(which_celltype_dw);
  if (isDeviceLevel(level)) {                                       // RMCRT_coarseLevel: trace the level Ray::coarsenAll derived
    auto it = d_deviceCoarsened.find(level->getIndex());
    if (it == d_deviceCoarsened.end() || it->second != new_dw->getID())
      throw InternalError("Ray::rayTrace: level " + std::to_string(level->getIndex()) + " was not coarsened this time step (sched_CoarsenAll missing)", __FILE__, __LINE__);
  } else {
    uploadLevel<T>(level, abskg_dw, sigmaT4_dw, celltype_dw, new_dw->getID());
  }
  if (d_radiometer)                                                 // Ray.cc:638-640
    for (int p = 0; p < patches->size(); ++p) d_radiometer->radiometerFlux<T>(patches->get(p), level, new_dw, modifies_divQ);
  traceCommon<T>(false, patches, new_dw, modifies_divQ);
}

//______________________________________________________________________
// Ray.cc:943-1274.  The coarse levels were derived on the device by coarsenAll (sched_CoarsenAll) and their extra cells
// set by setBoundaryConditions; the ROI / region boxes of Ray::computeExtents are evaluated inside rmcrt_trace_dataonion.
template <class T>
void Ray::rayTrace_dataOnion(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse* /*old_dw*/,
                             DataWarehouse* new_dw, bool modifies_divQ, Task::WhichDW which_abskg_dw,
                             Task::WhichDW which_sigmaT4_dw, Task::WhichDW which_celltype_dw) {
  const Level* fineLevel = getLevel(patches);
  DataWarehouse* abskg_dw = new_dw->getOtherDataWarehouse(which_abskg_dw);
  DataWarehouse* sigmaT4_dw = new_dw->getOtherDataWarehouse(which_sigmaT4_dw);
  DataWarehouse* celltype_dw = new_dw->getOtherDataWarehouse(which_celltype_dw);
  uploadLevel<T>(fineLevel, abskg_dw, sigmaT4_dw, celltype_dw, new_dw->getID());
  for (int l = 0; l < fineLevel->getIndex(); ++l) {
    auto it = d_deviceCoarsened.find(l);
    if (it == d_deviceCoarsened.end() || it->second != new_dw->getID())
      throw InternalError("Ray::rayTrace_dataOnion: level " + std::to_string(l) + " was not coarsened this time step (sched_CoarsenAll missing)", __FILE__, __LINE__);
  }
  traceCommon<T>(true, patches, new_dw, modifies_divQ);
}

//______________________________________________________________________
// Ray.cc:2048-2248 (sched_CoarsenAll -> coarsen_Q of abskg and sigmaT4) and 2252-2345 (computeCellType): one device kernel
// derives the coarse level's packed records from the finer level's (rmcrt_coarsen_level_device).  The coarse arrays are
// consumed only by the trace, so they are not copied back into the DataWarehouse.
void Ray::sched_CoarsenAll(const LevelP& coarseLevel, SchedulerP& sched, const bool /*modifies_abskg*/, const bool /*modifies_sigmaT4*/) {
  if (!coarseLevel->hasFinerLevel()) return;
  Task* tsk = nullptr;
  if (d_FLT_DBL == TypeDescription::double_type) tsk = new Task("Ray::coarsenAll", this, &Ray::coarsenAll<double>);
  else                                           tsk = new Task("Ray::coarsenAll", this, &Ray::coarsenAll<float>);
  Ghost::GhostType gn = Ghost::None;
  Task::DomainSpec ND = Task::NormalDomain;
  tsk->requiresVar(Task::NewDW, d_abskgLabel, nullptr, Task::FineLevel, 1, nullptr, ND, gn, 0);
  tsk->requiresVar(Task::NewDW, d_sigmaT4Label, nullptr, Task::FineLevel, 1, nullptr, ND, gn, 0);
  tsk->requiresVar(Task::NewDW, d_cellTypeLabel, nullptr, Task::FineLevel, 1, nullptr, ND, gn, 0);
  tsk->usesDevice(true);
  sched->addTask(tsk, coarseLevel->eachPatch(), d_matlSet, RMCRTCommon::TG_RMCRT);
}

template <class T>
void Ray::coarsenAll(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw) {
  const Level* coarse = getLevel(patches);
  const Level* fine = coarse->getFinerLevel();
  const int cL = coarse->getIndex(), fL = fine->getIndex();
  const long long stamp = new_dw->getID();
  // the finest level comes from the DataWarehouse; intermediate levels were derived on the device by an earlier coarsenAll
  if (!fine->hasFinerLevel()) {
    DataWarehouse* abskg_dw = get_abskg_dw(fL, d_abskgLabel, new_dw);
    uploadLevel<T>(fine, abskg_dw, new_dw, new_dw, stamp);
  }
  const rmcrt_level_desc desc = describe(coarse);
  rmcrt_params p;
  fillParams(p);
  for (auto& dev : devices()) {
    check(dev.ctx, rmcrt_define_level(dev.ctx, cL, &desc, recordKind(p)), "rmcrt_define_level");
    check(dev.ctx, rmcrt_coarsen_level_device(dev.ctx, fL, cL, nullptr), "rmcrt_coarsen_level_device");
    dev.levelStamp[cL] = stamp;
  }
  d_deviceCoarsened[cL] = stamp;
}

void Ray::sched_computeCellType(const LevelP&, SchedulerP&, const Ray::modifiesComputes) {
  // fused into the device coarsening kernel (ROUNDUP / ROUNDDOWN, Ray.cc:2315-2342); nothing to schedule
}

// Ray.cc:1947-1978: only on the finest level and only for the dynamic ROI
void Ray::sched_ROI_Extents(const LevelP& level, SchedulerP& scheduler) {
  const int maxLevels = level->getGrid()->numLevels() - 1;
  if (level->getIndex() != maxLevels || d_ROI_algo != dynamic) return;
  Task* tsk = nullptr;
  if (d_FLT_DBL == TypeDescription::double_type) tsk = new Task("Ray::ROI_Extents", this, &Ray::ROI_Extents<double>);
  else                                           tsk = new Task("Ray::ROI_Extents", this, &Ray::ROI_Extents<float>);
  Ghost::GhostType gac = Ghost::AroundCells;
  tsk->requiresVar(Task::NewDW, d_abskgLabel, gac, 1);
  tsk->requiresVar(Task::NewDW, d_sigmaT4Label, gac, 1);
  tsk->computesVar(d_mag_grad_abskgLabel);
  tsk->computesVar(d_mag_grad_sigmaT4Label);
  tsk->computesVar(d_flaggedCellsLabel);
  tsk->computesVar(d_ROI_LoCellLabel);
  tsk->computesVar(d_ROI_HiCellLabel);
  tsk->usesDevice(true);
  scheduler->addTask(tsk, level->eachPatch(), d_matlSet, RMCRTCommon::TG_RMCRT);
}

// Ray.cc:1984-2043.  The gradient magnitudes and the flags are computed on the device from the level's packed records
// (rmcrt_roi_extents: compute_Mag_gradient + the threshold test); the union of the flagged patches' extra-cell boxes is the
// ROI_loCell / ROI_hiCell reduction, kept in d_dynROI_* (the stand-in DataWarehouse has no reduction variables).
template <class T>
void Ray::ROI_Extents(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw) {
  const Level* level = getLevel(patches);
  const int L = level->getIndex();
  DataWarehouse* abskg_dw = get_abskg_dw(L, d_abskgLabel, new_dw);
  uploadLevel<T>(level, abskg_dw, new_dw, new_dw, new_dw->getID());
  if (d_dynROI_stamp != new_dw->getID()) {      // a new radiation solve: restart the min / max reduction
    d_dynROI_lo = IntVector(SHRT_MAX, SHRT_MAX, SHRT_MAX);
    d_dynROI_hi = IntVector(-SHRT_MAX, -SHRT_MAX, -SHRT_MAX);
    d_dynROI_stamp = new_dw->getID();
  }
  Device& dev = devices()[0];
  for (int p = 0; p < patches->size(); ++p) {
    const Patch* patch = patches->get(p);
    CCVariable<T> mag_grad_abskg, mag_grad_sigmaT4;
    CCVariable<int> flaggedCells;
    new_dw->allocateAndPut(mag_grad_abskg, d_mag_grad_abskgLabel, 0, patch);
    new_dw->allocateAndPut(mag_grad_sigmaT4, d_mag_grad_sigmaT4Label, 0, patch);
    new_dw->allocateAndPut(flaggedCells, d_flaggedCellsLabel, 0, patch);
    const IntVector lo = patch->getCellLowIndex(), hi = patch->getCellHighIndex();
    const IntVector olo = mag_grad_abskg.getLowIndex(), ohi = mag_grad_abskg.getHighIndex();
    rmcrt_params prm;
    fillParams(prm);
    const bool doubleRecords = recordKind(prm) == 1;      // FP64 records <=> the gradients come back as double
    int32_t flaggedPatch = 0;
    if (doubleRecords == (sizeof(T) == 8)) {
      check(dev.ctx, rmcrt_roi_extents(dev.ctx, L, lo.get_pointer(), hi.get_pointer(), d_abskg_thld, d_sigmaT4_thld, mag_grad_abskg.getPointer(),
                                       mag_grad_sigmaT4.getPointer(), flaggedCells.getPointer(), sizeof(T) == 4 ? 1 : 0, olo.get_pointer(),
                                       ohi.get_pointer(), &flaggedPatch, nullptr), "rmcrt_roi_extents");
    } else {      // double deck marched in FP32: float records, the visualisation variables are widened on the host
      CCVariable<float> ga, gs;
      ga.resize(olo, ohi); gs.resize(olo, ohi);
      check(dev.ctx, rmcrt_roi_extents(dev.ctx, L, lo.get_pointer(), hi.get_pointer(), d_abskg_thld, d_sigmaT4_thld, ga.getPointer(), gs.getPointer(),
                                       flaggedCells.getPointer(), 1, olo.get_pointer(), ohi.get_pointer(), &flaggedPatch, nullptr), "rmcrt_roi_extents");
      mag_grad_abskg.copyRegion(ga);
      mag_grad_sigmaT4.copyRegion(gs);
    }
    if (flaggedPatch) {                                   // Ray.cc:2032-2038: include BCs
      d_dynROI_lo = Min(d_dynROI_lo, patch->getExtraCellLowIndex());
      d_dynROI_hi = Max(d_dynROI_hi, patch->getExtraCellHighIndex());
    }
  }
}
template void Ray::ROI_Extents<double>(const ProcessorGroup*, const PatchSubset*, const MaterialSubset*, DataWarehouse*, DataWarehouse*);
template void Ray::ROI_Extents<float>(const ProcessorGroup*, const PatchSubset*, const MaterialSubset*, DataWarehouse*, DataWarehouse*);

// a level whose property records exist only on the device: every level below the finest one of a multi-level algorithm
bool Ray::isDeviceLevel(const Level* level) const {
  return level->hasFinerLevel() && (d_whichAlgo == dataOnion || d_whichAlgo == coarseLevel);
}

//______________________________________________________________________
// Ray.cc:1841-1867: same task name and dependencies (coarse-level divQ / boundFlux / radiationVolq with one ghost cell)
void Ray::sched_Refine_Q(SchedulerP& sched, const PatchSet* patches, const MaterialSet* matls) {
  const Level* fineLevel = getLevel(patches);
  const int L_indx = fineLevel->getIndex();
  if (L_indx > 0) {
    Task* task = new Task("Ray::refine_Q", this, &Ray::refine_Q);
    Task::DomainSpec ND = Task::NormalDomain;
    Ghost::GhostType gac = Ghost::AroundCells;
    task->requiresVar(Task::NewDW, d_divQLabel, nullptr, Task::CoarseLevel, 1, nullptr, ND, gac, 1);
    task->requiresVar(Task::NewDW, d_boundFluxLabel, nullptr, Task::CoarseLevel, 1, nullptr, ND, gac, 1);
    task->requiresVar(Task::NewDW, d_radiationVolqLabel, nullptr, Task::CoarseLevel, 1, nullptr, ND, gac, 1);
    task->computesVar(d_divQLabel);
    task->computesVar(d_boundFluxLabel);
    task->computesVar(d_radiationVolqLabel);
    task->usesDevice(true);
    sched->addTask(task, patches, matls, RMCRTCommon::TG_RMCRT);
  }
}

//______________________________________________________________________
// Ray.cc:1872-1945.  The interpolation (selectInterpolator, Core/Grid/AMR.h:708-745) runs in refine_kernel through rmcrt_refine_Q;
// like the reference, the Stencil7 boundary flux is only zero-initialised on the fine level (its interpolation is "#if 0"-ed out).
void Ray::refine_Q(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw) {
  const Level* fineLevel = getLevel(patches);
  const Level* coarseLevel = fineLevel->getCoarserLevel();
  Device& dev = devices()[0];
  IntVector gridLo, gridHi;
  coarseLevel->findCellIndexRange(gridLo, gridHi);
  const IntVector refineRatio = fineLevel->getRefinementRatio();
  for (int p = 0; p < patches->size(); p++) {
    const Patch* finePatch = patches->get(p);
    CCVariable<double> divQ_fine, radVolQ_fine;
    CCVariable<Stencil7> boundFlux_fine;
    new_dw->allocateAndPut(divQ_fine, d_divQLabel, d_matl, finePatch);
    new_dw->allocateAndPut(radVolQ_fine, d_radiationVolqLabel, d_matl, finePatch);
    new_dw->allocateAndPut(boundFlux_fine, d_boundFluxLabel, d_matl, finePatch);
    Stencil7 zero; zero.initialize(0.0);
    boundFlux_fine.initialize(zero);
    // getCoarseLevelRange (Core/Grid/AMR.cc:90-135) with one ghost cell, exclusive range: the coarse cells under the patch incl. its
    // extra cells and one more layer, clamped to the coarse level; the fine region is the patch interior
    const IntVector fl = finePatch->getCellLowIndex(), fh = finePatch->getCellHighIndex();
    const IntVector el = finePatch->getExtraCellLowIndex() - IntVector(1, 1, 1), eh = finePatch->getExtraCellHighIndex() + IntVector(1, 1, 1);
    IntVector cl = fineLevel->mapCellToCoarser(el), ch = fineLevel->mapCellToCoarser(eh);
    for (int d = 0; d < 3; ++d) {
      if (ch[d] != eh[d]) ch[d] += 1;
      cl[d] = std::max(cl[d], gridLo[d]);
      ch[d] = std::min(ch[d], gridHi[d]);
    }
    const IntVector olo = divQ_fine.getLowIndex(), ohi = divQ_fine.getHighIndex();
    constCCVariable<double> divQ_coarse, radVolQ_coarse;
    new_dw->getRegion(divQ_coarse, d_divQLabel, d_matl, coarseLevel, cl, ch);
    new_dw->getRegion(radVolQ_coarse, d_radiationVolqLabel, d_matl, coarseLevel, cl, ch);
    check(dev.ctx, rmcrt_refine_Q(dev.ctx, divQ_coarse.getPointer(), cl.get_pointer(), ch.get_pointer(), gridLo.get_pointer(), gridHi.get_pointer(),
                                  refineRatio.get_pointer(), d_orderOfInterpolation, fl.get_pointer(), fh.get_pointer(), divQ_fine.getPointer(),
                                  olo.get_pointer(), ohi.get_pointer(), nullptr), "rmcrt_refine_Q");
    check(dev.ctx, rmcrt_refine_Q(dev.ctx, radVolQ_coarse.getPointer(), cl.get_pointer(), ch.get_pointer(), gridLo.get_pointer(), gridHi.get_pointer(),
                                  refineRatio.get_pointer(), d_orderOfInterpolation, fl.get_pointer(), fh.get_pointer(), radVolQ_fine.getPointer(),
                                  olo.get_pointer(), ohi.get_pointer(), nullptr), "rmcrt_refine_Q");
  }
}

//______________________________________________________________________
// Ray.cc:1592-1615
void Ray::sched_setBoundaryConditions(const LevelP& level, SchedulerP& sched, Task::WhichDW temp_dw, const bool backoutTemp) {
  std::string taskname = "Ray::setBoundaryConditions";
  Task* tsk = nullptr;
  if (d_FLT_DBL == TypeDescription::double_type) tsk = new Task(taskname, this, &Ray::setBoundaryConditions<double>, temp_dw, backoutTemp);
  else                                           tsk = new Task(taskname, this, &Ray::setBoundaryConditions<float>, temp_dw, backoutTemp);
  const bool deviceLevel = isDeviceLevel(level.get());   // coarse Data-Onion / RMCRT_coarseLevel level: records live on the device
  if (!deviceLevel) {
    if (!backoutTemp) tsk->requiresVar(temp_dw, d_compTempLabel, Ghost::None, 0);
    tsk->modifiesVar(d_sigmaT4Label);
    tsk->modifiesVar(d_abskgLabel);
  }
  sched->addTask(tsk, level->eachPatch(), d_matlSet, RMCRTCommon::TG_RMCRT);
}

// Ray.cc:1662-1742.  Fine level: extra cells of the DataWarehouse variables abskg and sigmaT4 are set from the <BCType>
// entries (host glue over O(N^2) cells).  Coarse Data-Onion level: the same Dirichlet values go straight into the
// device records with rmcrt_set_level_extra_cells.
template <class T>
void Ray::setBoundaryConditions(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*,
                                DataWarehouse* new_dw, Task::WhichDW temp_dw, const bool backoutTemp) {
  if (d_onOff_SetBCs == false) return;
  const Level* level = getLevel(patches);
  const double sigma_over_pi = d_sigma / M_PI;

  if (isDeviceLevel(level)) {
    double abskg[6], sig[6];
    int32_t types[6];
    for (int f = 0; f < 6; ++f) {
      BCEntry a, t, c;
      if (!d_grid->getBC(f, d_abskgBC_tag, a) || !d_grid->getBC(f, d_compTempLabel->getName(), t))
        throw ProblemSetupException("Ray::setBoundaryConditions: abskg / temperature BC missing on a domain face", __FILE__, __LINE__);
      if (a.var != "Dirichlet" || t.var != "Dirichlet")
        throw ProblemSetupException("RMCRT (B200 back end): coarse-level boundary conditions must be Dirichlet", __FILE__, __LINE__);
      abskg[f] = a.value;
      sig[f] = sigma_over_pi * t.value * t.value * t.value * t.value;
      types[f] = d_grid->getBC(f, d_cellTypeLabel->getName(), c) ? (int32_t)c.value : 8;
    }
    for (auto& dev : devices())
      check(dev.ctx, rmcrt_set_level_extra_cells(dev.ctx, level->getIndex(), abskg, sig, types, nullptr), "rmcrt_set_level_extra_cells");
    return;
  }

  for (int p = 0; p < patches->size(); ++p) {
    const Patch* patch = patches->get(p);
    std::vector<Patch::FaceType> bf;
    patch->getBoundaryFaces(bf);
    if (bf.empty()) continue;
    CCVariable<double> temp;
    CCVariable<T> abskg, sigmaT4OverPi;
    new_dw->allocateTemporary(temp, patch);
    new_dw->getModifiable(abskg, d_abskgLabel, d_matl, patch);
    new_dw->getModifiable(sigmaT4OverPi, d_sigmaT4Label, d_matl, patch);
    if (backoutTemp) {        // temperature is not available on every level but sigmaT4 is (1696-1709)
      for (size_t i = 0; i < temp.cells(); ++i) temp.getPointer()[i] = std::pow((double)sigmaT4OverPi.getPointer()[i] / sigma_over_pi, 0.25);
    } else {
      DataWarehouse* t_dw = new_dw->getOtherDataWarehouse(temp_dw);
      constCCVariable<double> varTmp;
      t_dw->get(varTmp, d_compTempLabel, d_matl, patch, Ghost::None, 0);
      temp.copyData(varTmp);
    }
    setBC<T, double>(abskg, d_abskgBC_tag, patch, d_matl);
    setBC<double, double>(temp, d_compTempLabel->getName(), patch, d_matl);
    for (auto face : bf) {    // sigma T^4 / pi of the wall temperature on the extra cells incl. edges (1729-1739)
      IntVector lo, hi;
      patch->getFaceExtraCells(face, lo, hi);
      for (int k = lo.z(); k < hi.z(); ++k)
        for (int j = lo.y(); j < hi.y(); ++j)
          for (int i = lo.x(); i < hi.x(); ++i) {
            const IntVector c(i, j, k);
            const double T_sqrd = temp[c] * temp[c];
            sigmaT4OverPi[c] = (T)(sigma_over_pi * T_sqrd * T_sqrd);
          }
    }
  }
}

//______________________________________________________________________
// Ray.cc:1747-1867: Dirichlet sets the value, Neumann extrapolates with the given gradient; faces are visited in the
// order x-,x+,y-,y+,z-,z+ over extra cells plus edges, so later faces own the shared edges
template <class T, class V>
void Ray::setBC(CCVariable<T>& Q_CC, const std::string& desc, const Patch* patch, const int /*mat_id*/) {
  if (d_onOff_SetBCs == false) return;
  std::vector<Patch::FaceType> bf;
  patch->getBoundaryFaces(bf);
  const Grid* grid = patch->getLevel()->getGrid();
  const Vector dx = patch->dCell();
  for (auto face : bf) {
    BCEntry e;
    if (!grid->getBC((int)face, desc, e)) continue;
    IntVector lo, hi;
    patch->getFaceExtraCells(face, lo, hi);
    const int d = face / 2;
    const int inward = (face % 2 == 0) ? 1 : -1;
    for (int k = lo.z(); k < hi.z(); ++k)
      for (int j = lo.y(); j < hi.y(); ++j)
        for (int i = lo.x(); i < hi.x(); ++i) {
          const IntVector c(i, j, k);
          if (e.var == "Dirichlet") {
            Q_CC[c] = (T)(V)e.value;
          } else if (e.var == "Neumann") {
            IntVector in = c;
            in[d] += inward;
            if (Q_CC.contains(in)) Q_CC[c] = (T)(Q_CC[in] - (V)e.value * dx[d] * inward);
          } else {
            throw ProblemSetupException("Ray::setBC: unknown BC var '" + e.var + "' for " + desc, __FILE__, __LINE__);
          }
        }
  }
}
template void Ray::setBC<int, int>(CCVariable<int>&, const std::string&, const Patch*, const int);
template void Ray::setBC<double, double>(CCVariable<double>&, const std::string&, const Patch*, const int);
template void Ray::setBC<float, double>(CCVariable<float>&, const std::string&, const Patch*, const int);

}  // namespace Uintah
