// Radiometer.cc -- see Radiometer.h.  Reference behaviour followed: RMCRT/Radiometer.cc (line ranges cited per method).
#include "Radiometer.h"

#include <cfloat>
#include <climits>
#include <cmath>
#include <iostream>
#include <set>
#include <sstream>

namespace Uintah {

namespace {
// Core/Grid/Level-style box overlap test used by Radiometer.cc:464,582,608
inline bool doesIntersect(const IntVector& lo1, const IntVector& hi1, const IntVector& lo2, const IntVector& hi2) {
  return lo1.x() < hi2.x() && lo1.y() < hi2.y() && lo1.z() < hi2.z() && hi1.x() > lo2.x() && hi1.y() > lo2.y() && hi1.z() > lo2.z();
}
}  // namespace

//______________________________________________________________________
// Radiometer.cc:60-71
Radiometer::Radiometer(const TypeDescription::Type FLT_DBL, RMCRTCommon* backend) : RMCRTCommon(FLT_DBL) {
  d_VRFluxLabel = VarLabel::create("VRFlux", FLT_DBL);
  d_VRIntensityLabel = VarLabel::create("VRIntensity", FLT_DBL);
  std::cout << "  - Using " << (FLT_DBL == TypeDescription::double_type ? "double" : "float") << " implementation of the radiometer" << std::endl;
  shareDevicesOf(backend);
}

Radiometer::~Radiometer() {
  VarLabel::destroy(d_VRFluxLabel);
  VarLabel::destroy(d_VRIntensityLabel);
}

//______________________________________________________________________
// Radiometer.cc:91-287: one entry per <Radiometer> block, same required keys, same errors
void Radiometer::problemSetup(const ProblemSpecP&, const ProblemSpecP& rmcrtps, const GridP& grid) {
  ProblemSpecP rmcrt_ps = rmcrtps;
  for (ProblemSpecP rad_ps = rmcrt_ps->findBlock("Radiometer"); rad_ps != nullptr; rad_ps = rad_ps->findNextBlock("Radiometer")) {
    const Vector notInitialized(DBL_MAX);
    Vector unitNormalVec(notInitialized), normalVec(notInitialized);
    auto same = [](const Vector& a, const Vector& b) { return a.x() == b.x() && a.y() == b.y() && a.z() == b.z(); };
    double viewAngle;
    int nRays;
    Point start, end;
    rad_ps->require("viewAngle", viewAngle);
    rad_ps->require("nRays", nRays);
    rad_ps->require("locationsMin", start);
    rad_ps->require("locationsMax", end);
    rad_ps->get("unitNormalVector", unitNormalVec);
    rad_ps->get("normalVector", normalVec);

    BBox compDomain;
    grid->getSpatialRange(compDomain);
    const Point min = compDomain.min(), max = compDomain.max();
    if (start.x() < min.x() || start.y() < min.y() || start.z() < min.z() || end.x() > max.x() || end.y() > max.y() || end.z() > max.z()) {
      std::ostringstream warn;
      warn << "\n ERROR:Radiometer::problemSetup: the radiometer that you've specified " << start << " " << end
           << " begins or ends outside of the computational domain. \n";
      throw ProblemSetupException(warn.str(), __FILE__, __LINE__);
    }
    if (start.x() > end.x() || start.y() > end.y() || start.z() > end.z()) {
      std::ostringstream warn;
      warn << "\n ERROR:Radiometer::problemSetup: the radiometer that you've specified " << start << " " << end
           << " the starting point is > than the ending point \n";
      throw ProblemSetupException(warn.str(), __FILE__, __LINE__);
    }
    if (start.x() == end.x() || start.y() == end.y() || start.z() == end.z()) {
      std::ostringstream warn;
      warn << "\n ERROR:Radiometer::problemSetup: The specified radiometer has the same starting and ending points.  All of the directions must differ by one cell\n "
           << "                                start: " << start << " end: " << end;
      throw ProblemSetupException(warn.str(), __FILE__, __LINE__);
    }
    for (int L = 0; L < grid->numLevels(); L++) {
      const LevelP& level = grid->getLevel(L);
      const IntVector lo = level->getCellIndex(start), hi = level->getCellIndex(end), nCells = hi - lo;
      std::cout << "  - radiometer: lower cell " << lo << " upper cell " << hi << " nCells " << nCells << "\n";
      if (nCells.x() <= 0 || nCells.y() <= 0 || nCells.z() <= 0) {
        std::ostringstream warn;
        warn << "\n ERROR:Radiometer::problemSetup: The specified radiometer has the same starting and ending points.  There must be at least 1 cell between "
             << "locationMin & locationMax in each direction:\n                                 startCell: " << lo << " endCell: " << hi;
        throw ProblemSetupException(warn.str(), __FILE__, __LINE__);
      }
    }
    if (same(unitNormalVec, notInitialized) && same(normalVec, notInitialized))
      throw ProblemSetupException("\n ERROR:Radiometer::problemSetup: You must specify either:   <unitNormalVector> or <normalVector> \n to set the direction of the radiometer", __FILE__, __LINE__);
    if (viewAngle > 360) {
      std::ostringstream warn;
      warn << "ERROR:  VRViewAngle (" << viewAngle << ") exceeds the maximum acceptable value of 360 degrees.";
      throw ProblemSetupException(warn.str(), __FILE__, __LINE__);
    }
    if (nRays < int(15 + std::pow(5.4, viewAngle / 40)))
      std::cout << "    WARNING Number of radiometer rays:  (" << nRays << ") is less than the recommended number of ("
                << int(15 + std::pow(5.4, viewAngle / 40)) << "). Errors will exceed 1%. " << std::endl;

    // rotation / view-cone angles (Radiometer.cc:214-260) come from the library so that every binding derives the same numbers
    radiometer r;
    const bool isNormal = !same(normalVec, notInitialized);
    const Vector& v = isNormal ? normalVec : unitNormalVec;
    const double vec[3] = {v.x(), v.y(), v.z()};
    check(nullptr, rmcrt_radiometer_setup(viewAngle, vec, isNormal ? 1 : 0, &r.angles), "rmcrt_radiometer_setup");
    r.angles.nRays = nRays;
    r.locationsMin = start;
    r.locationsMax = end;
    std::cout << "                xi_rotate: " << r.angles.phi_rotate << " theta_rotate: " << r.angles.theta_rotate << " xi_rotate: " << r.angles.xi_rotate << std::endl;
    d_radiometers.push_back(r);
  }

  // Radiometer.cc:274-286
  ProblemSpecP root_ps = rmcrt_ps->getRootNode();
  Vector periodic(0, 0, 0);
  ProblemSpecP grid_ps = root_ps->findBlock("Grid");
  ProblemSpecP level_ps = grid_ps ? grid_ps->findBlock("Level") : nullptr;
  if (level_ps) level_ps->getWithDefault("periodic", periodic, Vector(0, 0, 0));
  if (periodic.length() != 0)
    throw ProblemSetupException("\nERROR RMCRT:\nPeriodic boundary conditions are not allowed with Radiometer.", __FILE__, __LINE__);
}

//______________________________________________________________________
// Radiometer.cc:293-310
void Radiometer::sched_initialize_VRFlux(const LevelP& level, SchedulerP& sched) {
  std::string taskname = "Radiometer::initialize_VRFlux";
  Task* tsk = nullptr;
  if (d_FLT_DBL == TypeDescription::double_type) tsk = new Task(taskname, this, &Radiometer::initialize_VRFlux<double>);
  else                                           tsk = new Task(taskname, this, &Radiometer::initialize_VRFlux<float>);
  tsk->computesVar(d_VRFluxLabel);
  tsk->computesVar(d_VRIntensityLabel);
  sched->addTask(tsk, level->eachPatch(), d_matlSet, RMCRTCommon::TG_RMCRT);
}

// Radiometer.cc:315-340
template <class T>
void Radiometer::initialize_VRFlux(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw) {
  for (int p = 0; p < patches->size(); p++) {
    const Patch* patch = patches->get(p);
    CCVariable<T> flux, intensity;
    new_dw->allocateAndPut(flux, d_VRFluxLabel, d_matl, patch);
    new_dw->allocateAndPut(intensity, d_VRIntensityLabel, d_matl, patch);
    flux.initialize((T)0.0);
    intensity.initialize((T)0.0);
  }
}

//______________________________________________________________________
// Radiometer.cc:346-398: temporal (TG_RMCRT) and spatial (only the patches that hold radiometer cells) scheduling
void Radiometer::sched_radiometer(const LevelP& level, SchedulerP& sched, Task::WhichDW /*notUsed*/, Task::WhichDW sigma_dw, Task::WhichDW celltype_dw) {
  sched_initialize_VRFlux(level, sched);   // there has to be a VRFlux / VRIntensity on every patch for output

  PatchSet radiometerPatchSet;
  getPatchSet(sched, level, radiometerPatchSet);

  const int L = level->getIndex();
  Task::WhichDW abskg_dw = get_abskg_whichDW(L, d_abskgLabel);
  std::string taskname = "Radiometer::radiometerTask";
  Task* tsk;
  if (d_FLT_DBL == TypeDescription::double_type) tsk = new Task(taskname, this, &Radiometer::radiometerTask<double>, abskg_dw, sigma_dw, celltype_dw);
  else                                           tsk = new Task(taskname, this, &Radiometer::radiometerTask<float>, abskg_dw, sigma_dw, celltype_dw);
  tsk->setType(Task::Spatial);

  Ghost::GhostType gac = Ghost::AroundCells;   // the whole level, like the ray trace (Radiometer.cc:383-385)
  tsk->requiresVar(abskg_dw, d_abskgLabel, gac, SHRT_MAX);
  tsk->requiresVar(sigma_dw, d_sigmaT4Label, gac, SHRT_MAX);
  tsk->requiresVar(celltype_dw, d_cellTypeLabel, gac, SHRT_MAX);
  tsk->modifiesVar(d_VRFluxLabel);
  tsk->modifiesVar(d_VRIntensityLabel);
  tsk->usesDevice(true);
  sched->addTask(tsk, sched->adopt(radiometerPatchSet), d_matlSet, RMCRTCommon::TG_RMCRT);
}

// Radiometer.cc:403-423
template <class T>
void Radiometer::radiometerTask(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw,
                                Task::WhichDW which_abskg_dw, Task::WhichDW which_sigmaT4_dw, Task::WhichDW which_celltype_dw) {
  if (patches->size() == 0) return;
  const Level* level = getLevel(patches);
  DataWarehouse* abskg_dw = new_dw->getOtherDataWarehouse(which_abskg_dw);
  DataWarehouse* sigmaT4_dw = new_dw->getOtherDataWarehouse(which_sigmaT4_dw);
  DataWarehouse* celltype_dw = new_dw->getOtherDataWarehouse(which_celltype_dw);
  uploadLevel<T>(level, abskg_dw, sigmaT4_dw, celltype_dw, new_dw->getID());
  for (int p = 0; p < patches->size(); p++) {
    const bool modifiesFlux = true;
    radiometerFlux<T>(patches->get(p), level, new_dw, modifiesFlux);
  }
}

//______________________________________________________________________
// Radiometer.cc:428-500
template <class T>
void Radiometer::radiometerFlux(const Patch* patch, const Level* level, DataWarehouse* new_dw, const bool modifiesFlux) {
  CCVariable<T> VRFlux, intensity;
  if (modifiesFlux) {
    new_dw->getModifiable(VRFlux, d_VRFluxLabel, d_matl, patch);
    new_dw->getModifiable(intensity, d_VRIntensityLabel, d_matl, patch);
  } else {
    new_dw->allocateAndPut(VRFlux, d_VRFluxLabel, d_matl, patch);
    new_dw->allocateAndPut(intensity, d_VRIntensityLabel, d_matl, patch);
    VRFlux.initialize((T)0.0);
    intensity.initialize((T)0.0);
  }
  const IntVector pLo = patch->getCellLowIndex(), pHi = patch->getCellHighIndex();
  std::vector<rmcrt_radiometer> rads;
  for (const radiometer& rad : d_radiometers) {
    const IntVector rLo = level->getCellIndex(rad.locationsMin), rHi = level->getCellIndex(rad.locationsMax);
    if (!doesIntersect(rLo, rHi, pLo, pHi)) continue;
    rmcrt_radiometer r = rad.angles;
    for (int d = 0; d < 3; ++d) { r.cellLo[d] = rLo[d]; r.cellHi[d] = rHi[d]; }
    rads.push_back(r);
  }
  if (rads.empty()) return;
  // patch -> device as in the ray trace (GPUGridVariableInfo.cc:610-632 deals patches round robin); radiometer patches are few
  auto& devs = devices();
  Device& dev = devs[(size_t)patch->getID() % devs.size()];
  check(dev.ctx, rmcrt_finalize_level(dev.ctx, level->getIndex(), nullptr), "rmcrt_finalize_level");
  const IntVector olo = VRFlux.getLowIndex(), ohi = VRFlux.getHighIndex();
  check(dev.ctx, rmcrt_trace_radiometer(dev.ctx, level->getIndex(), rads.data(), (int)rads.size(), pLo.get_pointer(), pHi.get_pointer(),
                                        VRFlux.getPointer(), intensity.getPointer(), sizeof(T) == 4 ? 1 : 0, olo.get_pointer(), ohi.get_pointer(),
                                        /*zeroInit*/ 0, (uint64_t)new_dw->getTimeStep(), nullptr), "rmcrt_trace_radiometer");
}
template void Radiometer::radiometerFlux<double>(const Patch*, const Level*, DataWarehouse*, const bool);
template void Radiometer::radiometerFlux<float>(const Patch*, const Level*, DataWarehouse*, const bool);

//______________________________________________________________________
// Radiometer.cc:565-620: the patches that contain radiometer cells
void Radiometer::getPatchSet(SchedulerP&, const LevelP& level, PatchSet& ps) {
  IntVector L_lo, L_hi;
  level->findInteriorCellIndexRange(L_lo, L_hi);
  std::set<int> chosen;
  for (const radiometer& rad : d_radiometers) {
    const IntVector rLo = level->getCellIndex(rad.locationsMin), rHi = level->getCellIndex(rad.locationsMax);
    if (!doesIntersect(rLo, rHi, L_lo, L_hi)) {
      std::ostringstream warn;
      warn << "\n ERROR:Radiometer::problemSetup: The VR locations:\n     VR min " << rad.locationsMin << "  cell index: " << rLo
           << "\n     VR max " << rad.locationsMax << "  cell index: " << rHi << "\n     Number of cells: " << rHi - rLo
           << "\n do not overlap with the domain, " << L_lo << " -> " << L_hi << "\n There must be at least 1 cell in each direction";
      throw ProblemSetupException(warn.str(), __FILE__, __LINE__);
    }
    for (int p = 0; p < level->numPatches(); ++p) {
      const Patch* patch = level->getPatch(p);
      if (doesIntersect(rLo, rHi, patch->getCellLowIndex(), patch->getCellHighIndex())) chosen.insert(p);
    }
  }
  for (int p : chosen) ps.add(level->getPatch(p));
}

}  // namespace Uintah
