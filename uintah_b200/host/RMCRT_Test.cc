// RMCRT_Test.cc -- stand-alone driver of the host plugin, mirroring the reference's example component
// CCA/Components/Examples/RMCRT_Test.{h,cc} (problemSetup 111-260, scheduleInitialize 317-341, scheduleTimeAdvance 365-600,
// initialize 619-821, sched_initProperties 944-958, task-graph choice 989-994) on top of UintahShim.h.
//
//   rmcrt_test <deck.ups> [--dump DIR] [--timesteps N] [--devices N] [--seed S] [--knob name=value]... [--parse-only]
//
// --parse-only stops after problemSetup + scheduling and prints every registered task with its dependencies (no GPU needed).
//
// It reads the same .ups decks (the <Grid>, <Temperature>/<abskg>/<benchmark>/<calc_frequency> keys and the <RMCRT> block),
// builds the grid, runs the time loop with the two task graphs (TG_CARRY_FORWARD / TG_RMCRT) and, with --dump, writes the
// finest level's divQ / radiationVolq / RMCRTboundFlux / abskg / sigmaT4 / cellType as raw arrays for the parity tests.
#include "DataArchive.h"
#include "DriverCommon.h"
#include "Ray.h"
#include "WasatchRadiationSource.h"

#include <sys/stat.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

using namespace Uintah;


namespace {

class RMCRT_Test {
 public:
  enum Algo { dataOnion, dataOnionSlim, coarseLevel, singleLevel, radiometerOnly };   // RMCRT_Test.h:166-171

  RMCRT_Test() {
    d_colorLabel = VarLabel::create("color", TypeDescription::double_type);
    d_compAbskgLabel = VarLabel::create("abskg", TypeDescription::double_type);
    d_divQLabel = VarLabel::create("divQ", TypeDescription::double_type);
    d_cellTypeLabel = VarLabel::create("cellType", TypeDescription::int_type);
  }
  ~RMCRT_Test() {
    d_wasatch.reset();
    d_RMCRT.reset();
    VarLabel::destroy(d_colorLabel);
    VarLabel::destroy(d_compAbskgLabel);
    VarLabel::destroy(d_divQLabel);
    VarLabel::destroy(d_cellTypeLabel);
  }

  //______________________________________________________________________ RMCRT_Test.cc:111-260
  void problemSetup(const ProblemSpecP& prob_spec, GridP& grid, SchedulerP& sched) {
    d_grid = grid;
    sched->setNumTaskGraphs(RMCRTCommon::NUM_GRAPHS);   // TG-0 carry forward, TG-1 RMCRT
    prob_spec->getWithDefault("calc_frequency", d_radCalc_freq, 1);
    prob_spec->require("Temperature", d_initColor);
    prob_spec->require("abskg", d_initAbskg);
    prob_spec->getWithDefault("benchmark", d_benchmark, 0);
    if (d_benchmark > 8 || d_benchmark < 0) {
      std::ostringstream warn;
      warn << "RMCRT:ERROR:  Benchmark (" << d_benchmark << ") only valid options are 1-7\n";
      throw ProblemSetupException(warn.str(), __FILE__, __LINE__);
    }
    if (ProblemSpecP bm_ps = prob_spec->findBlock("benchmark")) {        // 134-137: <benchmark dir="0|1|2">
      std::string dir;
      if (bm_ps->getAttribute("dir", dir)) d_pDir = std::stoi(dir);
    }
    if ((d_benchmark == 6 || d_benchmark == 7) && (d_pDir > 2 || d_pDir < 0)) {
      std::ostringstream warn;
      warn << "RMCRT:ERROR:  Benchmark (" << d_benchmark << ") requires the attribute 'dir' to be specified (0,1,2)\n";
      throw ProblemSetupException(warn.str(), __FILE__, __LINE__);
    }
    if (d_benchmark == 4 || d_benchmark == 5)     // Siegel isotropic scattering: needs a RAY_SCATTER build (Ray.cc:196-201)
      throw ProblemSetupException("rmcrt_test: benchmarks 4 and 5 need isotropic scattering, which this build does not have", __FILE__, __LINE__);
    // <Intrusion><geom_object> box | cylinder | sphere (221-225; GeometryPieceFactory::create)
    if (ProblemSpecP intr = prob_spec->findBlock("Intrusion"))
      if (ProblemSpecP go = intr->findBlock("geom_object"))
        for (const auto& c : go->children()) {
          Intrusion g;
          const std::string& nm = c->getNodeName();
          if (nm == "box") { g.kind = 0; c->require("min", g.a); c->require("max", g.b); }
          else if (nm == "cylinder") { g.kind = 1; c->require("bottom", g.a); c->require("top", g.b); c->require("radius", g.r); }
          else if (nm == "sphere") { g.kind = 2; c->require("origin", g.a); c->require("radius", g.r); }
          else throw ProblemSetupException("rmcrt_test: <Intrusion> geometry piece <" + nm + "> is not supported (box, cylinder, sphere)", __FILE__, __LINE__);
          d_intrusions.push_back(g);
        }

    // <initalizeUsingUda> (sic; 211-219) + its bullet proofing (253-310)
    if (ProblemSpecP uda_ps = prob_spec->findBlock("initalizeUsingUda")) {
      d_old_uda.reset(new useOldUdaData());
      uda_ps->get("uda_name", d_old_uda->udaName);
      uda_ps->get("timestep", d_old_uda->timestep);
      uda_ps->get("abskg_varName", d_old_uda->abskgName);
      uda_ps->get("temperature_varName", d_old_uda->temperatureName);
      uda_ps->getWithDefault("matl_index", d_old_uda->matl, 0);
      uda_ps->getWithDefault("cellType_varName", d_old_uda->cellTypeName, "NONE");
      if (!d_udaDirOverride.empty()) d_old_uda->udaName = d_udaDirOverride;
      DataArchive archive(d_old_uda->udaName);
      std::vector<int> index;
      std::vector<double> times;
      archive.queryTimesteps(index, times);
      d_old_uda->timeIndex = -9;
      for (unsigned int i = 0; i < index.size(); i++)
        if (d_old_uda->timestep == index[i]) d_old_uda->timeIndex = (int)i;
      if (d_old_uda->timeIndex < 0) {
        std::ostringstream warn;
        warn << "The timestep (" << d_old_uda->timestep << ") was not found in the uda\nThere are " << index.size() << " timesteps\n";
        throw ProblemSetupException(warn.str(), __FILE__, __LINE__);
      }
      areGridsEqual(archive.queryGrid(d_old_uda->timeIndex), grid);
      std::vector<std::string> vars;
      std::vector<int> num_matls;
      std::vector<TypeDescription::Type> types;
      archive.queryVariables(vars, num_matls, types);
      int vars_found = 0;
      for (unsigned int i = 0; i < vars.size(); i++) {
        if (d_old_uda->abskgName == vars[i]) { vars_found += 1; d_old_uda->subType = types[i]; }
        else if (d_old_uda->temperatureName == vars[i]) { vars_found += 1; d_old_uda->subType = types[i]; }
      }
      if (vars_found != 2 && vars_found != 3) {
        std::ostringstream warn;
        warn << "The variables (" << d_old_uda->abskgName << "), (" << d_old_uda->temperatureName << "), or optional variale cellType: ("
             << d_old_uda->cellTypeName << ") was not found in the uda";
        throw ProblemSetupException(warn.str(), __FILE__, __LINE__);
      }
    }

    ProblemSpecP rmcrt_ps = prob_spec->findBlock("RMCRT");
    if (!rmcrt_ps) throw ProblemSetupException("no <RMCRT> block", __FILE__, __LINE__);
    std::map<std::string, std::string> type;
    rmcrt_ps->getAttributes(type);
    d_RMCRT.reset(new Ray(type["type"] == "float" ? TypeDescription::float_type : TypeDescription::double_type));

    d_whichAlgo = singleLevel;
    if (ProblemSpecP alg_ps = rmcrt_ps->findBlock("algorithm")) {
      std::string t = "nullptr";
      alg_ps->getAttribute("type", t);
      if (t == "dataOnion") d_whichAlgo = dataOnion;
      else if (t == "dataOnionSlim") d_whichAlgo = dataOnionSlim;
      else if (t == "RMCRT_coarseLevel") d_whichAlgo = coarseLevel;
      else if (t == "singleLevel") d_whichAlgo = singleLevel;
      else if (t == "radiometerOnly") d_whichAlgo = radiometerOnly;
      if (d_whichAlgo == dataOnion) {
        ProblemSpecP amr = prob_spec->findBlock("AMR");
        bool lockstep = false;
        if (amr) amr->get("useLockStep", lockstep);
        if (!lockstep) throw ProblemSetupException("\n ERROR: You must add \n <useLockStep> true </useLockStep> \n inside of the <AMR> section. \n", __FILE__, __LINE__);
      }
    }
    // the caller-private enum is passed as a bare int; Ray only compares it with radiometerOnly
    int ray_algo = RMCRTCommon::singleLevel;
    if (d_whichAlgo == dataOnion || d_whichAlgo == dataOnionSlim) ray_algo = RMCRTCommon::dataOnion;
    else if (d_whichAlgo == coarseLevel) ray_algo = RMCRTCommon::coarseLevel;
    else if (d_whichAlgo == radiometerOnly) ray_algo = RMCRTCommon::radiometerOnly;
    if (d_wasatchCaller) {   // the Wasatch consumer: RadiationSource's constructor registers the variables and runs problemSetup itself
      if (d_whichAlgo != singleLevel) throw ProblemSetupException("rmcrt_test --caller wasatch: Wasatch drives RMCRT on a single level", __FILE__, __LINE__);
      d_wasatch.reset(new WasatchCore::RadiationSource("divQ", "color", "abskg", "cellType", d_RMCRT.get(), rmcrt_ps, grid));
    } else {
      d_RMCRT->registerVariables(0, d_compAbskgLabel, d_colorLabel, d_cellTypeLabel, d_divQLabel, ray_algo);
      d_RMCRT->problemSetup(prob_spec, rmcrt_ps, grid);
      d_RMCRT->BC_bulletproofing(rmcrt_ps, true, true);
    }

    // one extra cell (or periodicity) in each direction (RMCRT_Test.cc:238-251)
    const IntVector extraCells = grid->getLevel(0)->getExtraCells();
    for (int dir = 0; dir < 3; ++dir)
      if (grid->periodic[dir] != 1 && extraCells[dir] != 1) {
        std::ostringstream warn;
        warn << "\n \n INPUT FILE ERROR: \n You must have either a periodic boundary " << grid->periodic << " or one extraCell "
             << extraCells << " specified in each direction";
        throw ProblemSetupException(warn.str(), __FILE__, __LINE__);
      }
  }

  //______________________________________________________________________ RMCRT_Test.cc:317-341
  void scheduleInitialize(const LevelP& level, SchedulerP& sched) {
    Task* task = !d_old_uda ? new Task("RMCRT_Test::initialize", this, &RMCRT_Test::initialize)
                            : new Task("RMCRT_Test::initializeWithUda", this, &RMCRT_Test::initializeWithUda);   // 321-328
    task->computesVar(d_colorLabel);
    task->computesVar(d_compAbskgLabel);
    task->computesVar(d_cellTypeLabel);
    sched->addTask(task, level->eachPatch(), RMCRTCommon::d_matlSet);
    if (Radiometer* radiometer = d_RMCRT->getRadiometer()) radiometer->sched_initialize_VRFlux(level, sched);   // 337-340
  }

  //______________________________________________________________________ RMCRT_Test.cc:365-600
  void scheduleTimeAdvance(const LevelP& level, SchedulerP& sched) {
    if (level->getIndex() > 0) return;   // only schedule once
    GridP grid = d_grid;
    const int maxLevels = grid->numLevels();
    const LevelP& finestLevel = grid->getLevel(maxLevels - 1);
    sched_initProperties(finestLevel, sched);

    std::vector<const VarLabel*> fineLevelVarLabels = {RMCRTCommon::d_divQLabel, RMCRTCommon::d_boundFluxLabel,
                                                       RMCRTCommon::d_radiationVolqLabel, RMCRTCommon::d_sigmaT4Label};
    const bool includeExtraCells = true;
    Task::WhichDW notUsed = Task::None;
    Radiometer* radiometer = d_RMCRT->getRadiometer();                    // 384

    if (d_wasatchCaller) {     // RadiationSource::schedule_ray_tracing (Wasatch/Expressions/RadiationSource.cc:141-172): modifies_divQ = true
      // stands in for the expression tree, which owns the divQ field: allocate it (and radiationVolq) filled with a sentinel so the test
      // can see that the trace leaves every cell it does not compute untouched
      Task* alloc = new Task("Wasatch::allocate_fields", this, &RMCRT_Test::wasatchAllocate);
      alloc->computesVar(d_divQLabel);
      alloc->computesVar(RMCRTCommon::d_radiationVolqLabel);
      sched->addTask(alloc, level->eachPatch(), RMCRTCommon::d_matlSet, RMCRTCommon::TG_RMCRT);
      d_wasatch->schedule_ray_tracing(level, sched, RMCRTCommon::d_matlSet, 0);
      return;
    }

    if (d_whichAlgo == dataOnion || d_whichAlgo == dataOnionSlim) {      // 385-460
      Task::WhichDW temp_dw = Task::NewDW, sigmaT4_dw = Task::NewDW, celltype_dw = Task::NewDW;
      const bool backoutTemp = true, modifies_abskg = false, modifies_sigmaT4 = false, modifies_divQ = false;
      const LevelP& fineLevel = finestLevel;
      d_RMCRT->set_abskg_dw_perLevel(fineLevel, Task::NewDW);
      d_RMCRT->sched_DoubleToFloat(fineLevel, sched, notUsed);
      d_RMCRT->sched_sigmaT4(fineLevel, sched, temp_dw, includeExtraCells);
      d_RMCRT->sched_setBoundaryConditions(fineLevel, sched, temp_dw);
      d_RMCRT->sched_carryForward_VarLabels(fineLevel, sched, fineLevelVarLabels);
      // coarse-level abskg/sigmaT4 live on the device only (Ray::coarsenAll), so there is nothing to carry forward there
      for (int l = maxLevels - 2; l >= 0; l--) {
        const LevelP& lev = grid->getLevel(l);
        d_RMCRT->sched_CoarsenAll(lev, sched, modifies_abskg, modifies_sigmaT4);
        if (d_RMCRT->d_coarsenExtraCells == false) d_RMCRT->sched_setBoundaryConditions(lev, sched, notUsed, backoutTemp);
      }
      d_RMCRT->sched_ROI_Extents(fineLevel, sched);
      d_RMCRT->sched_rayTrace_dataOnion(fineLevel, sched, notUsed, sigmaT4_dw, celltype_dw, modifies_divQ);
    }

    if (d_whichAlgo == singleLevel) {                                     // 527-557
      Task::WhichDW temp_dw = Task::NewDW, sigmaT4_dw = Task::NewDW, celltype_dw = Task::NewDW;
      const bool modifies_divQ = false, backoutTemp = true;
      d_RMCRT->set_abskg_dw_perLevel(level, Task::NewDW);
      d_RMCRT->sched_carryForward_VarLabels(level, sched, fineLevelVarLabels);
      d_RMCRT->sched_DoubleToFloat(level, sched, notUsed);
      d_RMCRT->sched_sigmaT4(level, sched, temp_dw, includeExtraCells);
      d_RMCRT->sched_setBoundaryConditions(level, sched, temp_dw, backoutTemp);
      d_RMCRT->sched_rayTrace(level, sched, notUsed, sigmaT4_dw, celltype_dw, modifies_divQ);
      if (radiometer) {                                                   // 549-553
        d_RMCRT->sched_CarryForward_Var(level, sched, radiometer->d_VRFluxLabel, RMCRTCommon::TG_CARRY_FORWARD);
        radiometer->sched_radiometer(level, sched, notUsed, sigmaT4_dw, celltype_dw);
      }
    }
    if (d_whichAlgo == radiometerOnly) {                                  // 559-583: no other calculations
      if (!radiometer) throw ProblemSetupException("RMCRT_Test: algorithm radiometerOnly needs a <Radiometer> block", __FILE__, __LINE__);
      Task::WhichDW temp_dw = Task::NewDW, sigmaT4_dw = Task::NewDW, celltype_dw = Task::NewDW;
      const bool backoutTemp = true;
      d_RMCRT->set_abskg_dw_perLevel(level, Task::NewDW);
      d_RMCRT->sched_carryForward_VarLabels(level, sched, {RMCRTCommon::d_sigmaT4Label, radiometer->d_VRFluxLabel, radiometer->d_VRIntensityLabel});
      d_RMCRT->sched_DoubleToFloat(level, sched, notUsed);
      d_RMCRT->sched_sigmaT4(level, sched, temp_dw, includeExtraCells);
      d_RMCRT->sched_setBoundaryConditions(level, sched, temp_dw, backoutTemp);
      radiometer->sched_radiometer(level, sched, notUsed, sigmaT4_dw, celltype_dw);
    }
    if (d_whichAlgo == coarseLevel) {                                     // 464-519: RMCRT on the coarse level, divQ interpolated to the fine one
      Task::WhichDW temp_dw = Task::NewDW, sigmaT4_dw = Task::NewDW, celltype_dw = Task::NewDW;
      const bool modifies_divQ = false, backoutTemp = true, modifies_abskg = false, modifies_sigmaT4 = false;
      // coarse levels: sigmaT4 / abskg live on the device only (Ray::coarsenAll), so only the results are carried forward there
      for (int l = 0; l < maxLevels; l++)
        d_RMCRT->sched_carryForward_VarLabels(grid->getLevel(l), sched, l == maxLevels - 1 ? fineLevelVarLabels
                                              : std::vector<const VarLabel*>{RMCRTCommon::d_divQLabel, RMCRTCommon::d_boundFluxLabel, RMCRTCommon::d_radiationVolqLabel});
      const LevelP& fineLevel = finestLevel;
      d_RMCRT->set_abskg_dw_perLevel(fineLevel, Task::NewDW);
      d_RMCRT->sched_DoubleToFloat(fineLevel, sched, notUsed);
      d_RMCRT->sched_sigmaT4(fineLevel, sched, temp_dw, includeExtraCells);
      // the reference leaves the fine level's extra cells as initialised (no setBoundaryConditions there); they are never read:
      // coarsen_Q averages interior cells only
      for (int l = maxLevels - 2; l >= 0; l--) {                          // coarsest last: each level is coarsened from the next finer one
        const LevelP& lev = grid->getLevel(l);
        d_RMCRT->sched_CoarsenAll(lev, sched, modifies_abskg, modifies_sigmaT4);
        if (d_RMCRT->d_coarsenExtraCells == false) d_RMCRT->sched_setBoundaryConditions(lev, sched, temp_dw, backoutTemp);
      }
      for (int l = 0; l < maxLevels - 1; l++)
        d_RMCRT->sched_rayTrace(grid->getLevel(l), sched, notUsed, sigmaT4_dw, celltype_dw, modifies_divQ);
      for (int l = 1; l < maxLevels; l++) {                               // 512-518
        const LevelP& lev = grid->getLevel(l);
        d_RMCRT->sched_Refine_Q(sched, lev->eachPatch(), RMCRTCommon::d_matlSet);
      }
    }
  }

  // RMCRT_Test.cc:944-958: carry abskg, color, cellType Old -> New to mimic what a component would compute
  void sched_initProperties(const LevelP& level, SchedulerP& sched) {
    Task* tsk = new Task("RMCRT_Test::initProperties", this, &RMCRT_Test::initProperties);
    tsk->requiresVar(Task::OldDW, d_colorLabel);
    tsk->requiresVar(Task::OldDW, d_compAbskgLabel);
    tsk->requiresVar(Task::OldDW, d_cellTypeLabel);
    tsk->computesVar(d_colorLabel);
    tsk->computesVar(d_compAbskgLabel);
    tsk->computesVar(d_cellTypeLabel);
    sched->addTask(tsk, level->eachPatch(), RMCRTCommon::d_matlSet, Scheduler::ALL_GRAPHS);
  }
  void initProperties(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset* matls, DataWarehouse* old_dw, DataWarehouse* new_dw) {
    new_dw->transferFrom(old_dw, d_colorLabel, patches, matls);
    new_dw->transferFrom(old_dw, d_compAbskgLabel, patches, matls);
    new_dw->transferFrom(old_dw, d_cellTypeLabel, patches, matls);
  }

  //______________________________________________________________________ RMCRT_Test.cc:619-821
  void initialize(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw) {
    const Level* level = getLevel(patches);
    for (int p = 0; p < patches->size(); p++) {
      const Patch* patch = patches->get(p);
      const int matl = 0;
      CCVariable<double> color, abskg;
      CCVariable<int> cellType;
      new_dw->allocateAndPut(color, d_colorLabel, matl, patch);
      new_dw->allocateAndPut(abskg, d_compAbskgLabel, matl, patch);
      new_dw->allocateAndPut(cellType, d_cellTypeLabel, matl, patch);
      color.initialize(d_initColor);
      abskg.initialize(d_initAbskg);
      const IntVector lo = patch->getCellLowIndex(), hi = patch->getCellHighIndex();

      if (d_benchmark == 1 || d_benchmark == 3) {      // Burns & Christon (653-680)
        IntVector pLow, pHigh;
        level->findInteriorCellIndexRange(pLow, pHigh);
        const IntVector Nx = pHigh - pLow;
        const Vector Dx = patch->dCell();
        for (int k = lo.z(); k < hi.z(); ++k)
          for (int j = lo.y(); j < hi.y(); ++j)
            for (int i = lo.x(); i < hi.x(); ++i) {
              const IntVector c(i, j, k);
              abskg[c] = 0.90 * (1.0 - 2.0 * std::fabs((c[0] - (Nx.x() - 1.0) / 2.0) * Dx.x()))
                              * (1.0 - 2.0 * std::fabs((c[1] - (Nx.y() - 1.0) / 2.0) * Dx.y()))
                              * (1.0 - 2.0 * std::fabs((c[2] - (Nx.z() - 1.0) / 2.0) * Dx.z())) + 0.1;
            }
      }
      if (d_benchmark == 2) abskg.initialize(1);       // 682-684
      if (d_benchmark == 3)                            // 686-691
        for (int k = lo.z(); k < hi.z(); ++k)
          for (int j = lo.y(); j < hi.y(); ++j)
            for (int i = lo.x(); i < hi.x(); ++i) { const IntVector c(i, j, k); color[c] = 1000 * abskg[c]; }

      if (d_benchmark == 6 || d_benchmark == 7 || d_benchmark == 8) {      // 710-779: fields given at the cell centres
        const double twoPI = 2 * M_PI;
        if (d_benchmark == 6) abskg.initialize(5.0);
        for (int k = lo.z(); k < hi.z(); ++k)
          for (int j = lo.y(); j < hi.y(); ++j)
            for (int i = lo.x(); i < hi.x(); ++i) {
              const IntVector c(i, j, k);
              const Point p = level->getCellPosition(c);
              if (d_benchmark == 6) {             // sinusoidal temperature, abskg = 5
                color[c] = std::pow(1e7 * (1.0 + std::sin(4.5 * twoPI * p[d_pDir])), 0.25);
              } else if (d_benchmark == 7) {      // abskg = 5 exp(-2x-0.1), T^4 linear in x
                const double x = p[d_pDir];
                abskg[c] = 5.0 * std::exp(-2 * x - 0.1);
                color[c] = std::pow(1.0e7 * (2.0 * x + 1.0), 0.25);
              } else {                            // hot and cold clouds
                const double x = p[0], y = p[1], z = p[2];
                abskg[c] = (std::max(std::sin(x * 5 * M_PI) * std::abs((std::sin(z * 3 * M_PI + M_PI)) + (std::sin(y * 3 * M_PI))), 0.0)
                            + std::max(std::sin(x * 5.0 * M_PI + M_PI), 0.0) * std::max(std::sin(z * 3 * M_PI + M_PI) + std::sin(y * 3 * M_PI + M_PI), 0.0)) * 5.0;
                color[c] = (std::max(std::sin(x * 5.0 * M_PI) * std::max(std::sin(z * 3.0 * M_PI) + std::sin(y * 3.0 * M_PI), 0.0), 0.0)
                            + std::max(std::sin(x * 5.0 * M_PI), 0.0) * std::max(std::sin(z * 3.0 * M_PI + M_PI) + std::sin(y * 3.0 * M_PI + M_PI), 0.0)) * 1000.0;
              }
            }
      }

      // cellType: flow everywhere, intrusions (788-811: cellType only, abskg and temperature keep their values), then the
      // boundary conditions on the extra cells (776-786, 813)
      cellType.initialize(d_RMCRT->d_flowCell);
      if (!d_intrusions.empty()) {
        const IntVector elo = patch->getExtraCellLowIndex(), ehi = patch->getExtraCellHighIndex();
        for (const Intrusion& g : d_intrusions)
          for (int k = elo.z(); k < ehi.z(); ++k)
            for (int j = elo.y(); j < ehi.y(); ++j)
              for (int i = elo.x(); i < ehi.x(); ++i) {
                const IntVector c(i, j, k);
                if (g.inside(level->getCellPosition(c))) cellType[c] = d_wall_cell;
              }
      }
      d_RMCRT->setBC<int, int>(cellType, d_cellTypeLabel->getName(), patch, matl);
      d_RMCRT->setBC<double, double>(color, d_colorLabel->getName(), patch, matl);
      d_RMCRT->setBC<double, double>(abskg, d_compAbskgLabel->getName(), patch, matl);
    }
  }

  //______________________________________________________________________ RMCRT_Test.cc:824-940
  // initialise abskg / temperature (/ cellType) from a previously written uda; the archive may hold floats or doubles
  void initializeWithUda(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw) {
    DataArchive archive(d_old_uda->udaName);
    const int timeIndex = d_old_uda->timeIndex;
    GridP uda_grid = archive.queryGrid(timeIndex);
    const Level* uda_level = uda_grid->getLevel(0).get();   // there's only one level in these problems
    const int uda_matl = d_old_uda->matl, matl = 0;
    for (int p = 0; p < patches->size(); p++) {
      const Patch* patch = patches->get(p);
      const IntVector low = patch->getExtraCellLowIndex(), high = patch->getExtraCellHighIndex();
      CCVariable<int> cellType, uda_cellType;
      new_dw->allocateAndPut(cellType, d_cellTypeLabel, matl, patch);
      cellType.initialize(d_RMCRT->d_flowCell);
      if (d_old_uda->cellTypeName != "NONE") {
        archive.queryRegion(uda_cellType, d_old_uda->cellTypeName, uda_matl, uda_level, timeIndex, low, high);
        cellType.copyData(uda_cellType);
      }
      CCVariable<double> color, abskg;
      new_dw->allocateAndPut(color, d_colorLabel, matl, patch);
      new_dw->allocateAndPut(abskg, d_compAbskgLabel, matl, patch);
      if (d_old_uda->subType == TypeDescription::double_type) {
        CCVariable<double> uda_temp, uda_abskg;
        archive.queryRegion(uda_temp, d_old_uda->temperatureName, uda_matl, uda_level, timeIndex, low, high);
        archive.queryRegion(uda_abskg, d_old_uda->abskgName, uda_matl, uda_level, timeIndex, low, high);
        for (size_t i = 0; i < color.cells(); ++i) { color.getPointer()[i] = uda_temp.getPointer()[i]; abskg.getPointer()[i] = uda_abskg.getPointer()[i]; }
      } else {
        CCVariable<float> uda_temp, uda_abskg;
        archive.queryRegion(uda_temp, d_old_uda->temperatureName, uda_matl, uda_level, timeIndex, low, high);
        archive.queryRegion(uda_abskg, d_old_uda->abskgName, uda_matl, uda_level, timeIndex, low, high);
        for (size_t i = 0; i < color.cells(); ++i) { color.getPointer()[i] = uda_temp.getPointer()[i]; abskg.getPointer()[i] = uda_abskg.getPointer()[i]; }
      }
      d_RMCRT->setBC<double, double>(color, d_colorLabel->getName(), patch, matl);
      d_RMCRT->setBC<double, double>(abskg, d_compAbskgLabel->getName(), patch, matl);
    }
  }

  // RMCRT_Test.cc:1001-1043: the finest level of the input file against the (single) level of the initialisation uda
  static void areGridsEqual(const GridP& uda_grid, const GridP& grid) {
    bool areEqual = true;
    const LevelP level = grid->getLevel(grid->numLevels() - 1);
    const LevelP otherlevel = uda_grid->getLevel(0);
    if (level->numPatches() != otherlevel->numPatches()) areEqual = false;
    for (int p = 0; areEqual && p < level->numPatches(); ++p) {
      const Patch* patch = level->getPatch(p);
      const Patch* otherpatch = otherlevel->getPatch(p);
      const IntVector lo = patch->getCellLowIndex(), hi = patch->getCellHighIndex();
      const IntVector o_lo = otherpatch->getCellLowIndex(), o_hi = otherpatch->getCellHighIndex();
      if (lo != o_lo || hi != o_hi) areEqual = false;
      const Point a = level->getCellPosition(lo), b = otherlevel->getCellPosition(o_lo), c = level->getCellPosition(hi), d = otherlevel->getCellPosition(o_hi);
      for (int k = 0; k < 3; ++k)
        if (std::fabs(a[k] - b[k]) > 1e-12 * (1 + std::fabs(a[k])) || std::fabs(c[k] - d[k]) > 1e-12 * (1 + std::fabs(c[k]))) areEqual = false;
    }
    if (!areEqual)
      throw ProblemSetupException("\n\nERROR initalizeUsingUda: The grid defined in the input file is not equal to the initialization uda's grid\n"
                                  " For the Data Onion algorithm the finest level in the input file must be equal", __FILE__, __LINE__);
  }

  void wasatchAllocate(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw) {
    for (int p = 0; p < patches->size(); p++) {
      CCVariable<double> divQ, volq;
      new_dw->allocateAndPut(divQ, d_divQLabel, 0, patches->get(p));
      new_dw->allocateAndPut(volq, RMCRTCommon::d_radiationVolqLabel, 0, patches->get(p));
      divQ.initialize(123.0);
      volq.initialize(-7.0);
    }
  }
  bool d_wasatchCaller = false;      // --caller wasatch
  std::unique_ptr<WasatchCore::RadiationSource> d_wasatch;

  struct useOldUdaData {   // RMCRT_Test.h:174-183
    std::string udaName, cellTypeName, abskgName, temperatureName;
    int timestep = 0, matl = 0, timeIndex = -9;
    TypeDescription::Type subType = TypeDescription::double_type;
  };
  std::unique_ptr<useOldUdaData> d_old_uda;
  std::string d_udaDirOverride;     // --init-uda DIR: use this archive instead of <uda_name> (fixtures live at test-time paths)

  // RMCRT_Test.cc:989-994 decides the task graph of the NEXT step; equivalently, for step ts:
  int taskGraphFor(int ts) const { return (ts % d_radCalc_freq == 0 || ts == 1) ? RMCRTCommon::TG_RMCRT : RMCRTCommon::TG_CARRY_FORWARD; }

  std::unique_ptr<Ray> d_RMCRT;
  const VarLabel *d_colorLabel, *d_compAbskgLabel, *d_divQLabel, *d_cellTypeLabel;
  GridP d_grid;
  double d_initColor = 0, d_initAbskg = 0;
  int d_radCalc_freq = 1, d_benchmark = 0, d_whichAlgo = singleLevel;
  int d_pDir = -9;                 // RMCRT_Test.cc:82
  static constexpr int d_wall_cell = 8;   // RMCRT_Test.cc:80
  // the three GeometryPiece kinds the RMCRT decks use; inside() follows Core/GeometryPiece/{Box,Cylinder,Sphere}GeometryPiece.cc
  struct Intrusion {
    int kind = 0; Vector a, b; double r = 0;
    bool inside(const Point& p) const {
      if (kind == 0) { for (int d = 0; d < 3; ++d) if (p[d] < a[d] || p[d] > b[d]) return false; return true; }   // Box:91-100
      if (kind == 1) {                                                                                          // Cylinder:126-143
        const double ax[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, tb[3] = {p[0] - a[0], p[1] - a[1], p[2] - a[2]};
        const double h2 = ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2];
        const double h = tb[0] * ax[0] + tb[1] * ax[1] + tb[2] * ax[2];
        if (h < 0.0 || h > h2) return false;
        const double cx = ax[1] * tb[2] - ax[2] * tb[1], cy = ax[2] * tb[0] - ax[0] * tb[2], cz = ax[0] * tb[1] - ax[1] * tb[0];
        return (cx * cx + cy * cy + cz * cz) / h2 <= r * r;
      }
      const double dx = p[0] - a[0], dy = p[1] - a[1], dz = p[2] - a[2];                                         // Sphere:83-92
      return dx * dx + dy * dy + dz * dz <= r * r;
    }
  };
  std::vector<Intrusion> d_intrusions;
};

}  // namespace

int main(int argc, char** argv) {
  std::string deck, dumpDir, udaOut, initUda, caller = "rmcrt_test";
  int timesteps = -1, devices = 0;
  unsigned long long seed = 0;
  bool haveSeed = false, parseOnly = false;
  std::vector<std::pair<std::string, int>> knobs;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> std::string { if (i + 1 >= argc) { std::cerr << "missing value after " << a << "\n"; std::exit(2); } return argv[++i]; };
    if (a == "--parse-only") parseOnly = true;
    else if (a == "--dump") dumpDir = next();
    else if (a == "--uda") udaOut = next();
    else if (a == "--init-uda") initUda = next();
    else if (a == "--caller") caller = next();
    else if (a == "--timesteps") timesteps = std::atoi(next().c_str());
    else if (a == "--devices") devices = std::atoi(next().c_str());
    else if (a == "--seed") { seed = std::strtoull(next().c_str(), nullptr, 0); haveSeed = true; }
    else if (a == "--knob") {
      const std::string kv = next();
      const size_t eq = kv.find('=');
      if (eq == std::string::npos) { std::cerr << "--knob needs name=value\n"; return 2; }
      knobs.emplace_back(kv.substr(0, eq), std::atoi(kv.substr(eq + 1).c_str()));
    } else if (a[0] == '-') { std::cerr << "unknown option " << a << "\n"; return 2; }
    else deck = a;
  }
  if (deck.empty()) {
    std::cerr << "usage: rmcrt_test <deck.ups> [--dump DIR] [--uda DIR] [--init-uda DIR] [--caller rmcrt_test|wasatch] [--timesteps N] [--devices N] [--seed S] [--knob name=value]...\n";
    return 2;
  }
  try {
    ProblemSpecP ups = ProblemSpec::parseFile(deck);
    if (ups->getNodeName() != "Uintah_specification") throw ProblemSetupException("root element must be <Uintah_specification>", __FILE__, __LINE__);
    std::string comp;
    if (ProblemSpecP sc = ups->findBlock("SimulationComponent")) sc->getAttribute("type", comp);
    if (comp != "RMCRT_Test") throw ProblemSetupException("<SimulationComponent type> must be RMCRT_Test, got '" + comp + "'", __FILE__, __LINE__);
    int maxTimesteps = 1;
    if (ProblemSpecP time_ps = ups->findBlock("Time")) time_ps->getWithDefault("max_Timesteps", maxTimesteps, 1);
    if (timesteps > 0) maxTimesteps = timesteps;

    RMCRTCommon::setMaxDevices(devices);
    GridP grid = buildGrid(ups);
    SchedulerP sched = std::make_shared<Scheduler>();
    RMCRT_Test app;
    app.d_udaDirOverride = initUda;
    if (caller == "wasatch") app.d_wasatchCaller = true;
    else if (caller != "rmcrt_test") { std::cerr << "--caller must be rmcrt_test or wasatch\n"; return 2; }
    app.problemSetup(ups, grid, sched);
    if (haveSeed) app.d_RMCRT->setSeed(seed);
    for (auto& kv : knobs) app.d_RMCRT->setKnob(kv.first, kv.second);

    if (parseOnly) {
      for (int l = 0; l < grid->numLevels(); ++l) app.scheduleInitialize(grid->getLevel(l), sched);
      for (int l = 0; l < grid->numLevels(); ++l) app.scheduleTimeAdvance(grid->getLevel(l), sched);
      static const char* kinds[3] = {"requires", "computes", "modifies"};
      for (auto& e : sched->tasks()) {
        std::cout << "TASK " << e.task->getName() << " level=" << getLevel(e.patches)->getIndex() << " patches=" << e.patches->size()
                  << " tg=" << e.tg << " device=" << (e.task->usesDevice() ? 1 : 0) << "\n";
        for (auto& d : e.task->dependencies())
          std::cout << "  " << kinds[d.kind] << " " << d.label->getName() << " dw=" << (d.dw == Task::OldDW ? "OldDW" : "NewDW")
                    << " ghost=" << d.numGhost << " domain=" << (d.domain == Task::NormalDomain ? "normal" : d.domain == Task::CoarseLevel ? "coarse" : "fine")
                    << " offset=" << d.levelOffset << "\n";
      }
      return 0;
    }
    // initialisation "time step": every level gets its properties (sus: scheduleInitialize per level)
    sched->advanceDataWarehouse(0);
    for (int l = 0; l < grid->numLevels(); ++l) app.scheduleInitialize(grid->getLevel(l), sched);
    sched->execute(Scheduler::ALL_GRAPHS);
    sched->clearTasks();

    for (int l = 0; l < grid->numLevels(); ++l) app.scheduleTimeAdvance(grid->getLevel(l), sched);
    std::cout << "rmcrt_test: " << grid->numLevels() << " level(s), task graph TG_RMCRT:";
    for (auto& e : sched->tasks())
      if (e.tg == Scheduler::ALL_GRAPHS || e.tg == RMCRTCommon::TG_RMCRT) std::cout << " " << e.task->getName() << "[L" << getLevel(e.patches)->getIndex() << "]";
    std::cout << std::endl;

    for (int ts = 1; ts <= maxTimesteps; ++ts) {
      sched->advanceDataWarehouse(ts);
      const int tg = app.taskGraphFor(ts);
      std::cout << "timestep " << ts << ": task graph " << (tg == RMCRTCommon::TG_RMCRT ? "TG_RMCRT" : "TG_CARRY_FORWARD") << std::endl;
      sched->execute(tg);
    }

    const Level* fine = grid->getLevel(grid->numLevels() - 1).get();
    DataWarehouse* dw = sched->get_dw(Task::NewDW);
    const bool haveDivQ = dw->exists(app.d_divQLabel, 0, fine->getPatch(0));   // radiometerOnly computes no divQ
    if (haveDivQ) {   // centre line of divQ along x (what RMCRT.m / lineextract look at)
      constCCVariable<double> divQ;
      dw->getLevel(divQ, app.d_divQLabel, 0, fine);
      IntVector lo, hi;
      fine->findInteriorCellIndexRange(lo, hi);
      const int j = (lo.y() + hi.y()) / 2, k = (lo.z() + hi.z()) / 2;
      std::cout << "divQ centre line (j=" << j << ", k=" << k << "):";
      const int stride = std::max(1, (hi.x() - lo.x()) / 8);
      for (int i = lo.x(); i < hi.x(); i += stride) std::cout << " " << divQ[IntVector(i, j, k)];
      std::cout << std::endl;
    }
    if (!udaOut.empty()) {     // <DataArchiver>: the saved labels of the last time step, every level they exist on
      DataArchiver out(udaOut, grid.get());
      out.beginTimestep(maxTimesteps, (double)maxTimesteps, 1.0);
      auto saveOn = [&](const VarLabel* l, const Level* lev) {
        if (!ups->isLabelSaved(l->getName()) || !dw->exists(l, 0, lev->getPatch(0))) return;
        for (int p = 0; p < lev->numPatches(); ++p) {
          const Patch* patch = lev->getPatch(p);
          if (l == RMCRTCommon::d_boundFluxLabel) { constCCVariable<Stencil7> v; dw->get(v, l, 0, patch); out.save(l->getName(), 0, patch, v); }
          else if (l->typeDescription() == TypeDescription::int_type) { constCCVariable<int> v; dw->get(v, l, 0, patch); out.save(l->getName(), 0, patch, v); }
          else if (l->typeDescription() == TypeDescription::float_type) { constCCVariable<float> v; dw->get(v, l, 0, patch); out.save(l->getName(), 0, patch, v); }
          else { constCCVariable<double> v; dw->get(v, l, 0, patch); out.save(l->getName(), 0, patch, v); }
        }
      };
      for (int l = 0; l < grid->numLevels(); ++l)
        for (const VarLabel* lab : {app.d_divQLabel, RMCRTCommon::d_radiationVolqLabel, RMCRTCommon::d_boundFluxLabel, RMCRTCommon::d_sigmaT4Label,
                                    app.d_compAbskgLabel, app.d_colorLabel, app.d_cellTypeLabel})
          saveOn(lab, grid->getLevel(l).get());
      out.endTimestep();
      std::cout << "rmcrt_test: wrote " << udaOut << std::endl;
    }
    if (!dumpDir.empty()) {
      mkdir(dumpDir.c_str(), 0755);
      std::ofstream meta(dumpDir + "/meta.json");
      meta << "{\n";
      bool first = true;
      if (haveDivQ) {
        dumpVar<double>(dumpDir, "divQ", "f8", 1, dw, app.d_divQLabel, fine, meta, first);
        dumpVar<double>(dumpDir, "radiationVolq", "f8", 1, dw, RMCRTCommon::d_radiationVolqLabel, fine, meta, first);
        dumpVar<Stencil7>(dumpDir, "RMCRTboundFlux", "f8", 7, dw, RMCRTCommon::d_boundFluxLabel, fine, meta, first);
        for (int l = 0; l + 1 < grid->numLevels(); ++l) {      // RMCRT_coarseLevel: the coarse-level solve the fine divQ was interpolated from
          const Level* lev = grid->getLevel(l).get();
          if (dw->exists(app.d_divQLabel, 0, lev->getPatch(0)))
            dumpVar<double>(dumpDir, "divQ_L" + std::to_string(l), "f8", 1, dw, app.d_divQLabel, lev, meta, first);
        }
      }
      if (Radiometer* radiometer = app.d_RMCRT->getRadiometer()) {
        if (radiometer->d_VRFluxLabel->typeDescription() == TypeDescription::float_type) {
          dumpVar<float>(dumpDir, "VRFlux", "f4", 1, dw, radiometer->d_VRFluxLabel, fine, meta, first);
          dumpVar<float>(dumpDir, "VRIntensity", "f4", 1, dw, radiometer->d_VRIntensityLabel, fine, meta, first);
        } else {
          dumpVar<double>(dumpDir, "VRFlux", "f8", 1, dw, radiometer->d_VRFluxLabel, fine, meta, first);
          dumpVar<double>(dumpDir, "VRIntensity", "f8", 1, dw, radiometer->d_VRIntensityLabel, fine, meta, first);
        }
      }
      dumpVar<double>(dumpDir, "abskg", "f8", 1, dw, app.d_compAbskgLabel, fine, meta, first);
      dumpVar<int>(dumpDir, "cellType", "i4", 1, dw, app.d_cellTypeLabel, fine, meta, first);
      if (RMCRTCommon::d_sigmaT4Label->typeDescription() == TypeDescription::float_type)
        dumpVar<float>(dumpDir, "sigmaT4", "f4", 1, dw, RMCRTCommon::d_sigmaT4Label, fine, meta, first);
      else
        dumpVar<double>(dumpDir, "sigmaT4", "f8", 1, dw, RMCRTCommon::d_sigmaT4Label, fine, meta, first);
      const Ray::Report& r = app.d_RMCRT->lastReport();
      meta << ",\n  \"report\": {\"seconds\": " << r.seconds << ", \"nRaySteps\": " << r.nRaySteps << ", \"nFluxRaySteps\": " << r.nFluxRaySteps
           << ", \"nDivQCells\": " << r.nDivQCells << ", \"nFluxCells\": " << r.nFluxCells << ", \"devicesUsed\": " << r.devicesUsed << "}\n}\n";
    }
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "rmcrt_test: caught exception: " << e.what() << std::endl;
    return 1;
  }
}
