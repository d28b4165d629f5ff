// Arches_Test.cc -- stand-alone driver of the Arches-side mirror RMCRT_Radiation (ArchesRMCRT.{h,cc}): stands in for the part of
// Arches' ExplicitSolver that owns the radiation source term (ExplicitSolver.cc:262-280, 1257: problemSetup of the
// <src type="rmcrt_radiation"> block, sched_initialize, sched_computeSource on RK step 0, task graph chosen by calc_frequency).
//
//   arches_rmcrt_test <deck.ups> [--dump DIR] [--timesteps N] [--knob name=value]... [--parse-only]
//
// The fields Arches would compute (radiation_temperature, gas and particle absorption coefficients and temperatures, volFraction,
// cellType) are synthesised from integer index formulas (exact in floating point, so tests/test_arches_host.py rebuilds them bit
// for bit): see ArchesFields below.
#include <cstdlib>

#include "ArchesRMCRT.h"
#include "DriverCommon.h"

using namespace Uintah;

namespace {

inline int pmod(int a, int m) { return ((a % m) + m) % m; }

struct ArchesFields {     // the synthetic combustion-like state; c = cell index, N = interior cells per side of the level
  static bool intrusion(const IntVector& c, const IntVector& N) {
    return c.x() >= N.x() / 4 && c.x() < N.x() / 2 && c.y() >= N.y() / 4 && c.y() < N.y() / 2 && c.z() >= N.z() / 2 && c.z() < 3 * N.z() / 4;
  }
  static bool dark(const IntVector& c) { return pmod(c.x() + c.y() + c.z(), 11) == 0; }      // gas cells with no absorber at all
  static double gasTemp(const IntVector& c) { return 900.0 + 7.5 * pmod(3 * c.x() + 5 * c.y() + 7 * c.z(), 97); }
  static double gasAbsk(const IntVector& c) { return dark(c) ? 0.0 : 0.1 + 0.004 * pmod(c.x() + 2 * c.y() + 3 * c.z(), 50); }
  static double partAbsk(const IntVector& c, int q) { return dark(c) ? 0.0 : 0.02 * (q + 1) * (1 + pmod(c.x() + c.y() + c.z() + q, 3)); }
  static double partTemp(const IntVector& c, int q) { return gasTemp(c) + 50.0 * (q + 1); }
  static constexpr double wallTemp = 600.0, intrusionTemp = 700.0;
  static constexpr int wallType = 8;
};

class Arches_Test {
 public:
  Arches_Test() {
    d_labels.d_volFractionLabel = VarLabel::create("volFraction", TypeDescription::double_type);
    d_labels.d_cellTypeLabel = VarLabel::create("cellType", TypeDescription::int_type);
  }
  ~Arches_Test() {
    d_src.reset();
    VarLabel::destroy(d_labels.d_volFractionLabel);
    VarLabel::destroy(d_labels.d_cellTypeLabel);
    for (auto l : d_inputs) VarLabel::destroy(l);
  }

  // <CFD><ARCHES><TransportEqns><Sources><src label="divQ" type="rmcrt_radiation"> (SourceTermFactory; ExplicitSolver.cc:262-280)
  void problemSetup(const ProblemSpecP& root, GridP& grid, SchedulerP& sched) {
    d_grid = grid;
    sched->setNumTaskGraphs(RMCRT_Radiation::NUM_GRAPHS);
    ProblemSpecP ps = root;
    for (const char* tag : {"CFD", "ARCHES", "TransportEqns", "Sources"}) {
      ps = ps->findBlock(tag);
      if (!ps) throw ProblemSetupException(std::string("arches_rmcrt_test: <") + tag + "> block not found", __FILE__, __LINE__);
    }
    ProblemSpecP src_ps;
    std::string label, type;
    for (ProblemSpecP s = ps->findBlock("src"); s; s = s->findNextBlock("src")) {
      s->getAttribute("type", type);
      if (type == "rmcrt_radiation") { src_ps = s; s->getAttribute("label", label); break; }
    }
    if (!src_ps) throw ProblemSetupException("arches_rmcrt_test: no <src type=\"rmcrt_radiation\">", __FILE__, __LINE__);
    if (!src_ps->findBlock("calc_frequency")) throw ProblemSetupException("required input <calc_frequency> is missing in <src>", __FILE__, __LINE__);
    d_src.reset(new RMCRT_Radiation(label, &d_labels, {}, type));
    d_src->problemSetup(src_ps);
    d_src->extraSetup(grid);
    // the labels Arches' property models would own
    ProblemSpecP ac = src_ps->findBlock("absorptionCoeffs");
    std::string gas;
    ac->require("gas_absk", gas);
    d_inputs.push_back(VarLabel::create("radiation_temperature", TypeDescription::double_type));
    d_inputs.push_back(VarLabel::create(gas, TypeDescription::double_type));
    if (ProblemSpecP pr = ac->findBlock("includeParticleRad")) {
      std::string a, t;
      pr->require("particle_absk", a);
      pr->require("part_temp_label", t);
      root->findBlock("DQMOM") ? root->findBlock("DQMOM")->require("number_quad_nodes", d_nQn) : root->findBlock("CQMOM")->require("number_quad_nodes", d_nQn);
      for (int q = 0; q < d_nQn; ++q) {
        d_inputs.push_back(VarLabel::create(t + "_" + std::to_string(q), TypeDescription::double_type));
        d_inputs.push_back(VarLabel::create(a + "_" + std::to_string(q), TypeDescription::double_type));
      }
    }
  }

  void scheduleInitialize(const LevelP& level, SchedulerP& sched) {
    if (level->getIndex() != d_grid->numLevels() - 1) return;    // Arches lives on the finest level
    Task* task = new Task("Arches_Test::initialize", this, &Arches_Test::initialize);
    task->computesVar(d_labels.d_volFractionLabel);
    task->computesVar(d_labels.d_cellTypeLabel);
    for (auto l : d_inputs) task->computesVar(l);
    sched->addTask(task, level->eachPatch(), RMCRTCommon::d_matlSet);
    d_src->sched_initialize(level, sched);
  }

  void scheduleTimeAdvance(const LevelP& level, SchedulerP& sched) {
    if (level->getIndex() != d_grid->numLevels() - 1) return;
    // the properties are steady here: carry them Old -> New on every task graph (a real Arches recomputes them)
    Task* tsk = new Task("Arches_Test::carryProperties", this, &Arches_Test::carryProperties);
    std::vector<const VarLabel*> all = d_inputs;
    all.push_back(d_labels.d_volFractionLabel);
    all.push_back(d_labels.d_cellTypeLabel);
    for (auto l : all) { tsk->requiresVar(Task::OldDW, l); tsk->computesVar(l); }
    sched->addTask(tsk, level->eachPatch(), RMCRTCommon::d_matlSet, Scheduler::ALL_GRAPHS);
    d_src->sched_computeSource(level, sched, 0);
  }

  // ExplicitSolver.cc:1257: the graph of step ts
  int taskGraphFor(int ts) const { return ((ts % d_src->calcFrequency()) == 0 || ts == 1) ? RMCRT_Radiation::TG_RMCRT : RMCRT_Radiation::TG_CARRY_FORWARD; }

  void initialize(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw) {
    const Level* level = getLevel(patches);
    IntVector ilo, ihi;
    level->findInteriorCellIndexRange(ilo, ihi);
    const IntVector N = ihi - ilo;
    for (int p = 0; p < patches->size(); p++) {
      const Patch* patch = patches->get(p);
      CCVariable<double> vol;
      CCVariable<int> cellType;
      std::vector<CCVariable<double>> in(d_inputs.size());
      new_dw->allocateAndPut(vol, d_labels.d_volFractionLabel, 0, patch);
      new_dw->allocateAndPut(cellType, d_labels.d_cellTypeLabel, 0, patch);
      for (size_t i = 0; i < d_inputs.size(); ++i) new_dw->allocateAndPut(in[i], d_inputs[i], 0, patch);
      const IntVector lo = patch->getExtraCellLowIndex(), hi = patch->getExtraCellHighIndex();
      for (int k = lo.z(); k < hi.z(); ++k)
        for (int j = lo.y(); j < hi.y(); ++j)
          for (int i = lo.x(); i < hi.x(); ++i) {
            const IntVector c(i, j, k);
            bool wall = false;
            for (int d = 0; d < 3; ++d) wall = wall || c[d] < ilo[d] || c[d] >= ihi[d];
            const bool solid = wall || ArchesFields::intrusion(c, N);
            vol[c] = solid ? 0.0 : 1.0;
            cellType[c] = solid ? ArchesFields::wallType : -1;
            in[0][c] = wall ? ArchesFields::wallTemp : (solid ? ArchesFields::intrusionTemp : ArchesFields::gasTemp(c));
            in[1][c] = solid ? 0.0 : ArchesFields::gasAbsk(c);
            for (int q = 0; q < d_nQn; ++q) {
              in[2 + 2 * q][c] = solid ? in[0][c] : ArchesFields::partTemp(c, q);
              in[3 + 2 * q][c] = solid ? 0.0 : ArchesFields::partAbsk(c, q);
            }
          }
    }
  }
  void carryProperties(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset* matls, DataWarehouse* old_dw, DataWarehouse* new_dw) {
    for (auto l : d_inputs) new_dw->transferFrom(old_dw, l, patches, matls);
    new_dw->transferFrom(old_dw, d_labels.d_volFractionLabel, patches, matls);
    new_dw->transferFrom(old_dw, d_labels.d_cellTypeLabel, patches, matls);
  }

  ArchesLabel d_labels;
  std::unique_ptr<RMCRT_Radiation> d_src;
  std::vector<const VarLabel*> d_inputs;   // radiation_temperature, gas absk, then (temp_q, absk_q) pairs
  GridP d_grid;
  int d_nQn = 0;
};

}  // namespace

int main(int argc, char** argv) {
  std::string deck, dumpDir;
  int timesteps = -1;
  bool parseOnly = false;
  std::vector<std::pair<std::string, int>> knobs;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> std::string { if (i + 1 >= argc) { std::cerr << "missing value after " << a << "\n"; std::exit(2); } return argv[++i]; };
    if (a == "--parse-only") parseOnly = true;
    else if (a == "--dump") dumpDir = next();
    else if (a == "--timesteps") timesteps = std::atoi(next().c_str());
    else if (a == "--knob") {
      const std::string kv = next();
      const size_t eq = kv.find('=');
      if (eq == std::string::npos) { std::cerr << "--knob needs name=value\n"; return 2; }
      knobs.emplace_back(kv.substr(0, eq), std::atoi(kv.substr(eq + 1).c_str()));
    } else if (a[0] == '-') { std::cerr << "unknown option " << a << "\n"; return 2; }
    else deck = a;
  }
  if (deck.empty()) { std::cerr << "usage: arches_rmcrt_test <deck.ups> [--dump DIR] [--timesteps N] [--knob name=value]... [--parse-only]\n"; return 2; }
  try {
    ProblemSpecP ups = ProblemSpec::parseFile(deck);
    int maxTimesteps = 1;
    if (ProblemSpecP time_ps = ups->findBlock("Time")) time_ps->getWithDefault("max_Timesteps", maxTimesteps, 1);
    if (timesteps > 0) maxTimesteps = timesteps;
    RMCRTCommon::setMaxDevices(1);
    GridP grid = buildGrid(ups);
    SchedulerP sched = std::make_shared<Scheduler>();
    Arches_Test app;
    app.problemSetup(ups, grid, sched);
    for (auto& kv : knobs) app.d_src->getRay()->setKnob(kv.first, kv.second);

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
    sched->advanceDataWarehouse(0);
    for (int l = 0; l < grid->numLevels(); ++l) app.scheduleInitialize(grid->getLevel(l), sched);
    sched->execute(Scheduler::ALL_GRAPHS);
    sched->clearTasks();
    for (int l = 0; l < grid->numLevels(); ++l) app.scheduleTimeAdvance(grid->getLevel(l), sched);
    for (int ts = 1; ts <= maxTimesteps; ++ts) {
      sched->advanceDataWarehouse(ts);
      const int tg = app.taskGraphFor(ts);
      std::cout << "timestep " << ts << ": task graph " << (tg == RMCRT_Radiation::TG_RMCRT ? "TG_RMCRT" : "TG_CARRY_FORWARD") << std::endl;
      sched->execute(tg);
    }
    if (!dumpDir.empty()) {
      const Level* fine = grid->getLevel(grid->numLevels() - 1).get();
      DataWarehouse* dw = sched->get_dw(Task::NewDW);
      mkdir(dumpDir.c_str(), 0755);
      std::ofstream meta(dumpDir + "/meta.json");
      meta << "{\n";
      bool first = true;
      Ray* ray = app.d_src->getRay();
      dumpVar<double>(dumpDir, "divQ", "f8", 1, dw, app.d_src->getSrcLabel(), fine, meta, first);
      dumpVar<double>(dumpDir, "radiationVolq", "f8", 1, dw, RMCRTCommon::d_radiationVolqLabel, fine, meta, first);
      dumpVar<Stencil7>(dumpDir, "RMCRTboundFlux", "f8", 7, dw, RMCRTCommon::d_boundFluxLabel, fine, meta, first);
      if (ray->d_solveBoundaryFlux) {
        const VarLabel* six[6] = {app.d_src->m_radFluxE_Label, app.d_src->m_radFluxW_Label, app.d_src->m_radFluxN_Label,
                                  app.d_src->m_radFluxS_Label, app.d_src->m_radFluxT_Label, app.d_src->m_radFluxB_Label};
        for (auto l : six) dumpVar<double>(dumpDir, l->getName(), "f8", 1, dw, l, fine, meta, first);
      }
      if (RMCRTCommon::d_sigmaT4Label->typeDescription() == TypeDescription::float_type) {
        dumpVar<float>(dumpDir, "sigmaT4", "f4", 1, dw, RMCRTCommon::d_sigmaT4Label, fine, meta, first);
        dumpVar<float>(dumpDir, "RMCRT_sumAbsk", "f4", 1, dw, app.d_src->m_sumAbsk_Label, fine, meta, first);
      } else {
        dumpVar<double>(dumpDir, "sigmaT4", "f8", 1, dw, RMCRTCommon::d_sigmaT4Label, fine, meta, first);
        dumpVar<double>(dumpDir, "RMCRT_sumAbsk", "f8", 1, dw, app.d_src->m_sumAbsk_Label, fine, meta, first);
      }
      dumpVar<int>(dumpDir, "cellType", "i4", 1, dw, app.d_labels.d_cellTypeLabel, fine, meta, first);
      dumpVar<double>(dumpDir, "volFraction", "f8", 1, dw, app.d_labels.d_volFractionLabel, fine, meta, first);
      for (auto l : app.d_inputs) dumpVar<double>(dumpDir, l->getName(), "f8", 1, dw, l, fine, meta, first);
      const Ray::Report& r = ray->lastReport();
      meta << ",\n  \"report\": {\"seconds\": " << r.seconds << ", \"nRaySteps\": " << r.nRaySteps << ", \"nFluxRaySteps\": " << r.nFluxRaySteps
           << ", \"nDivQCells\": " << r.nDivQCells << ", \"nFluxCells\": " << r.nFluxCells << ", \"devicesUsed\": " << r.devicesUsed << "}\n}\n";
    }
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "arches_rmcrt_test: caught exception: " << e.what() << std::endl;
    return 1;
  }
}
