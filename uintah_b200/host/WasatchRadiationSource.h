// WasatchRadiationSource.h -- host-side mirror of the second consumer of the RMCRT plugin, Wasatch's RadiationSource expression
// (CCA/Components/Wasatch/Expressions/RadiationSource.{h,cc}).  It drives Ray differently from RMCRT_Test and Arches:
// divQ is a field the expression tree has already allocated, so the trace task is scheduled with modifies_divQ = TRUE
// (Ray::rayTrace fetches divQ / RMCRTboundFlux / radiationVolq with getModifiable and must leave every cell it does not compute
// untouched -- zeroInit = 0 at the C ABI), and the boundary temperature comes from the component's temperature field
// (sched_setBoundaryConditions(level, sched, tempDW, backoutTemp = false)).
// What is dropped: the ExprLib / SpatialOps plumbing (Expr::Expression<SVolField>, field requests, the Builder) -- none of it touches
// the path.  Line numbers cite RadiationSource.cc of the reference.
#pragma once
#include "Ray.h"

namespace WasatchCore {

class RadiationSource {
 public:
  // RadiationSource.cc:58-96: labels from the tag names, registerVariables(.., singleLevel), problemSetup, BC_bulletproofing
  RadiationSource(const std::string& radiationSourceName, const std::string& temperatureTag, const std::string& absorptionTag,
                  const std::string& celltypeTag, Uintah::Ray* rmcrt, const Uintah::ProblemSpecP& radiationSpec, Uintah::GridP grid)
      : temperatureLabel_(Uintah::VarLabel::create(temperatureTag, Uintah::TypeDescription::double_type)),
        absorptionLabel_(Uintah::VarLabel::create(absorptionTag, Uintah::TypeDescription::double_type)),
        celltypeLabel_(Uintah::VarLabel::create(celltypeTag, Uintah::TypeDescription::int_type)),
        divqLabel_(Uintah::VarLabel::create(radiationSourceName, Uintah::TypeDescription::double_type)),
        rmcrt_(rmcrt) {
    const int whichAlgo = Uintah::RMCRTCommon::Algorithm::singleLevel;
    rmcrt_->registerVariables(0, absorptionLabel_, temperatureLabel_, celltypeLabel_, divqLabel_, whichAlgo);
    rmcrt_->problemSetup(radiationSpec, radiationSpec, grid);
    rmcrt_->BC_bulletproofing(radiationSpec, true, true);
  }
  ~RadiationSource() {   // 131-137
    Uintah::VarLabel::destroy(temperatureLabel_);
    Uintah::VarLabel::destroy(absorptionLabel_);
    Uintah::VarLabel::destroy(celltypeLabel_);
    Uintah::VarLabel::destroy(divqLabel_);
  }

  // 100-109
  void schedule_setup_bndflux(const Uintah::LevelP& level, Uintah::SchedulerP sched, const Uintah::MaterialSet* const materials) {
    Uintah::Task* task = new Uintah::Task("RadiationSource: setup bndflux", this, &RadiationSource::setup_bndflux);
    task->computesVar(Uintah::RMCRTCommon::d_boundFluxLabel);
    sched->addTask(task, level->eachPatch(), materials, Uintah::RMCRTCommon::TG_RMCRT);
  }

  // 141-172
  void schedule_ray_tracing(const Uintah::LevelP& level, Uintah::SchedulerP sched, const Uintah::MaterialSet* const materials, const int RKStage) {
    using namespace Uintah;
    if (RKStage > 1) return;     // only on the first RK stage
    rmcrt_->sched_CarryForward_FineLevelLabels(level, sched);
    schedule_setup_bndflux(level, sched, materials);
    const bool modifiesDivQ = true;
    const LevelP& fineLevel = level;      // grid->getLevel(0): Wasatch is single level
    const Task::WhichDW tempDW = Task::NewDW, sigmaT4DW = Task::NewDW, celltypeDW = Task::NewDW, notUsed = Task::None;
    rmcrt_->set_abskg_dw_perLevel(fineLevel, Task::NewDW);
    rmcrt_->sched_DoubleToFloat(fineLevel, sched, notUsed);      // if needed convert absorptionLabel: double -> float
    rmcrt_->sched_sigmaT4(fineLevel, sched, tempDW);
    rmcrt_->sched_setBoundaryConditions(level, sched, tempDW, false);
    rmcrt_->sched_rayTrace(level, sched, notUsed, sigmaT4DW, celltypeDW, modifiesDivQ);
  }

 private:
  // 113-127
  void setup_bndflux(const Uintah::ProcessorGroup*, const Uintah::PatchSubset* patches, const Uintah::MaterialSubset*, Uintah::DataWarehouse*,
                     Uintah::DataWarehouse* newDW) {
    for (int p = 0; p < patches->size(); p++) {
      const Uintah::Patch* patch = patches->get(p);
      Uintah::CCVariable<Uintah::Stencil7> boundFlux;
      newDW->allocateAndPut(boundFlux, Uintah::RMCRTCommon::d_boundFluxLabel, 0, patch);
      Uintah::Stencil7 zero;
      zero.initialize(0.0);
      boundFlux.initialize(zero);
    }
  }
  const Uintah::VarLabel *temperatureLabel_, *absorptionLabel_, *celltypeLabel_, *divqLabel_;
  Uintah::Ray* rmcrt_;
};

}  // namespace WasatchCore
