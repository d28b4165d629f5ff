// Radiometer.h -- host-side mirror of the reference's virtual radiometer plugin class (RMCRT/Radiometer.h:42-170) over
// the B200 C ABI.  Same constructor, problemSetup (same <Radiometer> keys, same bullet proofing), sched_radiometer /
// sched_initialize_VRFlux (same task names, labels, spatial patch set) and the public VRFlux / VRIntensity labels; the ray
// loop of Radiometer::radiometerFlux<T> (Radiometer.cc:428-500) runs in radiometer_kernel behind rmcrt_trace_radiometer.
#pragma once
#include "RMCRTCommon.h"

namespace Uintah {

class Radiometer : public RMCRTCommon {
 public:
  /** `backend`: the RMCRTCommon whose device contexts / packed level records are used (the owning Ray); nullptr = own */
  explicit Radiometer(const TypeDescription::Type FLT_DBL, RMCRTCommon* backend = nullptr);
  ~Radiometer();

  void problemSetup(const ProblemSpecP& prob_spec, const ProblemSpecP& rmcrt_ps, const GridP& grid);

  void sched_radiometer(const LevelP& level, SchedulerP& sched, Task::WhichDW abskg_dw, Task::WhichDW sigma_dw, Task::WhichDW celltype_dw);
  void sched_initialize_VRFlux(const LevelP& level, SchedulerP& sched);

  /** VRFlux / VRIntensity of the radiometer cells of one patch (Radiometer.cc:428-500).  The level records must be on the
      devices already (uploadLevel); the reference's MTRand / property-array arguments have no counterpart here. */
  template <class T>
  void radiometerFlux(const Patch* patch, const Level* level, DataWarehouse* new_dw, const bool modifiesFlux);

  const VarLabel* d_VRFluxLabel{nullptr};
  const VarLabel* d_VRIntensityLabel{nullptr};

  size_t numRadiometers() const { return d_radiometers.size(); }

 private:
  struct radiometer {          // Radiometer.h:117-127
    rmcrt_radiometer angles;   // theta_rotate, phi_rotate, xi_rotate, theta_viewAngle, range, solidAngle (+ nRays)
    Point locationsMin, locationsMax;
  };
  std::vector<radiometer> d_radiometers;

  template <class T>
  void initialize_VRFlux(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw);
  template <class T>
  void radiometerTask(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse* old_dw, DataWarehouse* new_dw,
                      Task::WhichDW which_abskg_dw, Task::WhichDW which_sigmaT4_dw, Task::WhichDW which_celltype_dw);
  void getPatchSet(SchedulerP& sched, const LevelP& level, PatchSet& ps);
};

}  // namespace Uintah
