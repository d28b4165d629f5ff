// Ray.h -- host-side mirror of the reference's `Ray` plugin class (RMCRT/Ray.h:66-406) over the B200 C ABI.
//
// Drop-in surface kept verbatim: constructor, problemSetup (same <RMCRT> keys and defaults), sched_rayTrace /
// sched_rayTrace_dataOnion (same task names, same requires/computes on the same labels, same task graph), the boundary
// condition helpers, the multi-level scheduling entry points and the public members components read.  The task bodies
// obtain the same DataWarehouse variables the reference obtains and hand their raw x-fastest buffers to
// rmcrt_upload_level / rmcrt_trace_divq / rmcrt_trace_dataonion; no ray is traced on the host.
#pragma once
#include <climits>
#include "RMCRTCommon.h"
#include "Radiometer.h"

#include <cfloat>

namespace Uintah {

class Ray : public RMCRTCommon {
 public:
  explicit Ray(TypeDescription::Type FLT_DBL);   // this class can be float or double (Ray.h:70)
  ~Ray();

  bool d_solveBoundaryFlux{false};
  enum modifiesComputes { modifiesVar, computesVar };

  /** Interface to the input file (Ray.cc:154-364) */
  void problemSetup(const ProblemSpecP& prob_spec, const ProblemSpecP& rmcrt_ps, const GridP& grid);

  /** trace rays through a single level (Ray.cc:422-506) */
  void sched_rayTrace(const LevelP& level, SchedulerP& sched, Task::WhichDW abskg_dw, Task::WhichDW sigma_dw,
                      Task::WhichDW celltype_dw, bool modifies_divQ);
  /** multi-level Data Onion (Ray.cc:824-938) */
  void sched_rayTrace_dataOnion(const LevelP& level, SchedulerP& sched, Task::WhichDW abskg_dw, Task::WhichDW sigma_dw,
                                Task::WhichDW celltype_dw, bool modifies_divQ);

  //__________________________________ boundary conditions (Ray.cc:1592-1867)
  void setBC_onOff(const bool onOff) { d_onOff_SetBCs = onOff; }
  void sched_setBoundaryConditions(const LevelP& level, SchedulerP& sched, Task::WhichDW temp_dw, const bool backoutTemp = false);
  void BC_bulletproofing(const ProblemSpecP& rmcrtps, const bool chk_temp, const bool chk_absk);
  template <class T, class V>
  void setBC(CCVariable<T>& Q_CC, const std::string& desc, const Patch* patch, const int mat_id);

  //__________________________________ multi-level tasks
  void sched_Refine_Q(SchedulerP& sched, const PatchSet* patches, const MaterialSet* matls);
  void sched_CoarsenAll(const LevelP& coarseLevel, SchedulerP& sched, const bool modifies_abskg, const bool modifies_sigmaT4);
  void sched_computeCellType(const LevelP& coarseLevel, SchedulerP& sched, const Ray::modifiesComputes which);
  void sched_ROI_Extents(const LevelP& level, SchedulerP& scheduler);

  Radiometer* getRadiometer() { return d_radiometer; }   // Ray.h:151-153

  bool d_coarsenExtraCells{false};

  //__________________________________ B200 additions (no counterpart in the reference)
  struct Report {              // what the reference prints as "RMCRT REPORT" (Ray.cc:803-812)
    double seconds = 0;        // wall time of the last trace task
    unsigned long long nRaySteps = 0, nFluxRaySteps = 0, nDivQCells = 0, nFluxCells = 0;
    int devicesUsed = 0;
  };
  const Report& lastReport() const { return d_report; }
  void setSeed(unsigned long long s) { d_seed = s; }
  /** implementation knobs of rmcrt_params (rayCoherence, skipZeroWeightLegs, l2Persist, genericMarch), applied on the next solve */
  void setKnob(const std::string& name, int value);

 protected:
  void fillParams(rmcrt_params& p) const override;

 private:
  enum rayDirSampleAlgorithm { NAIVE, LATIN_HYPER_CUBE };
  enum ROI_algo { fixed, dynamic, patch_based, boundedRayLength, entireDomain };
  enum cellTypeCoarsenLogic { ROUNDUP, ROUNDDOWN };

  double d_sigmaT4_thld{DBL_MAX};
  double d_abskg_thld{DBL_MAX};
  int d_nDivQRays{10};
  int d_nFluxRays{1};
  int d_orderOfInterpolation{-9};
  IntVector d_haloCells{IntVector(-9, -9, -9)};
  double d_haloLength{-9};
  bool d_solveDivQ{true};
  bool d_CCRays{false};
  bool d_onOff_SetBCs{true};
  bool d_applyFilter{false};
  int d_rayDirSampleAlgo{NAIVE};
  Radiometer* d_radiometer{nullptr};
  int d_cellTypeCoarsenLogic{ROUNDUP};
  ROI_algo d_ROI_algo{entireDomain};
  Point d_ROI_minPt, d_ROI_maxPt;

  const VarLabel* m_timeStepLabel{nullptr};
  const VarLabel* d_mag_grad_abskgLabel{nullptr};
  const VarLabel* d_mag_grad_sigmaT4Label{nullptr};
  const VarLabel* d_flaggedCellsLabel{nullptr};
  const VarLabel* d_ROI_LoCellLabel{nullptr};
  const VarLabel* d_ROI_HiCellLabel{nullptr};
  const VarLabel* d_PPTimerLabel{nullptr};

  unsigned long long d_seed{0x9E3779B97F4A7C15ull};
  std::map<std::string, int> d_knobs;
  Report d_report;
  std::map<int, long long> d_deviceCoarsened;   // coarse level index -> DataWarehouse generation it was derived in
  // dynamic ROI: the ROI_loCell / ROI_hiCell reduction of Ray::ROI_Extents (Ray.cc:1990-1991, 2041-2042) and the generation it belongs to
  IntVector d_dynROI_lo{IntVector(SHRT_MAX, SHRT_MAX, SHRT_MAX)}, d_dynROI_hi{IntVector(-SHRT_MAX, -SHRT_MAX, -SHRT_MAX)};
  long long d_dynROI_stamp{-1};
  GridP d_grid;

  //__________________________________ task bodies
  template <class T>
  void rayTrace(const ProcessorGroup* pg, const PatchSubset* patches, const MaterialSubset* matls, DataWarehouse* old_dw,
                DataWarehouse* new_dw, bool modifies_divQ, Task::WhichDW which_abskg_dw, Task::WhichDW which_sigmaT4_dw,
                Task::WhichDW which_celltype_dw);
  template <class T>
  void rayTrace_dataOnion(const ProcessorGroup* pg, const PatchSubset* patches, const MaterialSubset* matls, DataWarehouse* old_dw,
                          DataWarehouse* new_dw, bool modifies_divQ, Task::WhichDW which_abskg_dw, Task::WhichDW which_sigmaT4_dw,
                          Task::WhichDW which_celltype_dw);
  template <class T>
  void setBoundaryConditions(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*,
                             DataWarehouse* new_dw, Task::WhichDW temp_dw, const bool backoutTemp);
  template <class T>
  void coarsenAll(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw);
  template <class T>
  void ROI_Extents(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw);   // Ray.h:283-288
  void refine_Q(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset* matls, DataWarehouse* old_dw, DataWarehouse* new_dw);   // Ray.h:307-311
  bool isDeviceLevel(const Level* level) const;

  //__________________________________ helpers
  int recordKind(const rmcrt_params& p) const override;
  /** run `fn(device, patchIndex)` for every patch, patches dealt round-robin to the devices, one host thread per device */
  void forEachPatchOnDevices(const PatchSubset* patches, const std::function<void(Device&, int)>& fn);
  template <class T>
  void traceCommon(bool dataOnion, const PatchSubset* patches, DataWarehouse* new_dw, bool modifies_divQ);
};

}  // namespace Uintah
