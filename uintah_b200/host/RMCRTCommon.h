// RMCRTCommon.h -- host-side mirror of the reference's RMCRTCommon plugin surface
// (CCA/Components/Models/Radiation/RMCRT/RMCRTCommon.h:65-320) over the B200 C ABI (include/rmcrt_b200.h).
//
// Same public names, argument meaning and error behaviour as the reference class, so a component that drives the
// reference's Ray (RMCRT_Test, Arches RMCRT_Radiation, Wasatch RadiationSource) drives this one unchanged.  What is
// gone: every per-ray routine (updateSumI, findRayDirection, ray_Origin, reflect, raySignStep) -- they live in the CUDA
// kernels now and are not part of the host surface any more.
#pragma once
#include <map>
#include <mutex>
#include <string>
#include <vector>

#ifdef RMCRT_USE_REAL_UINTAH
#include <CCA/Ports/DataWarehouse.h>
#include <CCA/Ports/Scheduler.h>
#include <Core/Grid/Task.h>
#else
#include "UintahShim.h"
#endif
#include "../../include/rmcrt_b200.h"

namespace Uintah {

class RMCRTCommon {
 public:
  explicit RMCRTCommon(TypeDescription::Type FLT_DBL);
  virtual ~RMCRTCommon();

  /** map the component VarLabels to RMCRT class VarLabels (RMCRTCommon.cc:157-203) */
  void registerVariables(int matl, const VarLabel* abskg, const VarLabel* temperature, const VarLabel* celltype,
                         const VarLabel* divQ, const int whichAlgo);

  /** sigmaT4 = (sigma/pi) T^4, extra cells included (RMCRTCommon.cc:279-343) */
  void sched_sigmaT4(const LevelP& level, SchedulerP& sched, Task::WhichDW temp_dw, const bool includeEC = true);
  template <class T>
  void sigmaT4(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse* old_dw,
               DataWarehouse* new_dw, Task::WhichDW which_temp_dw, const bool includeEC);

  /** zero-initialise a label on every patch of a level (RMCRTCommon.cc:345-376) */
  void sched_initialize_VarLabel(const LevelP& level, SchedulerP& sched, const VarLabel* label);
  template <class T>
  void initialize_VarLabel(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*,
                           DataWarehouse* new_dw, const VarLabel* label);

  /** abskg double -> float when the component's abskg is double and RMCRT runs in float (RMCRTCommon.cc:211-274) */
  void sched_DoubleToFloat(const LevelP& level, SchedulerP& sched, Task::WhichDW myDW);
  void DoubleToFloat(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse* old_dw,
                     DataWarehouse* new_dw, Task::WhichDW which_dw);

  /** the slim variant's pack task (RMCRTCommon.cc:378-427).  The packed record is produced on the device by
      rmcrt_upload_level, so this only registers the dependency set the reference task has. */
  void sched_combineAbskgSigmaT4CellType(const LevelP& level, SchedulerP& sched, Task::WhichDW temp_dw, const bool includeEC);

  /** carry results forward on non-radiation time steps (TG_CARRY_FORWARD; RMCRTCommon.cc:909-974) */
  void sched_CarryForward_FineLevelLabels(const LevelP& level, SchedulerP& sched);
  void sched_carryForward_VarLabels(const LevelP& level, SchedulerP& sched, const std::vector<const VarLabel*> varLabels);
  void sched_CarryForward_Var(const LevelP& level, SchedulerP& sched, const VarLabel* variable, const int tg_num = -1);
  void carryForward_VarLabels(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset* matls, DataWarehouse* old_dw,
                              DataWarehouse* new_dw, const std::vector<const VarLabel*> varLabels);

  /** which DataWarehouse holds abskg on each level (RMCRTCommon.cc:1033-1110) */
  void set_abskg_dw_perLevel(const LevelP& fineLevel, Task::WhichDW which_dw);
  DataWarehouse* get_abskg_dw(const int L, const VarLabel* label, DataWarehouse* new_dw);
  Task::WhichDW get_abskg_whichDW(const int L, const VarLabel* abskg);

  bool isDbgCell(IntVector me);

  //__________________________________ public variables used by components (RMCRTCommon.h:257-315)
  enum DIR { X = 0, Y = 1, Z = 2, NONE = -9 };
  // the reference's comment says "-x +x ..." but its code uses WEST for the -x neighbour
  enum FACE { EAST = 0, WEST = 1, NORTH = 2, SOUTH = 3, TOP = 4, BOT = 5, nFACES = 6 };
  enum GRAPH_TYPE { TG_CARRY_FORWARD = 0, TG_RMCRT = 1, NUM_GRAPHS };
  enum Algorithm { dataOnion, coarseLevel, singleLevel, radiometerOnly };

  double d_sigma_over_pi{0.0};
  int d_flowCell{-1};                       // HARDWIRED (RMCRTCommon.cc:116)
  Ghost::GhostType d_gn{Ghost::None};
  Ghost::GhostType d_gac{Ghost::AroundCells};
  TypeDescription::Type d_FLT_DBL;

  static std::vector<IntVector> d_dbgCells;
  static std::map<std::string, Task::WhichDW> d_abskg_dw;

  static double d_threshold;
  static double d_sigma;
  static double d_sigmaScat;
  static double d_maxRayLength;
  static bool d_isSeedRandom;
  static bool d_allowReflect;
  static int d_destructorCount;

  static int d_matl;
  static int d_whichAlgo;
  static MaterialSet* d_matlSet;
  static std::string d_abskgBC_tag;

  static const VarLabel* d_sigmaT4Label;
  static const VarLabel* d_abskgLabel;
  static const VarLabel* d_divQLabel;
  static const VarLabel* d_boundFluxLabel;
  static const VarLabel* d_radiationVolqLabel;
  static const VarLabel* d_abskgSigmaT4CellTypeLabel;
  static const VarLabel* d_compTempLabel;
  static const VarLabel* d_compAbskgLabel;
  static const VarLabel* d_cellTypeLabel;

  //__________________________________ B200 back end
  /** one rmcrt_ctx per visible device, created on first use with the current parameters; patch -> device is the
      round-robin of GPUGridVariableInfo.cc:610-632 */
  struct Device {
    int ordinal = 0;
    rmcrt_ctx* ctx = nullptr;
    std::map<int, long long> levelStamp;   // level index -> DataWarehouse generation its records were packed from
  };
  std::vector<Device>& devices();
  /** use the device contexts (and the level records already packed there) of another RMCRTCommon: the Radiometer that a
      Ray owns traces through the Ray's records instead of uploading the level a second time */
  void shareDevicesOf(RMCRTCommon* owner) { d_backend = owner ? owner : this; }
  /** limit the number of devices used (0 = all visible); call before the first solve */
  static void setMaxDevices(int n) { s_maxDevices = n; }
  /** throws InternalError / ProblemSetupException with rmcrt_last_error() */
  static void check(rmcrt_ctx* ctx, int rc, const char* what);

  /** getLevel() the three property arrays of `level` (the whole-level gathers of Ray.cc:549-553 / Radiometer.cc:411-417) and
      pack them on every device, once per DataWarehouse generation `stamp` */
  template <class T>
  void uploadLevel(const Level* level, DataWarehouse* abskg_dw, DataWarehouse* sigmaT4_dw, DataWarehouse* celltype_dw, long long stamp);
  static rmcrt_level_desc describe(const Level* level);

 protected:
  virtual void fillParams(rmcrt_params& p) const;   // Ray adds its members
  /** 0 float records, 1 double records (FP64 parity march), 2 half-packed records */
  virtual int recordKind(const rmcrt_params& p) const { return p.marchPrecision == 1 ? 1 : 0; }
  void releaseDevices();
  std::mutex d_deviceMutex;

 private:
  std::vector<Device> d_devices;
  RMCRTCommon* d_backend{this};
  static int s_maxDevices;
};

}  // namespace Uintah
