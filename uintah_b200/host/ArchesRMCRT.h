// ArchesRMCRT.h -- host-side mirror of the Arches caller of the RMCRT plugin, RMCRT_Radiation
// (CCA/Components/Arches/SourceTerms/RMCRT.{h,cc}): the source term that feeds Ray with sumAbsk / sigmaT4 / cellType, schedules
// the radiation solve on the Arches level and splits RMCRTboundFlux into the six radiationFlux{E,W,N,S,T,B} fields.
//
// Same public method names, .ups keys, task names, labels and dependencies as the reference class; the three task bodies that
// touch every cell (sigmaT4, sumAbsk, stencilToDBLs) run on the device through rmcrt_arches_sigmaT4 / rmcrt_arches_sumAbsk /
// rmcrt_stencil_to_dbls.  The Arches-internal collaborators the reference constructor takes (ArchesLabel, BoundaryCondition,
// TableLookup, MaterialManager, SourceTermBase) are reduced to what this path reads from them: the two labels of ArchesLabel.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "Ray.h"

namespace Uintah {

struct ArchesLabel {                              // Arches/ArchesLabel.h: only the members RMCRT.cc reads
  const VarLabel* d_volFractionLabel{nullptr};   // "volFraction": > 1e-16 in gas cells, 0 in walls / intrusions
  const VarLabel* d_cellTypeLabel{nullptr};      // "cellType"
};

class RMCRT_Radiation {
 public:
  using string_vector = std::vector<std::string>;
  enum GRAPH_TYPE { TG_CARRY_FORWARD = 0, TG_RMCRT = 1, NUM_GRAPHS };   // RMCRT.h:108-112
  enum Algorithm { dataOnion, dataOnionSlim, coarseLevel, singleLevel, radiometerOnly };

  RMCRT_Radiation(std::string srcName, ArchesLabel* labels, string_vector reqLabelNames, std::string type);   // RMCRT.cc:47-80
  ~RMCRT_Radiation();

  void problemSetup(const ProblemSpecP& db);                    // RMCRT.cc:103-230
  void extraSetup(GridP& grid);                                 // RMCRT.cc:236-346 (BoundaryCondition*, TableLookup* dropped)
  void sched_initialize(const LevelP& level, SchedulerP& sched);   // RMCRT.cc:574-660
  void sched_computeSource(const LevelP& level, SchedulerP& sched, int timeSubStep);   // RMCRT.cc:353-570

  void sched_sigmaT4(const LevelP& level, SchedulerP& sched);      // RMCRT.cc:900-934
  void sched_sumAbsk(const LevelP& level, SchedulerP& sched);      // RMCRT.cc:1023-1056
  void sched_fluxInit(const LevelP& level, SchedulerP& sched);     // RMCRT.cc:1298-1320
  void sched_stencilToDBLs(const LevelP& level, SchedulerP& sched);   // RMCRT.cc:1357-1384

  Ray* getRay() { return m_RMCRT; }
  const VarLabel* getSrcLabel() const { return _src_label; }
  int calcFrequency() const { return m_radCalc_freq; }

  const VarLabel* m_radFluxE_Label{nullptr};
  const VarLabel* m_radFluxW_Label{nullptr};
  const VarLabel* m_radFluxN_Label{nullptr};
  const VarLabel* m_radFluxS_Label{nullptr};
  const VarLabel* m_radFluxT_Label{nullptr};
  const VarLabel* m_radFluxB_Label{nullptr};
  const VarLabel* m_sumAbsk_Label{nullptr};

 private:
  void initialize(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw);
  void fluxInit(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw);
  template <class T>
  void sigmaT4(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse* old_dw, DataWarehouse* new_dw,
               Task::WhichDW which_dw);
  template <class T>
  void sumAbsk(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse* old_dw, DataWarehouse* new_dw,
               Task::WhichDW which_dw);
  void stencilToDBLs(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw);

  std::string m_srcName;
  ArchesLabel* m_labels;
  ProblemSpecP m_ps;
  GridP m_grid;
  Ray* m_RMCRT{nullptr};
  const VarLabel* _src_label{nullptr};
  const VarLabel* m_gasTemp_Label{nullptr};
  TypeDescription::Type m_FLT_DBL{TypeDescription::double_type};
  int m_matl{0};
  int m_archesLevelIndex{-9};
  int m_whichAlgo{singleLevel};
  int m_radCalc_freq{1};
  int m_nQn_part{0};
  int m_nPartGasLabels{0};
  bool m_all_rk{false};
  bool m_do_partRadiation{false};
  Ghost::GhostType m_gn{Ghost::None};
  string_vector m_partGas_temp_names, m_partGas_absk_names;
  std::vector<const VarLabel*> m_partGas_temp_Labels, m_partGas_absk_Labels;
};

}  // namespace Uintah
