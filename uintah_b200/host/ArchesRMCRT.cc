// ArchesRMCRT.cc -- see ArchesRMCRT.h.  Line numbers cite CCA/Components/Arches/SourceTerms/RMCRT.cc of the reference.
#include "ArchesRMCRT.h"

#include <iostream>
#include <sstream>

#include "Radiometer.h"

namespace Uintah {

//______________________________________________________________________ RMCRT.cc:47-80
RMCRT_Radiation::RMCRT_Radiation(std::string srcName, ArchesLabel* labels, string_vector /*reqLabelNames*/, std::string /*type*/)
    : m_srcName(srcName), m_labels(labels) {
  _src_label = VarLabel::create(srcName, TypeDescription::double_type);
  m_partGas_temp_names.push_back("radiation_temperature");   // HARDWIRED in the reference as well (66)
  m_matl = 0;                                                // the Arches material, index 0 (70-71)
  m_radFluxE_Label = VarLabel::create("radiationFluxE", TypeDescription::double_type);
  m_radFluxW_Label = VarLabel::create("radiationFluxW", TypeDescription::double_type);
  m_radFluxN_Label = VarLabel::create("radiationFluxN", TypeDescription::double_type);
  m_radFluxS_Label = VarLabel::create("radiationFluxS", TypeDescription::double_type);
  m_radFluxT_Label = VarLabel::create("radiationFluxT", TypeDescription::double_type);
  m_radFluxB_Label = VarLabel::create("radiationFluxB", TypeDescription::double_type);
}

RMCRT_Radiation::~RMCRT_Radiation() {   // 84-98
  delete m_RMCRT;
  VarLabel::destroy(_src_label);
  VarLabel::destroy(m_radFluxE_Label);
  VarLabel::destroy(m_radFluxW_Label);
  VarLabel::destroy(m_radFluxN_Label);
  VarLabel::destroy(m_radFluxS_Label);
  VarLabel::destroy(m_radFluxT_Label);
  VarLabel::destroy(m_radFluxB_Label);
  VarLabel::destroy(m_sumAbsk_Label);
  for (auto l : m_partGas_temp_Labels) VarLabel::destroy(l);
  for (auto l : m_partGas_absk_Labels) VarLabel::destroy(l);
}

//______________________________________________________________________ RMCRT.cc:103-230
void RMCRT_Radiation::problemSetup(const ProblemSpecP& inputdb) {
  m_ps = inputdb;
  m_ps->getWithDefault("calc_on_all_RKsteps", m_all_rk, false);
  m_ps->getWithDefault("calc_frequency", m_radCalc_freq, 1);   // read by ExplicitSolver.cc:270 in the reference

  ProblemSpecP ac_ps = m_ps->findBlock("absorptionCoeffs");   // gas absorption coefficient
  if (ac_ps == nullptr) throw ProblemSetupException("Error: RMCRT - <absorptionCoeffs> is not defined.", __FILE__, __LINE__);
  std::string gas_absk_name;
  ac_ps->require("gas_absk", gas_absk_name);
  m_partGas_absk_names.push_back(gas_absk_name);

  if (m_all_rk) throw ProblemSetupException("ERROR:  RMCRT_radiation only works if calc_on_all_RKstes = false", __FILE__, __LINE__);

  ProblemSpecP rmcrt_ps = m_ps->findBlock("RMCRT");
  if (!rmcrt_ps) throw ProblemSetupException("ERROR:  RMCRT_radiation, the xml tag <RMCRT> was not found", __FILE__, __LINE__);

  std::map<std::string, std::string> type;   // are we using floats for all-to-all variables
  rmcrt_ps->getAttributes(type);
  if (type["type"] == "float") m_FLT_DBL = TypeDescription::float_type;

  m_RMCRT = new Ray(m_FLT_DBL);
  m_RMCRT->setBC_onOff(false);

  if (ProblemSpecP alg_ps = rmcrt_ps->findBlock("algorithm")) {   // 147-179
    std::string t = "nullptr";
    alg_ps->getAttribute("type", t);
    if (t == "dataOnion") { m_whichAlgo = dataOnion; m_RMCRT->setBC_onOff(true); }
    else if (t == "dataOnionSlim") { m_whichAlgo = dataOnionSlim; m_RMCRT->setBC_onOff(true); }
    else if (t == "RMCRT_coarseLevel") { m_whichAlgo = coarseLevel; m_RMCRT->setBC_onOff(true); }
    else if (t == "singleLevel") m_whichAlgo = singleLevel;
    else if (t == "radiometerOnly") m_whichAlgo = radiometerOnly;
  }

  //__________________________________ particle contributions (183-228): absk_0..absk_{N-1}, temp_0..temp_{N-1}
  if (ProblemSpecP icpr_ps = ac_ps->findBlock("includeParticleRad")) {
    m_do_partRadiation = true;
    std::string pTempName, pAbskName;
    icpr_ps->require("particle_absk", pAbskName);
    icpr_ps->require("part_temp_label", pTempName);
    // ArchesCore::get_num_env: the number of DQMOM / CQMOM environments of the input file
    ProblemSpecP root = m_ps->getRootNode();
    ProblemSpecP dq = root->findBlock("DQMOM"), cq = root->findBlock("CQMOM");
    if (dq) dq->require("number_quad_nodes", m_nQn_part);
    else if (cq) cq->require("number_quad_nodes", m_nQn_part);
    else throw ProblemSetupException("RMCRT: This method only works for DQMOM/CQMOM.", __FILE__, __LINE__);
    for (int qn = 0; qn < m_nQn_part; qn++) {
      std::stringstream absk, temp;
      temp << pTempName << "_" << qn;
      absk << pAbskName << "_" << qn;
      m_partGas_temp_names.push_back(temp.str());
      m_partGas_absk_names.push_back(absk.str());
    }
  }
}

//______________________________________________________________________ RMCRT.cc:236-346
void RMCRT_Radiation::extraSetup(GridP& grid) {
  m_grid = grid;
  m_sumAbsk_Label = VarLabel::create("RMCRT_sumAbsk", m_FLT_DBL);
  for (size_t i = 0; i < m_partGas_absk_names.size(); i++) {    // VarLabel::find in the reference; the shim's create() finds or makes
    m_partGas_temp_Labels.push_back(VarLabel::create(m_partGas_temp_names[i], TypeDescription::double_type));
    m_partGas_absk_Labels.push_back(VarLabel::create(m_partGas_absk_names[i], TypeDescription::double_type));
  }
  m_gasTemp_Label = m_partGas_temp_Labels[0];
  m_nPartGasLabels = (int)m_partGas_absk_Labels.size();

  std::cout << "\n __________________________________ RMCRT SETTINGS\n"
            << "  - Temperature label:          " << m_partGas_temp_names[0] << "\n"
            << "  - gas absorption Coeff label: " << m_partGas_absk_names[0] << "\n"
            << "  - The boundary condition for the absorption coeff used in the RMCRT intensity calculation is 1.0.\n";
  if (!m_do_partRadiation) {
    std::cout << "  - sigmaT4 = (sigma/M_PI) * " << m_partGas_temp_names[0] << "^4\n\n";
  } else {
    std::cout << "  - Including the particle radiation contributions \n      sumT    = ";
    for (int i = 0; i < m_nPartGasLabels; i++)
      std::cout << "(" << m_partGas_absk_names[i] << " * " << m_partGas_temp_names[i] << "^4 ) " << (i < m_nPartGasLabels - 1 ? " + " : "\n");
    std::cout << "      sumAbsk = (";
    for (int i = 0; i < m_nPartGasLabels; i++) std::cout << m_partGas_absk_names[i] << (i < m_nPartGasLabels - 1 ? " + " : ")\n");
    std::cout << "      sigmaT4 = (sigma/M_PI) * sumT/sumAbsk\n\n";
  }

  m_RMCRT->registerVariables(m_matl, m_sumAbsk_Label, m_gasTemp_Label, m_labels->d_cellTypeLabel, _src_label, m_whichAlgo == dataOnionSlim ? (int)RMCRTCommon::dataOnion
                             : m_whichAlgo == dataOnion ? (int)RMCRTCommon::dataOnion : m_whichAlgo == coarseLevel ? (int)RMCRTCommon::coarseLevel
                             : m_whichAlgo == radiometerOnly ? (int)RMCRTCommon::radiometerOnly : (int)RMCRTCommon::singleLevel);
  ProblemSpecP rmcrt_ps = m_ps->findBlock("RMCRT");
  m_RMCRT->problemSetup(m_ps, rmcrt_ps, grid);
  m_RMCRT->BC_bulletproofing(rmcrt_ps, true, false);

  // dx must get smaller as the level index increases; Arches is always computed on the finest level (323-345)
  const int maxLevels = grid->numLevels();
  m_archesLevelIndex = maxLevels - 1;
  if (maxLevels > 1) {
    Vector dx_prev = grid->getLevel(0)->dCell();
    for (int L = 1; L < maxLevels; L++) {
      const Vector dx = grid->getLevel(L)->dCell();
      if (dx.x() / dx_prev.x() > 1 || dx.y() / dx_prev.y() > 1 || dx.z() / dx_prev.z() > 1) {
        std::ostringstream warn;
        warn << "RMCRT: ERROR Level-" << L << " cell spacing is not smaller than Level-" << L - 1;
        throw ProblemSetupException(warn.str(), __FILE__, __LINE__);
      }
      dx_prev = dx;
    }
  }
}

//______________________________________________________________________ RMCRT.cc:574-660
void RMCRT_Radiation::sched_initialize(const LevelP& level, SchedulerP& sched) {
  GridP grid = m_grid;
  const int maxLevels = grid->numLevels();
  if (m_whichAlgo == dataOnion && maxLevels == 1)
    throw ProblemSetupException("ERROR:  RMCRT_radiation, there must be more than 1 level if you're using the Data Onion algorithm", __FILE__, __LINE__);
  if (m_whichAlgo == radiometerOnly && maxLevels != 1)
    throw ProblemSetupException("ERROR:  RMCRT_radiation.  The virtual radiometer only works on 1 level", __FILE__, __LINE__);

  // the Arches level only: coarser levels keep sigmaT4 / abskg as device records (Ray::coarsenAll), nothing to initialise there
  m_RMCRT->sched_initialize_VarLabel(grid->getLevel(m_archesLevelIndex), sched, m_RMCRT->d_sigmaT4Label);
  m_RMCRT->sched_initialize_VarLabel(grid->getLevel(m_archesLevelIndex), sched, m_RMCRT->d_abskgLabel);

  if (m_whichAlgo == radiometerOnly) {
    m_RMCRT->getRadiometer()->sched_initialize_VRFlux(level, sched);
    return;
  }
  for (int l = 0; l < maxLevels; l++) {
    const LevelP& myLevel = grid->getLevel(l);
    const int L_index = myLevel->getIndex();
    std::ostringstream schedName;
    schedName << "RMCRT_Radiation::sched_initialize_L-" << L_index;
    Task* tsk = new Task(schedName.str(), this, &RMCRT_Radiation::initialize);
    tsk->computesVar(RMCRTCommon::d_radiationVolqLabel);
    tsk->computesVar(RMCRTCommon::d_boundFluxLabel);
    if (L_index == m_archesLevelIndex || m_whichAlgo == coarseLevel) tsk->computesVar(_src_label);   // divQ on all levels for coarseLevel
    sched->addTask(tsk, myLevel->eachPatch(), RMCRTCommon::d_matlSet, Scheduler::ALL_GRAPHS);
  }
  // cellType on the coarse levels: BoundaryCondition::sched_cellTypeInit + Ray::sched_computeCellType in the reference (645-656);
  // here the coarse cellType is part of the device records Ray::coarsenAll derives every radiation solve
  sched_fluxInit(level, sched);
}

void RMCRT_Radiation::initialize(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw) {   // 666-725
  const Level* level = getLevel(patches);
  const int L_index = level->getIndex();
  for (int p = 0; p < patches->size(); p++) {
    const Patch* patch = patches->get(p);
    CCVariable<double> radVolq, src;
    CCVariable<Stencil7> RMCRTboundFlux;
    new_dw->allocateAndPut(radVolq, RMCRTCommon::d_radiationVolqLabel, m_matl, patch);
    radVolq.initialize(0.0);
    new_dw->allocateAndPut(RMCRTboundFlux, RMCRTCommon::d_boundFluxLabel, m_matl, patch);
    Stencil7 zero; zero.initialize(0.0);
    RMCRTboundFlux.initialize(zero);
    if (L_index == m_archesLevelIndex || m_whichAlgo == coarseLevel) {
      new_dw->allocateAndPut(src, _src_label, m_matl, patch);
      src.initialize(0.0);
    }
  }
}

//______________________________________________________________________ RMCRT.cc:1298-1353
void RMCRT_Radiation::sched_fluxInit(const LevelP& level, SchedulerP& sched) {
  if (level->getIndex() != m_archesLevelIndex)
    throw InternalError("RMCRT_Radiation::sched_fluxInit.  You cannot schedule this task on a non-arches level", __FILE__, __LINE__);
  if (m_RMCRT->d_solveBoundaryFlux) {
    Task* tsk = new Task("RMCRT_Radiation::fluxInit", this, &RMCRT_Radiation::fluxInit);
    for (auto l : {m_radFluxE_Label, m_radFluxW_Label, m_radFluxN_Label, m_radFluxS_Label, m_radFluxT_Label, m_radFluxB_Label}) tsk->computesVar(l);
    sched->addTask(tsk, level->eachPatch(), RMCRTCommon::d_matlSet, Scheduler::ALL_GRAPHS);
  }
}
void RMCRT_Radiation::fluxInit(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw) {
  for (int p = 0; p < patches->size(); p++) {
    const Patch* patch = patches->get(p);
    for (auto l : {m_radFluxE_Label, m_radFluxW_Label, m_radFluxN_Label, m_radFluxS_Label, m_radFluxT_Label, m_radFluxB_Label}) {
      CCVariable<double> q;
      new_dw->allocateAndPut(q, l, m_matl, patch);
      q.initialize(0);
    }
  }
}

//______________________________________________________________________ RMCRT.cc:353-570
void RMCRT_Radiation::sched_computeSource(const LevelP& level, SchedulerP& sched, int timeSubStep) {
  GridP grid = m_grid;
  if (timeSubStep != 0 || level->getIndex() != m_archesLevelIndex) return;   // only on RK step 0 and on the arches level
  const int maxLevels = grid->numLevels();
  const bool modifies_divQ = false;

  typedef std::vector<const VarLabel*> VarLabelVec;
  VarLabelVec fineLevelVarLabels = {m_RMCRT->d_divQLabel, m_RMCRT->d_boundFluxLabel, m_RMCRT->d_radiationVolqLabel, m_RMCRT->d_abskgLabel,
                                    m_RMCRT->d_sigmaT4Label};
  // coarse levels: abskg / sigmaT4 / cellType are device records (not DataWarehouse variables), so there is nothing to carry forward
  // there except the results of an RMCRT_coarseLevel solve
  VarLabelVec coarseResults = {m_RMCRT->d_divQLabel, m_RMCRT->d_boundFluxLabel, m_RMCRT->d_radiationVolqLabel};
  Task::WhichDW notUsed = Task::None;

  if (m_whichAlgo == dataOnion || m_whichAlgo == dataOnionSlim) {   // 397-455
    Task::WhichDW celltype_dw = Task::NewDW, sigmaT4_dw = Task::NewDW;
    const bool backoutTemp = true, modifies_abskg = false, modifies_sigmaT4 = false;
    const LevelP& fineLevel = grid->getLevel(m_archesLevelIndex);
    m_RMCRT->set_abskg_dw_perLevel(fineLevel, Task::NewDW);
    sched_sigmaT4(fineLevel, sched);
    sched_sumAbsk(fineLevel, sched);
    m_RMCRT->sched_carryForward_VarLabels(fineLevel, sched, fineLevelVarLabels);
    for (int l = maxLevels - 2; l >= 0; l--) {
      const LevelP& lev = grid->getLevel(l);
      m_RMCRT->sched_CoarsenAll(lev, sched, modifies_abskg, modifies_sigmaT4);
      // RMCRT_Radiation::sched_setBoundaryConditions (1153-1290) takes the wall values from Arches' BoundaryCondition object; this mirror
      // has the <BoundaryConditions> of the input file instead and uses Ray's own task for the coarse-level extra cells
      if (m_RMCRT->d_coarsenExtraCells == false) m_RMCRT->sched_setBoundaryConditions(lev, sched, notUsed, backoutTemp);
    }
    m_RMCRT->sched_ROI_Extents(fineLevel, sched);
    m_RMCRT->sched_rayTrace_dataOnion(fineLevel, sched, notUsed, sigmaT4_dw, celltype_dw, modifies_divQ);
    sched_stencilToDBLs(fineLevel, sched);
  }

  if (m_whichAlgo == coarseLevel) {   // 461-508
    Task::WhichDW temp_dw = Task::OldDW, sigmaT4_dw = Task::NewDW, celltype_dw = Task::NewDW;
    const bool modifies_abskg = false, modifies_sigmaT4 = false, backoutTemp = true;
    for (int l = 0; l < maxLevels; l++)
      m_RMCRT->sched_carryForward_VarLabels(grid->getLevel(l), sched, l == m_archesLevelIndex ? fineLevelVarLabels : coarseResults);
    const LevelP& fineLevel = grid->getLevel(m_archesLevelIndex);
    m_RMCRT->set_abskg_dw_perLevel(fineLevel, Task::NewDW);
    sched_sigmaT4(fineLevel, sched);
    sched_sumAbsk(fineLevel, sched);
    for (int l = maxLevels - 2; l >= 0; l--) {
      const LevelP& lev = grid->getLevel(l);
      m_RMCRT->sched_CoarsenAll(lev, sched, modifies_abskg, modifies_sigmaT4);
      if (m_RMCRT->d_coarsenExtraCells == false) m_RMCRT->sched_setBoundaryConditions(lev, sched, temp_dw, backoutTemp);
    }
    for (int l = 0; l < maxLevels - 1; l++) m_RMCRT->sched_rayTrace(grid->getLevel(l), sched, notUsed, sigmaT4_dw, celltype_dw, modifies_divQ);
    for (int l = 0; l < maxLevels; l++) m_RMCRT->sched_Refine_Q(sched, grid->getLevel(l)->eachPatch(), RMCRTCommon::d_matlSet);
    sched_stencilToDBLs(fineLevel, sched);
  }

  if (m_whichAlgo == singleLevel) {   // 513-533
    Task::WhichDW sigmaT4_dw = Task::NewDW, celltype_dw = Task::NewDW;
    const LevelP& lev = grid->getLevel(m_archesLevelIndex);
    m_RMCRT->set_abskg_dw_perLevel(lev, Task::NewDW);
    m_RMCRT->sched_carryForward_VarLabels(lev, sched, fineLevelVarLabels);
    sched_sigmaT4(lev, sched);
    sched_sumAbsk(lev, sched);
    m_RMCRT->sched_rayTrace(lev, sched, notUsed, sigmaT4_dw, celltype_dw, modifies_divQ);
    sched_stencilToDBLs(lev, sched);
  }

  if (m_whichAlgo == radiometerOnly) {   // 538-562
    Radiometer* radiometer = m_RMCRT->getRadiometer();
    Task::WhichDW sigmaT4_dw = Task::NewDW, celltype_dw = Task::NewDW;
    const LevelP& lev = grid->getLevel(m_archesLevelIndex);
    m_RMCRT->set_abskg_dw_perLevel(lev, Task::NewDW);
    m_RMCRT->sched_carryForward_VarLabels(lev, sched, {m_RMCRT->d_abskgLabel, m_RMCRT->d_sigmaT4Label, radiometer->d_VRFluxLabel, radiometer->d_VRIntensityLabel});
    sched_sigmaT4(lev, sched);
    sched_sumAbsk(lev, sched);
    radiometer->sched_radiometer(lev, sched, notUsed, sigmaT4_dw, celltype_dw);
  }
}

//______________________________________________________________________ RMCRT.cc:900-934
void RMCRT_Radiation::sched_sigmaT4(const LevelP& level, SchedulerP& sched) {
  Task* tsk = nullptr;
  Task::WhichDW oldNew_dw = Task::OldDW;
  if (m_FLT_DBL == TypeDescription::double_type) tsk = new Task("RMCRT_Radiation::sigmaT4", this, &RMCRT_Radiation::sigmaT4<double>, oldNew_dw);
  else                                           tsk = new Task("RMCRT_Radiation::sigmaT4", this, &RMCRT_Radiation::sigmaT4<float>, oldNew_dw);
  tsk->requiresVar(oldNew_dw, m_labels->d_volFractionLabel, m_gn, 0);
  for (int i = 0; i < m_nPartGasLabels; i++) {
    tsk->requiresVar(oldNew_dw, m_partGas_absk_Labels[i], m_gn, 0);
    tsk->requiresVar(oldNew_dw, m_partGas_temp_Labels[i], m_gn, 0);
  }
  tsk->computesVar(m_RMCRT->d_sigmaT4Label);
  tsk->usesDevice(true);
  sched->addTask(tsk, level->eachPatch(), RMCRTCommon::d_matlSet, TG_RMCRT);
}

// RMCRT.cc:938-1016: the per-cell arithmetic runs in arches_sigmaT4_kernel
template <class T>
void RMCRT_Radiation::sigmaT4(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw,
                              Task::WhichDW which_dw) {
  DataWarehouse* oldNew_dw = new_dw->getOtherDataWarehouse(which_dw);
  RMCRTCommon::Device& dev = m_RMCRT->devices()[0];
  for (int p = 0; p < patches->size(); p++) {
    const Patch* patch = patches->get(p);
    constCCVariable<double> gasVolFrac;
    oldNew_dw->get(gasVolFrac, m_labels->d_volFractionLabel, m_matl, patch, m_gn, 0);
    CCVariable<T> sigmaT4;
    new_dw->allocateAndPut(sigmaT4, m_RMCRT->d_sigmaT4Label, m_matl, patch);
    std::vector<constCCVariable<double>> partGas_absk(m_nPartGasLabels), partGas_temp(m_nPartGasLabels);
    std::vector<const double*> absk(m_nPartGasLabels), temp(m_nPartGasLabels);
    for (int i = 0; i < m_nPartGasLabels; i++) {
      oldNew_dw->get(partGas_absk[i], m_partGas_absk_Labels[i], m_matl, patch, m_gn, 0);
      oldNew_dw->get(partGas_temp[i], m_partGas_temp_Labels[i], m_matl, patch, m_gn, 0);
      absk[i] = partGas_absk[i].getPointer();
      temp[i] = partGas_temp[i].getPointer();
    }
    RMCRTCommon::check(dev.ctx, rmcrt_arches_sigmaT4(dev.ctx, gasVolFrac.getPointer(), absk.data(), temp.data(), m_nPartGasLabels, m_do_partRadiation ? 1 : 0,
                                                     sigmaT4.getPointer(), sizeof(T) == 4 ? 1 : 0, (int64_t)sigmaT4.cells(), nullptr), "rmcrt_arches_sigmaT4");
  }
}

//______________________________________________________________________ RMCRT.cc:1023-1056
void RMCRT_Radiation::sched_sumAbsk(const LevelP& level, SchedulerP& sched) {
  Task* tsk = nullptr;
  Task::WhichDW oldNew_dw = Task::OldDW;
  if (m_FLT_DBL == TypeDescription::double_type) tsk = new Task("RMCRT_Radiation::sumAbsk", this, &RMCRT_Radiation::sumAbsk<double>, oldNew_dw);
  else                                           tsk = new Task("RMCRT_Radiation::sumAbsk", this, &RMCRT_Radiation::sumAbsk<float>, oldNew_dw);
  tsk->requiresVar(oldNew_dw, m_labels->d_volFractionLabel, m_gn, 0);
  for (int i = 0; i < m_nPartGasLabels; i++) tsk->requiresVar(oldNew_dw, m_partGas_absk_Labels[i], m_gn, 0);
  tsk->computesVar(m_sumAbsk_Label);
  tsk->usesDevice(true);
  sched->addTask(tsk, level->eachPatch(), RMCRTCommon::d_matlSet, TG_RMCRT);
}

// RMCRT.cc:1060-1130 in arches_sumAbsk_kernel: interior gas cells sum, solid cells and boundary-face cells get the hard-wired 1.
// A patch face that is not a domain boundary has no extra cells, so "outside the patch interior" is exactly ExtraPlusEdgeCells.
template <class T>
void RMCRT_Radiation::sumAbsk(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw,
                              Task::WhichDW which_dw) {
  DataWarehouse* oldNew_dw = new_dw->getOtherDataWarehouse(which_dw);
  RMCRTCommon::Device& dev = m_RMCRT->devices()[0];
  for (int p = 0; p < patches->size(); p++) {
    const Patch* patch = patches->get(p);
    constCCVariable<double> gasVolFrac;
    oldNew_dw->get(gasVolFrac, m_labels->d_volFractionLabel, m_matl, patch, m_gn, 0);
    std::vector<constCCVariable<double>> partGas_absk(m_nPartGasLabels);
    std::vector<const double*> absk(m_nPartGasLabels);
    for (int i = 0; i < m_nPartGasLabels; i++) {
      oldNew_dw->get(partGas_absk[i], m_partGas_absk_Labels[i], m_matl, patch, m_gn, 0);
      absk[i] = partGas_absk[i].getPointer();
    }
    CCVariable<T> sumAbsk;
    new_dw->allocateAndPut(sumAbsk, m_sumAbsk_Label, m_matl, patch);
    const IntVector lo = sumAbsk.getLowIndex(), hi = sumAbsk.getHighIndex(), ilo = patch->getCellLowIndex(), ihi = patch->getCellHighIndex();
    RMCRTCommon::check(dev.ctx, rmcrt_arches_sumAbsk(dev.ctx, gasVolFrac.getPointer(), absk.data(), m_nPartGasLabels, lo.get_pointer(), hi.get_pointer(),
                                                     ilo.get_pointer(), ihi.get_pointer(), sumAbsk.getPointer(), sizeof(T) == 4 ? 1 : 0, nullptr), "rmcrt_arches_sumAbsk");
  }
}

//______________________________________________________________________ RMCRT.cc:1357-1423
void RMCRT_Radiation::sched_stencilToDBLs(const LevelP& level, SchedulerP& sched) {
  if (level->getIndex() != m_archesLevelIndex)
    throw InternalError("RMCRT_Radiation::sched_stencilToDBLs.  You cannot schedule this task on a non-arches level", __FILE__, __LINE__);
  if (m_RMCRT->d_solveBoundaryFlux) {
    Task* tsk = new Task("RMCRT_Radiation::stencilToDBLs", this, &RMCRT_Radiation::stencilToDBLs);
    tsk->requiresVar(Task::NewDW, RMCRTCommon::d_boundFluxLabel, m_gn, 0);
    for (auto l : {m_radFluxE_Label, m_radFluxW_Label, m_radFluxN_Label, m_radFluxS_Label, m_radFluxT_Label, m_radFluxB_Label}) tsk->computesVar(l);
    tsk->usesDevice(true);
    sched->addTask(tsk, level->eachPatch(), RMCRTCommon::d_matlSet, Scheduler::ALL_GRAPHS);   // no graph index in the reference (1383): every graph
  }
}
void RMCRT_Radiation::stencilToDBLs(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*, DataWarehouse* new_dw) {
  RMCRTCommon::Device& dev = m_RMCRT->devices()[0];
  for (int p = 0; p < patches->size(); ++p) {
    const Patch* patch = patches->get(p);
    constCCVariable<Stencil7> boundaryFlux;
    new_dw->get(boundaryFlux, RMCRTCommon::d_boundFluxLabel, m_matl, patch, m_gn, 0);
    CCVariable<double> East, West, North, South, Top, Bot;
    new_dw->allocateAndPut(East, m_radFluxE_Label, m_matl, patch);
    new_dw->allocateAndPut(West, m_radFluxW_Label, m_matl, patch);
    new_dw->allocateAndPut(North, m_radFluxN_Label, m_matl, patch);
    new_dw->allocateAndPut(South, m_radFluxS_Label, m_matl, patch);
    new_dw->allocateAndPut(Top, m_radFluxT_Label, m_matl, patch);
    new_dw->allocateAndPut(Bot, m_radFluxB_Label, m_matl, patch);
    RMCRTCommon::check(dev.ctx, rmcrt_stencil_to_dbls(dev.ctx, reinterpret_cast<const double*>(boundaryFlux.getPointer()), (int64_t)East.cells(), East.getPointer(),
                                                      West.getPointer(), North.getPointer(), South.getPointer(), Top.getPointer(), Bot.getPointer(), nullptr),
                       "rmcrt_stencil_to_dbls");
  }
}

}  // namespace Uintah
