// RMCRTCommon.cc -- see RMCRTCommon.h.  Reference behaviour followed: RMCRT/RMCRTCommon.cc (line ranges cited per method).
#include "RMCRTCommon.h"
#include <cstdlib>

#include <cuda_runtime_api.h>

#include <cmath>
#include <cstring>
#include <iostream>
#include <sstream>

namespace Uintah {

// statics shared by Ray and Radiometer in the reference (RMCRTCommon.cc:62-91)
double RMCRTCommon::d_threshold = 0.01;
double RMCRTCommon::d_sigma = 5.67051e-8;
double RMCRTCommon::d_sigmaScat = 0;
double RMCRTCommon::d_maxRayLength = 1e308;
bool RMCRTCommon::d_isSeedRandom = true;
bool RMCRTCommon::d_allowReflect = true;
int RMCRTCommon::d_destructorCount = 0;
int RMCRTCommon::d_matl = 0;
int RMCRTCommon::d_whichAlgo = RMCRTCommon::singleLevel;
MaterialSet* RMCRTCommon::d_matlSet = nullptr;
std::string RMCRTCommon::d_abskgBC_tag;
std::vector<IntVector> RMCRTCommon::d_dbgCells;
std::map<std::string, Task::WhichDW> RMCRTCommon::d_abskg_dw;
const VarLabel* RMCRTCommon::d_sigmaT4Label = nullptr;
const VarLabel* RMCRTCommon::d_abskgLabel = nullptr;
const VarLabel* RMCRTCommon::d_divQLabel = nullptr;
const VarLabel* RMCRTCommon::d_boundFluxLabel = nullptr;
const VarLabel* RMCRTCommon::d_radiationVolqLabel = nullptr;
const VarLabel* RMCRTCommon::d_abskgSigmaT4CellTypeLabel = nullptr;
const VarLabel* RMCRTCommon::d_compTempLabel = nullptr;
const VarLabel* RMCRTCommon::d_compAbskgLabel = nullptr;
const VarLabel* RMCRTCommon::d_cellTypeLabel = nullptr;
int RMCRTCommon::s_maxDevices = 0;

//______________________________________________________________________
// RMCRTCommon.cc:97-122: the labels local to RMCRT.  "sigmaT4" holds sigma T^4 / pi in the algorithm's precision.
RMCRTCommon::RMCRTCommon(TypeDescription::Type FLT_DBL) : d_FLT_DBL(FLT_DBL) {
  d_sigmaT4Label = VarLabel::create("sigmaT4", FLT_DBL);
  d_abskgSigmaT4CellTypeLabel = VarLabel::create("abskgSigmaT4CellType", TypeDescription::double_type);
  d_boundFluxLabel = VarLabel::create("RMCRTboundFlux", TypeDescription::Stencil7_type);
  d_radiationVolqLabel = VarLabel::create("radiationVolq", TypeDescription::double_type);
  d_flowCell = -1;
  d_destructorCount = 0;
  // grid variables allocated from here on live in page-locked, device-mapped, pooled host memory (include/rmcrt_b200.h: rmcrt_host_alloc);
  // without a CUDA device the allocator returns NULL and the stand-in Array3 falls back to the heap (parse-only runs)
  if (!std::getenv("RMCRT_PAGEABLE_HOST")) {      // A/B switch: the staged path through pageable CCVariable buffers
    HostAllocator::get().alloc = rmcrt_host_alloc;
    HostAllocator::get().release = rmcrt_host_free;
  }
}

// RMCRTCommon.cc:128-152: Ray and Radiometer share the statics, so only the first destructor frees them
RMCRTCommon::~RMCRTCommon() {
  releaseDevices();
  if (d_destructorCount == 0) {
    VarLabel::destroy(d_sigmaT4Label);
    VarLabel::destroy(d_abskgSigmaT4CellTypeLabel);
    VarLabel::destroy(d_boundFluxLabel);
    VarLabel::destroy(d_radiationVolqLabel);
    if (d_abskgLabel && d_compAbskgLabel && d_abskgLabel != d_compAbskgLabel) VarLabel::destroy(d_abskgLabel);
    d_abskgLabel = nullptr;
    if (d_matlSet && d_matlSet->removeReference()) { delete d_matlSet; d_matlSet = nullptr; }
  }
  d_destructorCount++;
}

//______________________________________________________________________
// RMCRTCommon.cc:157-203
void RMCRTCommon::registerVariables(int matlIndex, const VarLabel* abskg, const VarLabel* temperature, const VarLabel* celltype,
                                    const VarLabel* divQ, const int whichAlgo) {
  d_matl = matlIndex;
  d_compAbskgLabel = abskg;
  d_compTempLabel = temperature;
  d_cellTypeLabel = celltype;
  d_divQLabel = divQ;
  d_whichAlgo = whichAlgo;
  d_abskgBC_tag = d_compAbskgLabel->getName();   // the BC specs name the component's variable
  // float RMCRT with a double abskg: RMCRT keeps its own float copy "abskgRMCRT" (188-191)
  if (d_FLT_DBL == TypeDescription::float_type && abskg->typeDescription() == TypeDescription::double_type)
    d_abskgLabel = VarLabel::create("abskgRMCRT", TypeDescription::float_type);
  else
    d_abskgLabel = d_compAbskgLabel;
  if (!d_matlSet) {
    d_matlSet = new MaterialSet();
    d_matlSet->addAll(std::vector<int>{matlIndex});
    d_matlSet->addReference();
  }
}

//______________________________________________________________________
// back end plumbing
void RMCRTCommon::check(rmcrt_ctx* ctx, int rc, const char* what) {
  if (rc == RMCRT_OK) return;
  const char* m = rmcrt_last_error(ctx);
  std::ostringstream msg;
  msg << "RMCRT (B200 back end) " << what << " failed [" << rc << "]: " << (m ? m : "");
  if (rc == RMCRT_ERR_INVALID) throw ProblemSetupException(msg.str(), __FILE__, __LINE__);
  throw InternalError(msg.str(), __FILE__, __LINE__);
}

void RMCRTCommon::fillParams(rmcrt_params& p) const {
  rmcrt_default_params(&p);
  p.threshold = d_threshold;
  p.sigma = d_sigma;
  p.allowReflect = d_allowReflect ? 1 : 0;
  p.isSeedRandom = d_isSeedRandom ? 1 : 0;
  p.flowCell = d_flowCell;
  // the reference marches in FP64 whatever T is; the production solve marches in FP32 (DESIGN.md "precision"),
  // <RMCRT type="double"> decks in deterministic regression mode get the FP64 march so they can be compared to round-off
  p.marchPrecision = (d_FLT_DBL == TypeDescription::double_type && !d_isSeedRandom) ? 1 : 0;
}

std::vector<RMCRTCommon::Device>& RMCRTCommon::devices() {
  if (d_backend != this) return d_backend->devices();
  std::lock_guard<std::mutex> lock(d_deviceMutex);
  if (d_devices.empty()) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
      throw InternalError(std::string("RMCRT: no CUDA device available (") + cudaGetErrorString(e) +
                          "); the B200 RMCRT solve has no CPU fallback", __FILE__, __LINE__);
    if (s_maxDevices > 0 && n > s_maxDevices) n = s_maxDevices;
    rmcrt_params p;
    fillParams(p);
    for (int d = 0; d < n; ++d) {
      Device dev;
      dev.ordinal = d;
      int rc = rmcrt_create(d, &p, &dev.ctx);
      if (rc != RMCRT_OK) check(nullptr, rc, "rmcrt_create");
      d_devices.push_back(dev);
    }
  }
  return d_devices;
}

rmcrt_level_desc RMCRTCommon::describe(const Level* level) {
  rmcrt_level_desc d;
  IntVector lo, hi, ilo, ihi;
  level->findCellIndexRange(lo, hi);
  level->findInteriorCellIndexRange(ilo, ihi);
  const Vector dx = level->dCell();
  const Point anchor = level->getAnchor();
  const IntVector rr = level->hasCoarserLevel() ? level->getRefinementRatio() : IntVector(1, 1, 1);
  for (int k = 0; k < 3; ++k) {
    d.lo[k] = lo[k]; d.hi[k] = hi[k];
    d.interiorLo[k] = ilo[k]; d.interiorHi[k] = ihi[k];
    d.dx[k] = dx[k]; d.anchor[k] = anchor[k];
    d.refineRatio[k] = rr[k];
  }
  return d;
}

// the three whole-level gathers of Ray.cc:549-553 followed by the H2D + pack that replaces the GPU DataWarehouse staging
template <class T>
void RMCRTCommon::uploadLevel(const Level* level, DataWarehouse* abskg_dw, DataWarehouse* sigmaT4_dw, DataWarehouse* celltype_dw, long long stamp) {
  if (d_backend != this) { d_backend->uploadLevel<T>(level, abskg_dw, sigmaT4_dw, celltype_dw, stamp); return; }
  const int L = level->getIndex();
  auto& devs = devices();
  bool needed = false;
  for (auto& dev : devs) {
    auto it = dev.levelStamp.find(L);
    if (it == dev.levelStamp.end() || it->second != stamp) needed = true;
  }
  if (!needed) return;
  constCCVariable<T> abskg, sigmaT4OverPi;
  constCCVariable<int> celltype;
  abskg_dw->getLevel(abskg, d_abskgLabel, d_matl, level);
  sigmaT4_dw->getLevel(sigmaT4OverPi, d_sigmaT4Label, d_matl, level);
  celltype_dw->getLevel(celltype, d_cellTypeLabel, d_matl, level);
  const rmcrt_level_desc desc = describe(level);
  rmcrt_params p;
  fillParams(p);
  std::vector<rmcrt_ctx*> ctxs;
  for (auto& dev : devs) {
    check(dev.ctx, rmcrt_set_params(dev.ctx, &p), "rmcrt_set_params");
    check(dev.ctx, rmcrt_define_level(dev.ctx, L, &desc, recordKind(p)), "rmcrt_define_level");
    ctxs.push_back(dev.ctx);
  }
  // every device needs the whole level (Ray.cc:479-481): each uploads and packs its z-slab, one NCCL all-gather replicates the records
  check(ctxs[0], rmcrt_replicate_level(ctxs.data(), (int)ctxs.size(), L, abskg.getPointer(), sigmaT4OverPi.getPointer(), celltype.getPointer(),
                                       sizeof(T) == 4 ? 1 : 0, desc.lo, desc.hi), "rmcrt_replicate_level");
  for (auto& dev : devs) dev.levelStamp[L] = stamp;
}
template void RMCRTCommon::uploadLevel<double>(const Level*, DataWarehouse*, DataWarehouse*, DataWarehouse*, long long);
template void RMCRTCommon::uploadLevel<float>(const Level*, DataWarehouse*, DataWarehouse*, DataWarehouse*, long long);

void RMCRTCommon::releaseDevices() {
  std::lock_guard<std::mutex> lock(d_deviceMutex);
  for (auto& d : d_devices) rmcrt_destroy(d.ctx);
  d_devices.clear();
}

//______________________________________________________________________
// RMCRTCommon.cc:279-300
void RMCRTCommon::sched_sigmaT4(const LevelP& level, SchedulerP& sched, Task::WhichDW temp_dw, const bool includeEC) {
  std::string taskname = "RMCRTCommon::sigmaT4";
  Task* tsk = nullptr;
  if (d_FLT_DBL == TypeDescription::double_type) tsk = new Task(taskname, this, &RMCRTCommon::sigmaT4<double>, temp_dw, includeEC);
  else                                           tsk = new Task(taskname, this, &RMCRTCommon::sigmaT4<float>, temp_dw, includeEC);
  tsk->requiresVar(temp_dw, d_compTempLabel, d_gn, 0);
  tsk->computesVar(d_sigmaT4Label);
  tsk->usesDevice(true);
  sched->addTask(tsk, level->eachPatch(), d_matlSet, TG_RMCRT);
}

// RMCRTCommon.cc:304-343; the arithmetic runs in sigmaT4_kernel through rmcrt_sigmaT4 (H2D, kernel, D2H)
template <class T>
void RMCRTCommon::sigmaT4(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*,
                          DataWarehouse* new_dw, Task::WhichDW which_temp_dw, const bool /*includeEC*/) {
  Device& dev = devices()[0];
  for (int p = 0; p < patches->size(); ++p) {
    const Patch* patch = patches->get(p);
    DataWarehouse* temp_dw = new_dw->getOtherDataWarehouse(which_temp_dw);
    constCCVariable<double> temp;
    CCVariable<T> sigmaT4;
    temp_dw->get(temp, d_compTempLabel, d_matl, patch, Ghost::None, 0);
    new_dw->allocateAndPut(sigmaT4, d_sigmaT4Label, d_matl, patch);
    // the patch variable always covers the extra cells (includeEC == true in every caller); interior-only callers
    // simply ignore the extra-cell values
    if (temp.cells() != sigmaT4.cells()) throw InternalError("sigmaT4: temperature and sigmaT4 windows differ", __FILE__, __LINE__);
    check(dev.ctx, rmcrt_sigmaT4(dev.ctx, temp.getPointer(), sigmaT4.getPointer(), sizeof(T) == 4 ? 1 : 0,
                                 (int64_t)temp.cells(), nullptr), "rmcrt_sigmaT4");
  }
}
template void RMCRTCommon::sigmaT4<double>(const ProcessorGroup*, const PatchSubset*, const MaterialSubset*, DataWarehouse*, DataWarehouse*, Task::WhichDW, const bool);
template void RMCRTCommon::sigmaT4<float>(const ProcessorGroup*, const PatchSubset*, const MaterialSubset*, DataWarehouse*, DataWarehouse*, Task::WhichDW, const bool);

//______________________________________________________________________
// RMCRTCommon.cc:345-376
void RMCRTCommon::sched_initialize_VarLabel(const LevelP& level, SchedulerP& sched, const VarLabel* label) {
  Task* tsk = nullptr;
  switch (label->typeDescription()) {
    case TypeDescription::double_type: tsk = new Task("RMCRTCommon::initialize_VarLabel", this, &RMCRTCommon::initialize_VarLabel<double>, label); break;
    case TypeDescription::float_type:  tsk = new Task("RMCRTCommon::initialize_VarLabel", this, &RMCRTCommon::initialize_VarLabel<float>, label); break;
    default: throw InternalError("RMCRTCommon::sched_initialize_VarLabel: (" + label->getName() + ") could not find task for this var type", __FILE__, __LINE__);
  }
  tsk->computesVar(label);
  sched->addTask(tsk, level->eachPatch(), d_matlSet);
}
template <class T>
void RMCRTCommon::initialize_VarLabel(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*,
                                      DataWarehouse* new_dw, const VarLabel* label) {
  for (int p = 0; p < patches->size(); ++p) {
    CCVariable<T> var;
    new_dw->allocateAndPut(var, label, d_matl, patches->get(p));
    var.initialize((T)0);
  }
}

//______________________________________________________________________
// RMCRTCommon.cc:211-237: only needed when RMCRT is float and the component's abskg is double
void RMCRTCommon::sched_DoubleToFloat(const LevelP& level, SchedulerP& sched, Task::WhichDW /*notUsed*/) {
  const VarLabel* abskg = d_compAbskgLabel;
  if (abskg->typeDescription() == TypeDescription::double_type && d_FLT_DBL == TypeDescription::float_type) {
    Task::WhichDW myDW = get_abskg_whichDW(level->getIndex(), d_compAbskgLabel);
    Task* tsk = new Task("RMCRTCommon::DoubleToFloat", this, &RMCRTCommon::DoubleToFloat, myDW);
    tsk->requiresVar(myDW, d_compAbskgLabel, d_gn, 0);
    tsk->computesVar(d_abskgLabel);
    sched->addTask(tsk, level->eachPatch(), d_matlSet, TG_RMCRT);
  }
}
// RMCRTCommon.cc:241-274, including its non-finite bullet proofing.  A cast of one array per radiation solve: host glue.
void RMCRTCommon::DoubleToFloat(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset*, DataWarehouse*,
                                DataWarehouse* new_dw, Task::WhichDW which_dw) {
  for (int p = 0; p < patches->size(); ++p) {
    const Patch* patch = patches->get(p);
    constCCVariable<double> abskg_D;
    CCVariable<float> abskg_F;
    DataWarehouse* myDW = new_dw->getOtherDataWarehouse(which_dw);
    myDW->get(abskg_D, d_compAbskgLabel, d_matl, patch, Ghost::None, 0);
    new_dw->allocateAndPut(abskg_F, d_abskgLabel, d_matl, patch);
    const double* s = abskg_D.getPointer();
    float* d = abskg_F.getPointer();
    for (size_t i = 0; i < abskg_D.cells(); ++i) {
      d[i] = (float)s[i];
      if (std::isinf(d[i]) || std::isnan(d[i])) {
        std::ostringstream warn;
        warn << "RMCRTCommon::DoubleToFloat A non-physical abskg detected (" << d[i] << ") on patch " << patch->getID();
        throw InternalError(warn.str(), __FILE__, __LINE__);
      }
    }
  }
}

//______________________________________________________________________
// RMCRTCommon.cc:378-414: kept so callers written for the slim variant still schedule; the packed {abskg|wall, sigmaT4/pi}
// record is built on the device by rmcrt_upload_level when the trace task first touches a level.
void RMCRTCommon::sched_combineAbskgSigmaT4CellType(const LevelP&, SchedulerP&, Task::WhichDW, const bool) {}

//______________________________________________________________________
// RMCRTCommon.cc:909-974
void RMCRTCommon::sched_CarryForward_FineLevelLabels(const LevelP& level, SchedulerP& sched) {
  sched_carryForward_VarLabels(level, sched, {d_divQLabel, d_boundFluxLabel, d_radiationVolqLabel, d_sigmaT4Label});
}
void RMCRTCommon::sched_carryForward_VarLabels(const LevelP& level, SchedulerP& sched, const std::vector<const VarLabel*> varLabels) {
  Task* tsk = new Task("RMCRTCommon::carryForward_VarLabels", this, &RMCRTCommon::carryForward_VarLabels, varLabels);
  for (auto l : varLabels) {
    tsk->requiresVar(Task::OldDW, l, d_gn, 0);
    tsk->computesVar(l);
  }
  sched->addTask(tsk, level->eachPatch(), d_matlSet, TG_CARRY_FORWARD);
}
void RMCRTCommon::sched_CarryForward_Var(const LevelP& level, SchedulerP& sched, const VarLabel* variable, const int tg_num) {
  Task* tsk = new Task("RMCRTCommon::carryForward_Var", this, &RMCRTCommon::carryForward_VarLabels, std::vector<const VarLabel*>{variable});
  tsk->requiresVar(Task::OldDW, variable, d_gn, 0);
  tsk->computesVar(variable);
  sched->addTask(tsk, level->eachPatch(), d_matlSet, tg_num);
}
void RMCRTCommon::carryForward_VarLabels(const ProcessorGroup*, const PatchSubset* patches, const MaterialSubset* matls,
                                         DataWarehouse* old_dw, DataWarehouse* new_dw, const std::vector<const VarLabel*> varLabels) {
  for (auto l : varLabels) new_dw->transferFrom(old_dw, l, patches, matls);
}

//______________________________________________________________________
// RMCRTCommon.cc:1033-1110: abskg lives in which_dw on the fine level and in NewDW on every coarser one
void RMCRTCommon::set_abskg_dw_perLevel(const LevelP& fineLevel, Task::WhichDW fineLevel_abskg_dw) {
  const int maxLevels = fineLevel->getGrid()->numLevels();
  d_abskg_dw.clear();
  for (int L = 0; L < maxLevels; ++L) {
    std::ostringstream key;
    key << d_compAbskgLabel->getName() << "_L-" << L;       // the component's abskg
    d_abskg_dw[key.str()] = (L == fineLevel->getIndex()) ? fineLevel_abskg_dw : Task::NewDW;
    std::ostringstream key1;
    key1 << d_abskgLabel->getName() << "_L-" << L;          // RMCRT's own abskg is always computed into the NewDW
    d_abskg_dw[key1.str()] = Task::NewDW;
  }
}
DataWarehouse* RMCRTCommon::get_abskg_dw(const int L, const VarLabel* label, DataWarehouse* new_dw) {
  return new_dw->getOtherDataWarehouse(get_abskg_whichDW(L, label));
}
Task::WhichDW RMCRTCommon::get_abskg_whichDW(const int L, const VarLabel* label) {
  std::ostringstream key;
  key << label->getName() << "_L-" << L;
  auto it = d_abskg_dw.find(key.str());
  if (it == d_abskg_dw.end())
    throw InternalError("RMCRTCommon::get_abskg_whichDW: set_abskg_dw_perLevel was not called for " + key.str(), __FILE__, __LINE__);
  return it->second;
}

bool RMCRTCommon::isDbgCell(IntVector me) {
  for (auto& c : d_dbgCells) if (c == me) return true;
  return false;
}

}  // namespace Uintah
