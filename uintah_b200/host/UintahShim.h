// UintahShim.h -- the narrow slice of Uintah's framework types that the RMCRT plugin (Ray / RMCRTCommon) touches,
// written from scratch so the host side of the B200 solve can be compiled, driven from a .ups deck and tested
// WITHOUT the Uintah tree (sus needs MPI + libxml2, neither is in this image: SURVEY.md fact 2).
//
// Only names, argument meaning and semantics follow the reference; none of its code is used.  When the plugin is
// dropped into a real Uintah tree this header is replaced by the real ones:
//   IntVector/Vector/Point   Core/Geometry/{IntVector,Vector,Point}.h
//   CCVariable/Array3        Core/Grid/Variables/{CCVariable,Array3,Array3Data}.h  (x-fastest window, Array3Data.h:93,212-220)
//   Stencil7                 Core/Grid/Variables/Stencil7.h:38-76
//   Level / Patch / Grid     Core/Grid/{Level,Patch,Grid}.h  (getCellPosition Level.cc:473-477, mapCellToCoarser 1285-1307)
//   VarLabel, Ghost, Task    Core/Grid/Variables/VarLabel.h, Core/Grid/Ghost.h, Core/Grid/Task.h:217-279
//   DataWarehouse            CCA/Ports/DataWarehouse.h (getLevel / getRegion / allocateAndPut / getModifiable / transferFrom)
//   Scheduler                CCA/Ports/Scheduler.h (addTask with a task-graph index, setNumTaskGraphs)
//   ProblemSpec              Core/ProblemSpec/ProblemSpec.h (findBlock / get / getWithDefault / require / getAttribute)
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace Uintah {

//----------------------------------------------------------------------------------------------------- exceptions
class Exception : public std::runtime_error {
 public:
  Exception(const std::string& kind, const std::string& msg, const char* file, int line)
      : std::runtime_error(kind + ": " + msg + " (" + file + ":" + std::to_string(line) + ")") {}
};
class ProblemSetupException : public Exception {
 public:
  ProblemSetupException(const std::string& m, const char* f, int l) : Exception("ProblemSetupException", m, f, l) {}
};
class InternalError : public Exception {
 public:
  InternalError(const std::string& m, const char* f, int l) : Exception("InternalError", m, f, l) {}
};

//----------------------------------------------------------------------------------------------------- geometry
class IntVector {
 public:
  IntVector() : v_{0, 0, 0} {}
  IntVector(int x, int y, int z) : v_{x, y, z} {}
  int x() const { return v_[0]; }
  int y() const { return v_[1]; }
  int z() const { return v_[2]; }
  int& operator[](int i) { return v_[i]; }
  int operator[](int i) const { return v_[i]; }
  const int* get_pointer() const { return v_; }
  IntVector operator+(const IntVector& o) const { return {v_[0] + o.v_[0], v_[1] + o.v_[1], v_[2] + o.v_[2]}; }
  IntVector operator-(const IntVector& o) const { return {v_[0] - o.v_[0], v_[1] - o.v_[1], v_[2] - o.v_[2]}; }
  IntVector operator*(const IntVector& o) const { return {v_[0] * o.v_[0], v_[1] * o.v_[1], v_[2] * o.v_[2]}; }
  bool operator==(const IntVector& o) const { return v_[0] == o.v_[0] && v_[1] == o.v_[1] && v_[2] == o.v_[2]; }
  bool operator!=(const IntVector& o) const { return !(*this == o); }
 private:
  int v_[3];
};
inline IntVector Min(const IntVector& a, const IntVector& b) { return {std::min(a.x(), b.x()), std::min(a.y(), b.y()), std::min(a.z(), b.z())}; }
inline IntVector Max(const IntVector& a, const IntVector& b) { return {std::max(a.x(), b.x()), std::max(a.y(), b.y()), std::max(a.z(), b.z())}; }
inline int Max(int a, int b, int c) { return std::max(a, std::max(b, c)); }
inline std::ostream& operator<<(std::ostream& o, const IntVector& v) { return o << "[" << v.x() << "," << v.y() << "," << v.z() << "]"; }

class Vector {
 public:
  Vector() : v_{0, 0, 0} {}
  explicit Vector(double a) : v_{a, a, a} {}
  Vector(double x, double y, double z) : v_{x, y, z} {}
  double x() const { return v_[0]; }
  double y() const { return v_[1]; }
  double z() const { return v_[2]; }
  double& operator[](int i) { return v_[i]; }
  double operator[](int i) const { return v_[i]; }
  Vector operator+(const Vector& o) const { return {v_[0] + o.v_[0], v_[1] + o.v_[1], v_[2] + o.v_[2]}; }
  Vector operator-(const Vector& o) const { return {v_[0] - o.v_[0], v_[1] - o.v_[1], v_[2] - o.v_[2]}; }
  Vector operator/(const Vector& o) const { return {v_[0] / o.v_[0], v_[1] / o.v_[1], v_[2] / o.v_[2]}; }
  Vector operator*(double s) const { return {v_[0] * s, v_[1] * s, v_[2] * s}; }
  double length() const { return std::sqrt(v_[0] * v_[0] + v_[1] * v_[1] + v_[2] * v_[2]); }
 private:
  double v_[3];
};
typedef Vector Point;
// Core/Geometry/BBox.h: axis-aligned box grown by extend()
class BBox {
 public:
  void extend(const Point& p) {
    if (!valid_) { min_ = max_ = p; valid_ = true; return; }
    for (int d = 0; d < 3; ++d) { min_[d] = std::min(min_[d], p[d]); max_[d] = std::max(max_[d], p[d]); }
  }
  Point min() const { return min_; }
  Point max() const { return max_; }
 private:
  Point min_, max_;
  bool valid_ = false;
};
inline std::ostream& operator<<(std::ostream& o, const Vector& v) { return o << "[" << v.x() << "," << v.y() << "," << v.z() << "]"; }

//----------------------------------------------------------------------------------------------------- type tags, labels
struct TypeDescription {
  enum Type { double_type, float_type, int_type, Stencil7_type, Unknown };
};

namespace Ghost {
enum GhostType { None, AroundCells };
}

// VarLabel.cc:56-122: labels are unique per name and reference counted -- create() of an existing name returns the same
// object (Ray and Radiometer both construct an RMCRTCommon and must see the same "sigmaT4" label), destroy() drops a reference.
class VarLabel {
 public:
  static VarLabel* create(const std::string& name, TypeDescription::Type type) {
    auto& reg = registry();
    auto it = reg.find(name);
    if (it != reg.end()) { ++it->second->refs_; return it->second; }
    VarLabel* l = new VarLabel(name, type);
    reg[name] = l;
    return l;
  }
  static void destroy(const VarLabel* l) {
    if (!l) return;
    VarLabel* m = const_cast<VarLabel*>(l);
    if (--m->refs_ > 0) return;
    registry().erase(m->name_);
    delete m;
  }
  const std::string& getName() const { return name_; }
  TypeDescription::Type typeDescription() const { return type_; }
  bool equals(const VarLabel* o) const { return this == o; }
 private:
  VarLabel(const std::string& n, TypeDescription::Type t) : name_(n), type_(t) {}
  static std::map<std::string, VarLabel*>& registry() { static std::map<std::string, VarLabel*> r; return r; }
  std::string name_;
  TypeDescription::Type type_;
  int refs_ = 1;
};

struct Stencil7 {   // Stencil7.h:38-76: AoS of 7 doubles; RMCRTboundFlux uses w,e,s,n,b,t for the faces and p as the "has a boundary" mark
  double w, e, s, n, b, t, p;
  double& operator[](int i) { return (&w)[i]; }
  const double& operator[](int i) const { return (&w)[i]; }
  void initialize(double v) { w = e = s = n = b = t = p = v; }
};

//----------------------------------------------------------------------------------------------------- Array3 / CCVariable
// x-fastest storage over the index window [lo,hi); copies share the data (ref-counted like Array3Data).
// Where grid-variable storage comes from (Array3Data's allocation, Array3Data.h:93-125).  Default: the C++ heap.  A back end may install
// its own pair -- the RMCRT back end installs rmcrt_host_alloc / rmcrt_host_free (page-locked, device-mapped, pooled) so that uploads run
// at full PCIe rate and the trace kernels write divQ / radiationVolq in place; in a real Uintah tree the same two calls go where
// Array3Data allocates.
struct HostAllocator {
  void* (*alloc)(size_t) = nullptr;
  void (*release)(void*) = nullptr;
  static HostAllocator& get() { static HostAllocator a; return a; }
};
template <class T>
struct HookAlloc {
  typedef T value_type;
  HookAlloc() {}
  template <class U> HookAlloc(const HookAlloc<U>&) {}
  T* allocate(size_t n) {
    HostAllocator& h = HostAllocator::get();
    if (h.alloc) {
      // a 16-byte header remembers which release function belongs to the block (the hook may be installed after older blocks were made)
      char* raw = static_cast<char*>(h.alloc(n * sizeof(T) + 16));
      if (raw) { *reinterpret_cast<void (**)(void*)>(raw) = h.release; return reinterpret_cast<T*>(raw + 16); }
    }
    char* raw = static_cast<char*>(::operator new(n * sizeof(T) + 16));
    *reinterpret_cast<void (**)(void*)>(raw) = nullptr;
    return reinterpret_cast<T*>(raw + 16);
  }
  void deallocate(T* p, size_t) {
    char* raw = reinterpret_cast<char*>(p) - 16;
    void (*rel)(void*) = *reinterpret_cast<void (**)(void*)>(raw);
    if (rel) rel(raw); else ::operator delete(raw);
  }
  template <class U> bool operator==(const HookAlloc<U>&) const { return true; }
  template <class U> bool operator!=(const HookAlloc<U>&) const { return false; }
};

template <class T>
class Array3 {
 public:
  Array3() {}
  void resize(const IntVector& lo, const IntVector& hi) {
    lo_ = lo; hi_ = hi;
    const IntVector n = hi - lo;
    data_ = std::make_shared<std::vector<T, HookAlloc<T>>>((size_t)std::max(n.x(), 0) * std::max(n.y(), 0) * std::max(n.z(), 0));
  }
  IntVector getLowIndex() const { return lo_; }
  IntVector getHighIndex() const { return hi_; }
  IntVector size() const { return hi_ - lo_; }
  bool allocated() const { return (bool)data_; }
  size_t cells() const { return data_ ? data_->size() : 0; }
  T* getPointer() { return data_ ? data_->data() : nullptr; }
  const T* getPointer() const { return data_ ? data_->data() : nullptr; }
  size_t offset(const IntVector& c) const {
    const IntVector n = hi_ - lo_;
    return (size_t)(c.x() - lo_.x()) + (size_t)n.x() * ((size_t)(c.y() - lo_.y()) + (size_t)n.y() * (size_t)(c.z() - lo_.z()));
  }
  bool contains(const IntVector& c) const {
    return c.x() >= lo_.x() && c.y() >= lo_.y() && c.z() >= lo_.z() && c.x() < hi_.x() && c.y() < hi_.y() && c.z() < hi_.z();
  }
  T& operator[](const IntVector& c) { return (*data_)[offset(c)]; }
  const T& operator[](const IntVector& c) const { return (*data_)[offset(c)]; }
  void initialize(const T& v) { if (data_) std::fill(data_->begin(), data_->end(), v); }
  // copy the overlap of src into this window
  template <class S>
  void copyRegion(const Array3<S>& src) {
    const IntVector lo = Max(lo_, src.getLowIndex()), hi = Min(hi_, src.getHighIndex());
    for (int k = lo.z(); k < hi.z(); ++k)
      for (int j = lo.y(); j < hi.y(); ++j)
        for (int i = lo.x(); i < hi.x(); ++i) (*this)[IntVector(i, j, k)] = (T)src[IntVector(i, j, k)];
  }
  void copyData(const Array3<T>& src) { copyRegion(src); }
 protected:
  IntVector lo_, hi_;
  std::shared_ptr<std::vector<T, HookAlloc<T>>> data_;
};
template <class T> class CCVariable : public Array3<T> {};
template <class T> class constCCVariable : public Array3<T> {
 public:
  constCCVariable() {}
  constCCVariable(const Array3<T>& a) : Array3<T>(a) {}
};

//----------------------------------------------------------------------------------------------------- grid
class Level;
class Grid;

class Patch {
 public:
  enum FaceType { xminus = 0, xplus, yminus, yplus, zminus, zplus, numFaces };
  Patch(int id, const Level* level, const IntVector& lo, const IntVector& hi, const IntVector& exLo, const IntVector& exHi)
      : id_(id), level_(level), lo_(lo), hi_(hi), exLo_(exLo), exHi_(exHi) {}
  int getID() const { return id_; }
  const Level* getLevel() const { return level_; }
  IntVector getCellLowIndex() const { return lo_; }          // interior cells of the patch
  IntVector getCellHighIndex() const { return hi_; }
  IntVector getExtraCellLowIndex() const { return exLo_; }   // incl. extra cells on domain faces
  IntVector getExtraCellHighIndex() const { return exHi_; }
  int getNumExtraCells() const { return (exHi_ - exLo_).x() * (exHi_ - exLo_).y() * (exHi_ - exLo_).z(); }
  void getBoundaryFaces(std::vector<FaceType>& f) const {
    f.clear();
    for (int d = 0; d < 3; ++d) {
      if (exLo_[d] < lo_[d]) f.push_back((FaceType)(2 * d));
      if (exHi_[d] > hi_[d]) f.push_back((FaceType)(2 * d + 1));
    }
  }
  // cells of the extra-cell layer on `face`, edges and corners included (Patch::ExtraPlusEdgeCells)
  void getFaceExtraCells(FaceType face, IntVector& lo, IntVector& hi) const {
    lo = exLo_; hi = exHi_;
    const int d = face / 2;
    if (face % 2 == 0) hi[d] = lo_[d]; else lo[d] = hi_[d];
  }
  Vector dCell() const;
 private:
  int id_;
  const Level* level_;
  IntVector lo_, hi_, exLo_, exHi_;
};

class PatchSubset {
 public:
  int size() const { return (int)p_.size(); }
  const Patch* get(int i) const { return p_[i]; }
  void add(const Patch* p) { p_.push_back(p); }
 private:
  std::vector<const Patch*> p_;
};
typedef PatchSubset PatchSet;   // one subset per set is all the shim needs

class MaterialSubset {
 public:
  int size() const { return (int)m_.size(); }
  int get(int i) const { return m_[i]; }
  void add(int m) { m_.push_back(m); }
 private:
  std::vector<int> m_;
};
class MaterialSet {
 public:
  void addAll(const std::vector<int>& m) { for (int x : m) s_.add(x); }
  const MaterialSubset* getUnion() const { return &s_; }
  void addReference() { ++ref_; }
  bool removeReference() { return --ref_ == 0; }
 private:
  MaterialSubset s_;
  int ref_ = 0;
};

class Level {
 public:
  Level(Grid* g, int index, const Point& anchor, const Vector& dcell, const IntVector& extra, const IntVector& rr)
      : grid_(g), index_(index), anchor_(anchor), dcell_(dcell), extra_(extra), rr_(rr) {}
  int getIndex() const { return index_; }
  Grid* getGrid() const { return grid_; }
  Vector dCell() const { return dcell_; }
  Point getAnchor() const { return anchor_; }
  IntVector getExtraCells() const { return extra_; }
  IntVector getRefinementRatio() const { return rr_; }     // to the next coarser level
  // Level.cc:473-477
  Point getCellPosition(const IntVector& c) const {
    return Point(anchor_.x() + dcell_.x() * c.x() + dcell_.x() * 0.5, anchor_.y() + dcell_.y() * c.y() + dcell_.y() * 0.5,
                 anchor_.z() + dcell_.z() * c.z() + dcell_.z() * 0.5);   // m_anchor + m_dcell * v + m_dcell * 0.5, in that order
  }
  // Level.cc:480-485: RoundDown((p - anchor) / dcell)
  IntVector getCellIndex(const Point& p) const {
    return IntVector((int)std::floor((p.x() - anchor_.x()) / dcell_.x()), (int)std::floor((p.y() - anchor_.y()) / dcell_.y()),
                     (int)std::floor((p.z() - anchor_.z()) / dcell_.z()));
  }
  // Level.cc getSpatialRange: the boxes of all patches, extra cells included
  void getSpatialRange(BBox& b) const {
    const IntVector lo = lo_ - extra_, hi = hi_ + extra_;
    b.extend(Point(anchor_.x() + dcell_.x() * lo.x(), anchor_.y() + dcell_.y() * lo.y(), anchor_.z() + dcell_.z() * lo.z()));
    b.extend(Point(anchor_.x() + dcell_.x() * hi.x(), anchor_.y() + dcell_.y() * hi.y(), anchor_.z() + dcell_.z() * hi.z()));
  }
  void findCellIndexRange(IntVector& lo, IntVector& hi) const { lo = lo_ - extra_; hi = hi_ + extra_; }
  // Level::containsCell: any cell of the level, extra cells included
  bool containsCell(const IntVector& c) const {
    const IntVector lo = lo_ - extra_, hi = hi_ + extra_;
    return c.x() >= lo.x() && c.y() >= lo.y() && c.z() >= lo.z() && c.x() < hi.x() && c.y() < hi.y() && c.z() < hi.z();
  }
  void findInteriorCellIndexRange(IntVector& lo, IntVector& hi) const { lo = lo_; hi = hi_; }
  // Level.cc:1285-1307: truncating division, negative indices shifted by (int)fmod(idx, ratio)
  IntVector mapCellToCoarser(const IntVector& c) const {
    IntVector r;
    for (int d = 0; d < 3; ++d) {
      int q = c[d] / rr_[d];
      if (c[d] < 0 && rr_[d] > 1) q += (int)std::fmod((double)c[d], (double)rr_[d]);
      r[d] = q;
    }
    return r;
  }
  bool hasCoarserLevel() const { return index_ > 0; }
  bool hasFinerLevel() const;
  const Level* getCoarserLevel() const;
  const Level* getFinerLevel() const;
  const PatchSet* eachPatch() const { return &all_; }
  int numPatches() const { return (int)patches_.size(); }
  const Patch* getPatch(int i) const { return patches_[i].get(); }
  // shim-only construction helpers
  void setInterior(const IntVector& lo, const IntVector& hi) { lo_ = lo; hi_ = hi; }
  void addPatch(std::unique_ptr<Patch> p) { all_.add(p.get()); patches_.push_back(std::move(p)); }
 private:
  Grid* grid_;
  int index_;
  Point anchor_;
  Vector dcell_;
  IntVector extra_, rr_, lo_, hi_;
  std::vector<std::unique_ptr<Patch>> patches_;
  PatchSubset all_;
};
typedef std::shared_ptr<Level> LevelP;
inline Vector Patch::dCell() const { return level_->dCell(); }

// one Dirichlet / Neumann entry of <Grid><BoundaryConditions><Face side=..><BCType label=.. var=..><value>
struct BCEntry { std::string var; double value = 0; };

class Grid {
 public:
  int numLevels() const { return (int)levels_.size(); }
  void getSpatialRange(BBox& b) const { for (auto& l : levels_) l->getSpatialRange(b); }   // Grid.cc:726-732
  const LevelP& getLevel(int i) const { return levels_[i]; }
  void addLevel(const LevelP& l) { levels_.push_back(l); }
  // face in Patch::FaceType order; returns false when the deck has no BCType for (face, label)
  bool getBC(int face, const std::string& label, BCEntry& out) const {
    auto it = bc_.find(std::make_pair(face, label));
    if (it == bc_.end()) return false;
    out = it->second;
    return true;
  }
  void setBC(int face, const std::string& label, const BCEntry& e) { bc_[std::make_pair(face, label)] = e; }
  IntVector periodic{0, 0, 0};
 private:
  std::vector<LevelP> levels_;
  std::map<std::pair<int, std::string>, BCEntry> bc_;
};
typedef std::shared_ptr<Grid> GridP;

inline bool Level::hasFinerLevel() const { return index_ < grid_->numLevels() - 1; }
inline const Level* Level::getCoarserLevel() const { return grid_->getLevel(index_ - 1).get(); }
inline const Level* Level::getFinerLevel() const { return grid_->getLevel(index_ + 1).get(); }
inline const Level* getLevel(const PatchSubset* p) { return p->get(0)->getLevel(); }

class ProcessorGroup {
 public:
  int myRank() const { return 0; }
  int nRanks() const { return 1; }
};

//----------------------------------------------------------------------------------------------------- DataWarehouse
class DataWarehouse {
 public:
  explicit DataWarehouse(int generation) : generation_(generation) {}
  int getID() const { return generation_; }
  void setTimeStep(int t) { timestep_ = t; }
  int getTimeStep() const { return timestep_; }    // the reference reads the timeStep sole-variable; the shim keeps it here

  DataWarehouse* getOtherDataWarehouse(int whichDW);   // Task::WhichDW
  void setOthers(DataWarehouse* oldDW, DataWarehouse* newDW) { old_ = oldDW; new_ = newDW; }

  bool exists(const VarLabel* l, int matl, const Patch* p) const { return vars_.count(key(l, matl, p)) != 0; }

  template <class T>
  void allocateAndPut(CCVariable<T>& v, const VarLabel* l, int matl, const Patch* p) {
    v.resize(p->getExtraCellLowIndex(), p->getExtraCellHighIndex());
    put(v, l, matl, p);
  }
  template <class T>
  void allocateTemporary(CCVariable<T>& v, const Patch* p) { v.resize(p->getExtraCellLowIndex(), p->getExtraCellHighIndex()); }
  template <class T>
  void put(const CCVariable<T>& v, const VarLabel* l, int matl, const Patch* p) {
    vars_[key(l, matl, p)] = std::make_shared<Holder<T>>(v);
  }
  template <class T>
  void getModifiable(CCVariable<T>& v, const VarLabel* l, int matl, const Patch* p) { v = fetch<T>(l, matl, p); }
  template <class T>
  void get(constCCVariable<T>& v, const VarLabel* l, int matl, const Patch* p, Ghost::GhostType = Ghost::None, int = 0) {
    v = constCCVariable<T>(fetch<T>(l, matl, p));
  }
  // whole-level gather: a fresh array over the level's cell range incl. extra cells (OnDemandDataWarehouse.cc:2423-2534)
  template <class T>
  void getLevel(constCCVariable<T>& v, const VarLabel* l, int matl, const Level* level) {
    IntVector lo, hi;
    level->findCellIndexRange(lo, hi);
    getRegion(v, l, matl, level, lo, hi);
  }
  template <class T>
  void getRegion(constCCVariable<T>& v, const VarLabel* l, int matl, const Level* level, const IntVector& lo, const IntVector& hi) {
    Array3<T> a;
    a.resize(lo, hi);
    for (int i = 0; i < level->numPatches(); ++i) {
      const Patch* p = level->getPatch(i);
      if (!exists(l, matl, p)) throw InternalError("getRegion: " + l->getName() + " missing on patch " + std::to_string(p->getID()), __FILE__, __LINE__);
      a.copyRegion(fetch<T>(l, matl, p));
    }
    v = constCCVariable<T>(a);
  }
  // new_dw->transferFrom(old_dw, label, patches, matls, replace): share the old variable (carry forward)
  void transferFrom(DataWarehouse* from, const VarLabel* l, const PatchSubset* patches, const MaterialSubset* matls) {
    for (int p = 0; p < patches->size(); ++p)
      for (int m = 0; m < matls->size(); ++m) {
        auto it = from->vars_.find(key(l, matls->get(m), patches->get(p)));
        if (it != from->vars_.end()) vars_[it->first] = it->second;
      }
  }

 private:
  struct HolderBase { virtual ~HolderBase() {} };
  template <class T> struct Holder : HolderBase { CCVariable<T> v; explicit Holder(const CCVariable<T>& x) : v(x) {} };
  typedef std::tuple<std::string, int, int> Key;
  static Key key(const VarLabel* l, int matl, const Patch* p) { return Key(l->getName(), matl, p->getID()); }
  template <class T>
  CCVariable<T>& fetch(const VarLabel* l, int matl, const Patch* p) {
    auto it = vars_.find(key(l, matl, p));
    if (it == vars_.end()) throw InternalError("DataWarehouse: unknown variable " + l->getName() + " on patch " + std::to_string(p->getID()), __FILE__, __LINE__);
    Holder<T>* h = dynamic_cast<Holder<T>*>(it->second.get());
    if (!h) throw InternalError("DataWarehouse: type mismatch for " + l->getName(), __FILE__, __LINE__);
    return h->v;
  }
  std::map<Key, std::shared_ptr<HolderBase>> vars_;
  DataWarehouse* old_ = nullptr;
  DataWarehouse* new_ = nullptr;
  int generation_;
  int timestep_ = 0;
};

//----------------------------------------------------------------------------------------------------- Task / Scheduler
class Task {
 public:
  enum WhichDW { OldDW = 0, NewDW = 1, None = -1 };
  enum DomainSpec { NormalDomain, CoarseLevel, FineLevel };
  struct Dependency {
    enum Kind { Requires, Computes, Modifies } kind;
    WhichDW dw;
    const VarLabel* label;
    Ghost::GhostType gtype;
    int numGhost;
    DomainSpec domain;
    int levelOffset;
  };
  typedef std::function<void(const ProcessorGroup*, const PatchSubset*, const MaterialSubset*, DataWarehouse*, DataWarehouse*)> Action;

  // Task(name, object, &Class::method, args...) -- the callback shape of Task::ActionNonPortable (Task.h:217-279)
  template <class C, class... Args, class... Bound>
  Task(const std::string& name, C* obj,
       void (C::*pmf)(const ProcessorGroup*, const PatchSubset*, const MaterialSubset*, DataWarehouse*, DataWarehouse*, Args...),
       Bound... bound)
      : name_(name) {
    action_ = [obj, pmf, bound...](const ProcessorGroup* pg, const PatchSubset* p, const MaterialSubset* m, DataWarehouse* o, DataWarehouse* n) {
      (obj->*pmf)(pg, p, m, o, n, bound...);
    };
  }
  const std::string& getName() const { return name_; }
  void requiresVar(WhichDW dw, const VarLabel* l, Ghost::GhostType g = Ghost::None, int n = 0) {
    deps_.push_back({Dependency::Requires, dw, l, g, n, NormalDomain, 0});
  }
  // requiresVar(dw, label, patches, CoarseLevel, offset, matls, NormalDomain, gtype, nGhost): a coarser level's variable
  void requiresVar(WhichDW dw, const VarLabel* l, const PatchSubset*, DomainSpec dom, int levelOffset, const MaterialSubset*, DomainSpec,
                   Ghost::GhostType g, int n) {
    deps_.push_back({Dependency::Requires, dw, l, g, n, dom, levelOffset});
  }
  void computesVar(const VarLabel* l) { deps_.push_back({Dependency::Computes, NewDW, l, Ghost::None, 0, NormalDomain, 0}); }
  void modifiesVar(const VarLabel* l) { deps_.push_back({Dependency::Modifies, NewDW, l, Ghost::None, 0, NormalDomain, 0}); }
  enum TaskType { Normal, Spatial };                 // Task.h: Spatial = scheduled on a subset of the level's patches
  void setType(TaskType t) { type_ = t; }
  TaskType getType() const { return type_; }
  void usesDevice(bool on, int streams = 1) { usesDevice_ = on; deviceStreams_ = streams; }
  bool usesDevice() const { return usesDevice_; }
  int deviceStreams() const { return deviceStreams_; }
  const std::vector<Dependency>& dependencies() const { return deps_; }
  void doit(const ProcessorGroup* pg, const PatchSubset* p, const MaterialSubset* m, DataWarehouse* o, DataWarehouse* n) { action_(pg, p, m, o, n); }
 private:
  std::string name_;
  Action action_;
  std::vector<Dependency> deps_;
  bool usesDevice_ = false;
  int deviceStreams_ = 1;
  TaskType type_ = Normal;
};

inline DataWarehouse* DataWarehouse::getOtherDataWarehouse(int whichDW) { return whichDW == Task::OldDW ? old_ : new_; }

// Serial scheduler: tasks of a graph run in the order they were added, each once over its whole patch set (the host code
// underneath fans the patches out over the visible GPUs).  Dependencies are recorded and checked, not used for ordering.
class Scheduler {
 public:
  struct Entry { std::unique_ptr<Task> task; const PatchSet* patches; const MaterialSet* matls; int tg; };
  enum { ALL_GRAPHS = -1 };
  void setNumTaskGraphs(int n) { nGraphs_ = n; }
  int getNumTaskGraphs() const { return nGraphs_; }
  void addTask(Task* t, const PatchSet* patches, const MaterialSet* matls, int tg = ALL_GRAPHS) {
    entries_.push_back(Entry{std::unique_ptr<Task>(t), patches, matls, tg});
  }
  const std::vector<Entry>& tasks() const { return entries_; }
  void clearTasks() { entries_.clear(); ownedSets_.clear(); }
  // spatial tasks build their own PatchSet (Radiometer::sched_radiometer); the real scheduler reference-counts it, the
  // shim takes a copy that lives as long as the task list
  const PatchSet* adopt(const PatchSet& ps) { ownedSets_.push_back(std::make_unique<PatchSet>(ps)); return ownedSets_.back().get(); }
  DataWarehouse* get_dw(int which) { return which == Task::OldDW ? old_.get() : new_.get(); }
  // OldDW <- NewDW, fresh NewDW
  void advanceDataWarehouse(int timestep) {
    old_ = new_ ? new_ : std::make_shared<DataWarehouse>(generation_++);
    new_ = std::make_shared<DataWarehouse>(generation_++);
    old_->setOthers(old_.get(), new_.get());
    new_->setOthers(old_.get(), new_.get());
    new_->setTimeStep(timestep);
  }
  // run task graph `tg` (tasks registered for every graph run in all of them)
  void execute(int tg) {
    ProcessorGroup pg;
    for (auto& e : entries_) {
      if (tg != ALL_GRAPHS && e.tg != ALL_GRAPHS && e.tg != tg) continue;   // execute(ALL_GRAPHS): the initialisation graph runs everything
      for (const auto& d : e.task->dependencies())
        if (d.kind == Task::Dependency::Requires && d.domain == Task::NormalDomain) {
          DataWarehouse* dw = get_dw(d.dw);
          for (int p = 0; p < e.patches->size(); ++p)
            if (!dw || !dw->exists(d.label, e.matls->getUnion()->get(0), e.patches->get(p)))
              throw InternalError("task " + e.task->getName() + " requires " + d.label->getName() + " which no earlier task computed", __FILE__, __LINE__);
        }
      e.task->doit(&pg, e.patches, e.matls->getUnion(), old_.get(), new_.get());
    }
  }
 private:
  std::vector<Entry> entries_;
  std::vector<std::unique_ptr<PatchSet>> ownedSets_;
  std::shared_ptr<DataWarehouse> old_, new_;
  int nGraphs_ = 1;
  int generation_ = 0;
};
typedef std::shared_ptr<Scheduler> SchedulerP;

//----------------------------------------------------------------------------------------------------- ProblemSpec
// A tolerant reader for the XML subset the RMCRT decks use: elements, attributes, text, comments, <?xml?>.
class ProblemSpec;
typedef std::shared_ptr<ProblemSpec> ProblemSpecP;

class ProblemSpec : public std::enable_shared_from_this<ProblemSpec> {
 public:
  static ProblemSpecP parseString(const std::string& xml);
  static ProblemSpecP parseFile(const std::string& path);

  const std::string& getNodeName() const { return name_; }
  ProblemSpecP getRootNode() { ProblemSpecP n = shared_from_this(); while (n->parent_.lock()) n = n->parent_.lock(); return n; }
  ProblemSpecP findBlock(const std::string& name) const {
    for (auto& c : children_) if (c->name_ == name) return c;
    return nullptr;
  }
  // next sibling with the same name (iterate <Face>, <BCType>, <Level> lists)
  ProblemSpecP findNextBlock(const std::string& name) const {
    ProblemSpecP par = parent_.lock();
    if (!par) return nullptr;
    bool seen = false;
    for (auto& c : par->children_) {
      if (c.get() == this) { seen = true; continue; }
      if (seen && c->name_ == name) return c;
    }
    return nullptr;
  }
  const std::vector<ProblemSpecP>& children() const { return children_; }
  bool getAttribute(const std::string& name, std::string& out) const {
    auto it = attrs_.find(name);
    if (it == attrs_.end()) return false;
    out = it->second;
    return true;
  }
  void getAttributes(std::map<std::string, std::string>& out) const { out = attrs_; }
  const std::string& text() const { return text_; }

  template <class T> bool get(const std::string& name, T& out) const {
    ProblemSpecP b = findBlock(name);
    if (!b) return false;
    return parseValue(b->text_, out, name);
  }
  template <class T> void getWithDefault(const std::string& name, T& out, const T& dflt) const { if (!get(name, out)) out = dflt; }
  void getWithDefault(const std::string& name, std::string& out, const char* dflt) const { if (!get(name, out)) out = dflt; }
  template <class T> void require(const std::string& name, T& out) const {
    if (!get(name, out)) throw ProblemSetupException("required input <" + name + "> is missing in <" + name_ + ">", __FILE__, __LINE__);
  }
  // <DataArchiver><save label="..."/>
  bool isLabelSaved(const std::string& label) {
    ProblemSpecP da = getRootNode()->findBlock("DataArchiver");
    if (!da) return false;
    for (auto& c : da->children_) { std::string l; if (c->name_ == "save" && c->getAttribute("label", l) && l == label) return true; }
    return false;
  }

  static std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
  }
 private:
  static bool parseValue(const std::string& t, std::string& out, const std::string&) { out = trim(t); return true; }
  static bool parseValue(const std::string& t, double& out, const std::string& n) {
    try { size_t p = 0; const std::string s = trim(t); out = std::stod(s, &p); if (p != s.size()) throw 1; } catch (...) { bad(n, t); }
    return true;
  }
  static bool parseValue(const std::string& t, int& out, const std::string& n) {
    try { size_t p = 0; const std::string s = trim(t); out = std::stoi(s, &p); if (p != s.size()) throw 1; } catch (...) { bad(n, t); }
    return true;
  }
  static bool parseValue(const std::string& t, bool& out, const std::string& n) {
    std::string s = trim(t);
    std::transform(s.begin(), s.end(), s.begin(), ::tolower);
    if (s == "true" || s == "1" || s == "t" || s == "yes") out = true;
    else if (s == "false" || s == "0" || s == "f" || s == "no") out = false;
    else bad(n, t);
    return true;
  }
  static std::array<double, 3> triple(const std::string& t, const std::string& n) {
    std::string s = trim(t);
    for (char& c : s) if (c == '[' || c == ']' || c == ',') c = ' ';
    std::istringstream is(s);
    std::array<double, 3> v;
    if (!(is >> v[0] >> v[1] >> v[2])) bad(n, t);
    return v;
  }
  static bool parseValue(const std::string& t, IntVector& out, const std::string& n) {
    auto v = triple(t, n);
    out = IntVector((int)std::lround(v[0]), (int)std::lround(v[1]), (int)std::lround(v[2]));
    return true;
  }
  static bool parseValue(const std::string& t, Vector& out, const std::string& n) { auto v = triple(t, n); out = Vector(v[0], v[1], v[2]); return true; }
  [[noreturn]] static void bad(const std::string& n, const std::string& t) {
    throw ProblemSetupException("cannot parse the value of <" + n + ">: '" + trim(t) + "'", __FILE__, __LINE__);
  }
  friend struct XmlReader;
  std::string name_, text_;
  std::map<std::string, std::string> attrs_;
  std::vector<ProblemSpecP> children_;
  std::weak_ptr<ProblemSpec> parent_;
};

struct XmlReader {
  const std::string& s;
  size_t i = 0;
  explicit XmlReader(const std::string& src) : s(src) {}
  [[noreturn]] void fail(const std::string& m) const {
    size_t line = 1 + std::count(s.begin(), s.begin() + std::min(i, s.size()), '\n');
    throw ProblemSetupException("XML parse error at line " + std::to_string(line) + ": " + m, __FILE__, __LINE__);
  }
  void skipMisc() {
    for (;;) {
      while (i < s.size() && isspace((unsigned char)s[i])) ++i;
      if (s.compare(i, 4, "<!--") == 0) { size_t e = s.find("-->", i + 4); if (e == std::string::npos) fail("unterminated comment"); i = e + 3; }
      else if (s.compare(i, 2, "<?") == 0) { size_t e = s.find("?>", i + 2); if (e == std::string::npos) fail("unterminated <? ?>"); i = e + 2; }
      else if (s.compare(i, 2, "<!") == 0) { size_t e = s.find('>', i); if (e == std::string::npos) fail("unterminated <! >"); i = e + 1; }
      else break;
    }
  }
  std::string nameTok() {
    size_t b = i;
    while (i < s.size() && (isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == '-' || s[i] == ':' || s[i] == '.')) ++i;
    if (b == i) fail("name expected");
    return s.substr(b, i - b);
  }
  ProblemSpecP element(const ProblemSpecP& parent) {
    if (s[i] != '<') fail("'<' expected");
    ++i;
    auto n = std::make_shared<ProblemSpec>();
    n->name_ = nameTok();
    n->parent_ = parent;
    for (;;) {
      while (i < s.size() && isspace((unsigned char)s[i])) ++i;
      if (i >= s.size()) fail("unterminated tag <" + n->name_ + ">");
      if (s[i] == '/') { if (s.compare(i, 2, "/>") != 0) fail("'/>' expected"); i += 2; return n; }
      if (s[i] == '>') { ++i; break; }
      const std::string an = nameTok();
      while (i < s.size() && isspace((unsigned char)s[i])) ++i;
      if (i >= s.size() || s[i] != '=') fail("'=' expected after attribute " + an);
      ++i;
      while (i < s.size() && isspace((unsigned char)s[i])) ++i;
      if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) fail("quoted attribute value expected");
      const char q = s[i++];
      size_t e = s.find(q, i);
      if (e == std::string::npos) fail("unterminated attribute value");
      n->attrs_[an] = s.substr(i, e - i);
      i = e + 1;
    }
    for (;;) {   // content
      size_t lt = s.find('<', i);
      if (lt == std::string::npos) fail("missing </" + n->name_ + ">");
      n->text_ += s.substr(i, lt - i);
      i = lt;
      if (s.compare(i, 4, "<!--") == 0) { size_t e = s.find("-->", i + 4); if (e == std::string::npos) fail("unterminated comment"); i = e + 3; continue; }
      if (s.compare(i, 2, "</") == 0) {
        i += 2;
        const std::string cn = nameTok();
        if (cn != n->name_) fail("</" + cn + "> closes <" + n->name_ + ">");
        while (i < s.size() && isspace((unsigned char)s[i])) ++i;
        if (i >= s.size() || s[i] != '>') fail("'>' expected");
        ++i;
        return n;
      }
      n->children_.push_back(element(n));
    }
  }
};

inline ProblemSpecP ProblemSpec::parseString(const std::string& xml) {
  XmlReader r(xml);
  r.skipMisc();
  if (r.i >= xml.size()) throw ProblemSetupException("empty input deck", __FILE__, __LINE__);
  return r.element(nullptr);
}

}  // namespace Uintah
