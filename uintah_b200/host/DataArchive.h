// DataArchive.h -- reader and writer of Uintah's on-disk data archive (UDA), the file format on either side of the RMCRT path
// (SURVEY 8f-3): RMCRT_Test::initializeWithUda (Examples/RMCRT_Test.cc:824-940) initialises abskg / temperature / cellType from an
// archive, and every deck's <DataArchiver> block saves divQ and friends into one.
//
// Layout (Core/DataArchive/DataArchive.cc, CCA/Components/DataArchiver/DataArchiver.cc; fixture inputs/Examples/restartUda_RT.uda):
//   <uda>/index.xml                  <Uintah_DataArchive>: <Meta> (endianness, nBits), <variables> (name, type), <timesteps>
//   <uda>/tNNNNN/timestep.xml        <Uintah_timestep>: <Time>, <Grid> with one <Level> per level (anchor, cellspacing, extraCells) and
//                                    one <Patch> each (id, lowIndex/highIndex incl. extra cells, interiorLow/HighIndex when different)
//   <uda>/tNNNNN/lL/pNNNNN.xml       <Uintah_Output>: one <Variable> per (variable, material): byte range [start,end) in pNNNNN.data
//   <uda>/tNNNNN/lL/pNNNNN.data      the arrays, x-fastest over [lowIndex,highIndex); with <compression>gzip: an nBits/8-byte
//                                    little-endian uncompressed size followed by a zlib stream
// Only cell-centred variables of int / float / double / Stencil7 are read or written: what the RMCRT components save and load.
#pragma once
#include <sys/stat.h>
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <sstream>

#include "RMCRTCommon.h"

namespace Uintah {

namespace uda_detail {
inline std::string readFile(const std::string& path, bool binary = false) {
  std::ifstream in(path, binary ? std::ios::binary : std::ios::in);
  if (!in) throw InternalError("DataArchive: cannot open " + path, __FILE__, __LINE__);
  std::stringstream ss;
  ss << in.rdbuf();
  return ss.str();
}
inline std::string decodeEntities(std::string s) {
  for (auto& kv : {std::pair<const char*, const char*>{"&lt;", "<"}, {"&gt;", ">"}, {"&amp;", "&"}})
    for (size_t p = 0; (p = s.find(kv.first, p)) != std::string::npos; p += 1) s.replace(p, strlen(kv.first), kv.second);
  return s;
}
inline std::string encodeEntities(std::string s) {
  for (size_t p = 0; (p = s.find('<', p)) != std::string::npos;) s.replace(p, 1, "&lt;");
  for (size_t p = 0; (p = s.find('>', p)) != std::string::npos;) s.replace(p, 1, "&gt;");
  return s;
}
inline IntVector getIV(const ProblemSpecP& ps, const std::string& tag, const IntVector& dflt) {
  IntVector v = dflt;
  ps->get(tag, v);
  return v;
}
template <class T> struct TypeName;
template <> struct TypeName<int> { static const char* sub() { return "int"; } };
template <> struct TypeName<float> { static const char* sub() { return "float"; } };
template <> struct TypeName<double> { static const char* sub() { return "double"; } };
template <> struct TypeName<Stencil7> { static const char* sub() { return "Stencil7"; } };
inline std::string pad5(int n) { std::ostringstream s; s << std::setw(5) << std::setfill('0') << n; return s.str(); }
}  // namespace uda_detail

//______________________________________________________________________ reader: the DataArchive calls RMCRT_Test uses
class DataArchive {
 public:
  explicit DataArchive(const std::string& filebase) : base_(filebase) {
    ProblemSpecP idx = ProblemSpec::parseString(uda_detail::readFile(base_ + "/index.xml"));
    if (idx->getNodeName() != "Uintah_DataArchive") throw InternalError("DataArchive: " + base_ + "/index.xml is not a <Uintah_DataArchive>", __FILE__, __LINE__);
    if (ProblemSpecP meta = idx->findBlock("Meta")) {
      std::string e;
      if (meta->get("endianness", e) && e != "little_endian") throw InternalError("DataArchive: only little_endian archives are supported", __FILE__, __LINE__);
      meta->get("nBits", nBits_);
    }
    if (ProblemSpecP vs = idx->findBlock("variables"))
      for (auto& v : vs->children()) {
        if (v->getNodeName() != "variable") continue;
        std::string n, t;
        v->getAttribute("name", n);
        v->getAttribute("type", t);
        varNames_.push_back(n);
        varTypes_.push_back(uda_detail::decodeEntities(t));
      }
    if (ProblemSpecP ts = idx->findBlock("timesteps"))
      for (auto& t : ts->children()) {
        if (t->getNodeName() != "timestep") continue;
        std::string href, time;
        t->getAttribute("href", href);
        t->getAttribute("time", time);
        tsIndex_.push_back(std::stoi(ProblemSpec::trim(t->text())));
        tsTime_.push_back(time.empty() ? 0.0 : std::stod(time));
        tsHref_.push_back(href);
      }
  }

  void queryTimesteps(std::vector<int>& index, std::vector<double>& times) const { index = tsIndex_; times = tsTime_; }

  // variable names, number of materials (1: these archives hold the single Arches / RMCRT material) and element sub types
  void queryVariables(std::vector<std::string>& names, std::vector<int>& num_matls, std::vector<TypeDescription::Type>& subTypes) const {
    names = varNames_;
    num_matls.assign(varNames_.size(), 1);
    subTypes.clear();
    for (auto& t : varTypes_)
      subTypes.push_back(t.find("<float>") != std::string::npos ? TypeDescription::float_type
                         : t.find("<int>") != std::string::npos ? TypeDescription::int_type : TypeDescription::double_type);
  }
  const std::string& typeName(const std::string& var) const {
    for (size_t i = 0; i < varNames_.size(); ++i) if (varNames_[i] == var) return varTypes_[i];
    throw InternalError("DataArchive: variable " + var + " is not in " + base_, __FILE__, __LINE__);
  }

  // the grid of a time step, rebuilt from timestep.xml
  GridP queryGrid(int timeIndex) {
    checkIndex(timeIndex);
    if (grids_.count(timeIndex)) return grids_[timeIndex];
    // newer archives keep <Grid> in tNNNNN/grid.xml, older ones inside timestep.xml (DataArchive.cc:431-457)
    const std::string tsPath = base_ + "/" + tsHref_[timeIndex];
    const std::string gridPath = tsPath.substr(0, tsPath.find_last_of('/') + 1) + "grid.xml";
    ProblemSpecP grid_ps;
    {
      std::ifstream gx(gridPath, std::ios::binary);
      if (gx) {
        const std::string text = uda_detail::readFile(gridPath);
        if (text.find("<Grid") == std::string::npos)
          throw InternalError("DataArchive: " + gridPath + " is a binary grid file (GRID_MAGIC_NUMBER); only the xml form is read", __FILE__, __LINE__);
        ProblemSpecP doc = ProblemSpec::parseString(text);
        grid_ps = doc->getNodeName() == "Grid" ? doc : doc->findBlock("Grid");
      }
    }
    if (!grid_ps) {
      ProblemSpecP ts = ProblemSpec::parseString(uda_detail::readFile(tsPath));
      grid_ps = ts->findBlock("Grid");
    }
    if (!grid_ps) throw InternalError("DataArchive: no <Grid> in " + tsHref_[timeIndex] + " (nor a grid.xml next to it)", __FILE__, __LINE__);
    auto grid = std::make_shared<Grid>();
    int L = 0;
    Vector prevDx(0, 0, 0);
    for (ProblemSpecP lev = grid_ps->findBlock("Level"); lev; lev = lev->findNextBlock("Level"), ++L) {
      Vector anchor(0, 0, 0), dx(1, 1, 1);
      IntVector extra(0, 0, 0);
      lev->get("anchor", anchor);
      lev->require("cellspacing", dx);
      lev->get("extraCells", extra);
      IntVector rr(1, 1, 1);
      if (L > 0) for (int d = 0; d < 3; ++d) rr[d] = (int)std::lround(prevDx[d] / dx[d]);
      auto level = std::make_shared<Level>(grid.get(), L, anchor, dx, extra, rr);
      IntVector ilo(INT_MAX, INT_MAX, INT_MAX), ihi(INT_MIN, INT_MIN, INT_MIN);
      for (ProblemSpecP p = lev->findBlock("Patch"); p; p = p->findNextBlock("Patch")) {
        int id = 0, proc = -1;
        p->require("id", id);
        p->get("proc", proc);             // the rank that wrote the patch: its data is in lL/p<proc>.{xml,data} (DataArchive.cc:2310-2314)
        procOf_[timeIndex][id] = proc;
        IntVector lo, hi;
        p->require("lowIndex", lo);
        p->require("highIndex", hi);
        const IntVector inLo = uda_detail::getIV(p, "interiorLowIndex", lo), inHi = uda_detail::getIV(p, "interiorHighIndex", hi);
        level->addPatch(std::unique_ptr<Patch>(new Patch(id, level.get(), inLo, inHi, lo, hi)));
        ilo = Min(ilo, inLo); ihi = Max(ihi, inHi);
      }
      level->setInterior(ilo, ihi);
      grid->addLevel(level);
      prevDx = dx;
    }
    grids_[timeIndex] = grid;
    return grid;
  }

  // DataArchive::query: one variable of one patch, over the patch's window incl. extra cells
  template <class T>
  void query(CCVariable<T>& var, const std::string& name, int matl, const Patch* patch, int timeIndex) {
    checkIndex(timeIndex);
    const std::string want = std::string("CCVariable<") + uda_detail::TypeName<T>::sub() + ">";
    if (typeName(name) != want) throw InternalError("DataArchive::query: " + name + " is a " + typeName(name) + ", not a " + want, __FILE__, __LINE__);
    const std::string tdir = base_ + "/" + tsHref_[timeIndex].substr(0, tsHref_[timeIndex].find_last_of('/'));
    const std::string ldir = tdir + "/l" + std::to_string(patch->getLevel()->getIndex());
    // Output files are named after the RANK that wrote them, not after the patch: lL/p<proc>.xml lists the variables of every patch that
    // rank owned (DataArchive.cc:2310-2314).  Archives without <proc> (old format): guess the file of the same index, then scan the
    // level's files (2330-2347).
    if (!grids_.count(timeIndex)) queryGrid(timeIndex);
    std::vector<std::string> candidates;
    const auto pit = procOf_[timeIndex].find(patch->getID());
    if (pit != procOf_[timeIndex].end() && pit->second >= 0) candidates.push_back(ldir + "/p" + uda_detail::pad5(pit->second) + ".xml");
    else {
      candidates.push_back(ldir + "/p" + uda_detail::pad5(patch->getID()) + ".xml");
      for (int r = 0; r < 100000; ++r) {
        const std::string f = ldir + "/p" + uda_detail::pad5(r) + ".xml";
        if (!std::ifstream(f)) { if (r > patch->getLevel()->numPatches()) break; else continue; }
        if (r != patch->getID()) candidates.push_back(f);
      }
    }
    std::string pxml;
    for (const std::string& cand : candidates) {
    if (!std::ifstream(cand)) continue;
    pxml = cand;
    ProblemSpecP out = ProblemSpec::parseString(uda_detail::readFile(pxml));
    for (auto& v : out->children()) {
      if (v->getNodeName() != "Variable") continue;
      std::string vn;
      int index = 0, pid = patch->getID();
      v->get("variable", vn);
      v->get("index", index);
      v->get("patch", pid);
      if (vn != name || index != matl || pid != patch->getID()) continue;
      double start = 0, end = 0;     // byte offsets can exceed INT_MAX
      std::string comp, file;
      v->require("start", start);
      v->require("end", end);
      v->get("compression", comp);
      v->require("filename", file);
      var.resize(patch->getExtraCellLowIndex(), patch->getExtraCellHighIndex());
      const size_t want_bytes = var.cells() * sizeof(T);
      std::ifstream in(ldir + "/" + file, std::ios::binary);
      if (!in) throw InternalError("DataArchive: cannot open " + ldir + "/" + file, __FILE__, __LINE__);
      std::vector<unsigned char> raw((size_t)(end - start));
      in.seekg((std::streamoff)start);
      in.read(reinterpret_cast<char*>(raw.data()), (std::streamsize)raw.size());
      if ((size_t)in.gcount() != raw.size()) throw InternalError("DataArchive: short read of " + name + " in " + file, __FILE__, __LINE__);
      if (comp.empty() || comp == "none") {
        if (raw.size() != want_bytes) throw InternalError("DataArchive: " + name + " has " + std::to_string(raw.size()) + " bytes, the patch needs " + std::to_string(want_bytes), __FILE__, __LINE__);
        memcpy(var.getPointer(), raw.data(), want_bytes);
      } else if (comp == "gzip") {
        const size_t hdr = (size_t)nBits_ / 8;
        unsigned long long usize = 0;
        if (raw.size() < hdr) throw InternalError("DataArchive: truncated gzip block of " + name, __FILE__, __LINE__);
        memcpy(&usize, raw.data(), hdr);
        if (usize != want_bytes) throw InternalError("DataArchive: " + name + " inflates to " + std::to_string(usize) + " bytes, the patch needs " + std::to_string(want_bytes), __FILE__, __LINE__);
        uLongf dlen = (uLongf)want_bytes;
        const int rc = uncompress(reinterpret_cast<Bytef*>(var.getPointer()), &dlen, raw.data() + hdr, (uLong)(raw.size() - hdr));
        if (rc != Z_OK || dlen != want_bytes) throw InternalError("DataArchive: zlib error " + std::to_string(rc) + " inflating " + name, __FILE__, __LINE__);
      } else {
        throw InternalError("DataArchive: unsupported <compression> " + comp, __FILE__, __LINE__);
      }
      return;
    }
    }
    throw InternalError("DataArchive::query: " + name + " (matl " + std::to_string(matl) + ", patch " + std::to_string(patch->getID()) + ") is not in " +
                        (pxml.empty() ? ldir + "/p*.xml (no such file)" : pxml), __FILE__, __LINE__);
  }

  // DataArchive::queryRegion: gather [low,high) of a level from every patch that overlaps it
  template <class T>
  void queryRegion(CCVariable<T>& var, const std::string& name, int matl, const Level* level, int timeIndex, const IntVector& low, const IntVector& high) {
    var.resize(low, high);
    std::vector<char> filled(var.cells(), 0);
    const IntVector n = high - low;
    for (int p = 0; p < level->numPatches(); ++p) {
      const Patch* patch = level->getPatch(p);
      const IntVector lo = Max(low, patch->getExtraCellLowIndex()), hi = Min(high, patch->getExtraCellHighIndex());
      if (lo.x() >= hi.x() || lo.y() >= hi.y() || lo.z() >= hi.z()) continue;
      CCVariable<T> pv;
      query(pv, name, matl, patch, timeIndex);
      for (int k = lo.z(); k < hi.z(); ++k)
        for (int j = lo.y(); j < hi.y(); ++j)
          for (int i = lo.x(); i < hi.x(); ++i) {
            const IntVector c(i, j, k);
            var[c] = pv[c];
            filled[(size_t)(i - low.x()) + (size_t)n.x() * ((size_t)(j - low.y()) + (size_t)n.y() * (size_t)(k - low.z()))] = 1;
          }
    }
    for (char f : filled) if (!f) throw InternalError("DataArchive::queryRegion: the region is not covered by the archive's patches (" + name + ")", __FILE__, __LINE__);
  }

 private:
  void checkIndex(int t) const {
    if (t < 0 || t >= (int)tsIndex_.size()) throw InternalError("DataArchive: time index " + std::to_string(t) + " out of range", __FILE__, __LINE__);
  }
  std::string base_;
  int nBits_ = 64;
  std::vector<std::string> varNames_, varTypes_, tsHref_;
  std::vector<int> tsIndex_;
  std::vector<double> tsTime_;
  std::map<int, GridP> grids_;
  std::map<int, std::map<int, int>> procOf_;      // time index -> patch id -> rank whose files hold the patch (-1: not recorded)
};

//______________________________________________________________________ writer: what <DataArchiver> produces for the saved labels
class DataArchiver {
 public:
  DataArchiver(const std::string& filebase, const Grid* grid, bool gzip = true) : base_(filebase), grid_(grid), gzip_(gzip) {
    mkdir(base_.c_str(), 0755);
  }
  void beginTimestep(int timestep, double time, double delt) {
    ts_ = timestep; time_ = time; delt_ = delt;
    tdir_ = "t" + uda_detail::pad5(timestep);
    mkdir((base_ + "/" + tdir_).c_str(), 0755);
    for (int l = 0; l < grid_->numLevels(); ++l) mkdir((base_ + "/" + tdir_ + "/l" + std::to_string(l)).c_str(), 0755);
    patchXml_.clear();
    offsets_.clear();
  }
  // var must cover the patch incl. its extra cells (the allocateAndPut window)
  template <class T>
  void save(const std::string& name, int matl, const Patch* patch, const Array3<T>& var) {
    const std::string type = std::string("CCVariable<") + uda_detail::TypeName<T>::sub() + ">";
    bool known = false;
    for (auto& v : vars_) known = known || v.first == name;
    if (!known) vars_.emplace_back(name, type);
    // this writer is ONE rank (rank 0): like sus it keeps one file pair per level and rank, lL/p00000.{xml,data}, holding all its patches
    const std::string key = "l" + std::to_string(patch->getLevel()->getIndex()) + "/p" + uda_detail::pad5(0);
    const std::string data = base_ + "/" + tdir_ + "/" + key + ".data";
    size_t& off = offsets_[key];
    std::ofstream out(data, off == 0 ? (std::ios::binary | std::ios::trunc) : (std::ios::binary | std::ios::app));
    const size_t bytes = var.cells() * sizeof(T);
    size_t written = bytes;
    if (gzip_) {
      uLongf clen = compressBound((uLong)bytes);
      std::vector<unsigned char> buf(clen);
      if (compress(buf.data(), &clen, reinterpret_cast<const Bytef*>(var.getPointer()), (uLong)bytes) != Z_OK)
        throw InternalError("DataArchiver: zlib compress failed for " + name, __FILE__, __LINE__);
      const unsigned long long usize = bytes;
      out.write(reinterpret_cast<const char*>(&usize), 8);
      out.write(reinterpret_cast<const char*>(buf.data()), (std::streamsize)clen);
      written = 8 + clen;
    } else {
      out.write(reinterpret_cast<const char*>(var.getPointer()), (std::streamsize)bytes);
    }
    std::ostringstream x;
    x << "  <Variable type=\"" << uda_detail::encodeEntities(type) << "\">\n    <variable>" << name << "</variable>\n    <index>" << matl << "</index>\n    <patch>"
      << patch->getID() << "</patch>\n    <start>" << off << "</start>\n" << (gzip_ ? "    <compression>gzip</compression>\n" : "") << "    <end>" << off + written
      << "</end>\n    <filename>p" << uda_detail::pad5(0) << ".data</filename>\n  </Variable>\n";
    patchXml_[key] += x.str();
    off += written;
  }
  void endTimestep() {
    for (auto& kv : patchXml_) {
      std::ofstream x(base_ + "/" + tdir_ + "/" + kv.first + ".xml");
      x << "<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n<Uintah_Output>\n" << kv.second << "</Uintah_Output>\n";
    }
    std::ofstream t(base_ + "/" + tdir_ + "/timestep.xml");
    t << std::setprecision(17) << "<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n<Uintah_timestep>\n  <Meta>\n    <endianness>little_endian</endianness>\n    <nBits>64</nBits>\n"
      << "    <numProcs>1</numProcs>\n  </Meta>\n  <Time>\n    <timestepNumber>" << ts_ << "</timestepNumber>\n    <currentTime>" << time_
      << "</currentTime>\n    <oldDelt>" << delt_ << "</oldDelt>\n  </Time>\n  <Grid>\n    <numLevels>" << grid_->numLevels() << "</numLevels>\n";
    auto iv = [](const IntVector& v) { std::ostringstream s; s << "[" << v.x() << ", " << v.y() << ", " << v.z() << "]"; return s.str(); };
    auto vv = [](const Vector& v) { std::ostringstream s; s << std::setprecision(17) << "[" << v.x() << ", " << v.y() << ", " << v.z() << "]"; return s.str(); };
    for (int l = 0; l < grid_->numLevels(); ++l) {
      const Level* lev = grid_->getLevel(l).get();
      IntVector lo, hi;
      lev->findCellIndexRange(lo, hi);
      t << "    <Level>\n      <numPatches>" << lev->numPatches() << "</numPatches>\n      <totalCells>" << (long long)(hi - lo).x() * (hi - lo).y() * (hi - lo).z()
        << "</totalCells>\n      <extraCells>" << iv(lev->getExtraCells()) << "</extraCells>\n      <anchor>" << vv(lev->getAnchor()) << "</anchor>\n      <id>" << l
        << "</id>\n      <cellspacing>" << vv(lev->dCell()) << "</cellspacing>\n";
      for (int p = 0; p < lev->numPatches(); ++p) {
        const Patch* pa = lev->getPatch(p);
        t << "      <Patch>\n        <id>" << pa->getID() << "</id>\n        <proc>0</proc>\n        <lowIndex>" << iv(pa->getExtraCellLowIndex())
          << "</lowIndex>\n        <highIndex>" << iv(pa->getExtraCellHighIndex()) << "</highIndex>\n";
        if (pa->getCellLowIndex() != pa->getExtraCellLowIndex()) t << "        <interiorLowIndex>" << iv(pa->getCellLowIndex()) << "</interiorLowIndex>\n";
        if (pa->getCellHighIndex() != pa->getExtraCellHighIndex()) t << "        <interiorHighIndex>" << iv(pa->getCellHighIndex()) << "</interiorHighIndex>\n";
        t << "        <totalCells>" << pa->getNumExtraCells() << "</totalCells>\n      </Patch>\n";
      }
      t << "    </Level>\n";
    }
    t << "  </Grid>\n  <Data>\n";
    for (auto& kv : patchXml_) t << "    <Datafile href=\"" << kv.first << ".xml\" proc=\"0\"/>\n";
    t << "  </Data>\n</Uintah_timestep>\n";
    steps_.emplace_back(ts_, time_, delt_);
    writeIndex();
  }

 private:
  void writeIndex() const {
    std::ofstream i(base_ + "/index.xml");
    i << std::setprecision(17) << "<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n<Uintah_DataArchive>\n  <numberOfProcessors>1</numberOfProcessors>\n  <Meta>\n"
      << "    <endianness>little_endian</endianness>\n    <nBits>64</nBits>\n  </Meta>\n  <variables>\n";
    for (auto& v : vars_) i << "    <variable type=\"" << uda_detail::encodeEntities(v.second) << "\" name=\"" << v.first << "\"/>\n";
    i << "  </variables>\n  <timesteps>\n";
    for (auto& s : steps_)
      i << "    <timestep href=\"t" << uda_detail::pad5(std::get<0>(s)) << "/timestep.xml\" time=\"" << std::get<1>(s) << "\" oldDelt=\"" << std::get<2>(s) << "\">"
        << std::get<0>(s) << "</timestep>\n";
    i << "  </timesteps>\n</Uintah_DataArchive>\n";
  }
  std::string base_, tdir_;
  const Grid* grid_;
  bool gzip_;
  int ts_ = 0;
  double time_ = 0, delt_ = 0;
  std::vector<std::pair<std::string, std::string>> vars_;
  std::vector<std::tuple<int, double, double>> steps_;
  std::map<std::string, std::string> patchXml_;
  std::map<std::string, size_t> offsets_;
};

}  // namespace Uintah
