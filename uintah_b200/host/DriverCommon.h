// DriverCommon.h -- what the stand-alone drivers (rmcrt_test, arches_rmcrt_test) share: reading a .ups file, building the Grid from
// its <Grid> block (levels, patches, extra cells, <BoundaryConditions> side= faces) and dumping DataWarehouse variables as raw arrays.
#pragma once
#include <sys/stat.h>

#include <climits>
#include <cmath>
#include <fstream>
#include <iostream>
#include <sstream>

#include "RMCRTCommon.h"

namespace Uintah {
inline ProblemSpecP ProblemSpec::parseFile(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw ProblemSetupException("cannot open input deck " + path, __FILE__, __LINE__);
  std::stringstream ss;
  ss << in.rdbuf();
  return parseString(ss.str());
}

//______________________________________________________________________ <Grid>
inline GridP buildGrid(const ProblemSpecP& root) {
  ProblemSpecP grid_ps = root->findBlock("Grid");
  if (!grid_ps) throw ProblemSetupException("no <Grid> block", __FILE__, __LINE__);
  auto grid = std::make_shared<Grid>();
  static const char* sides[6] = {"x-", "x+", "y-", "y+", "z-", "z+"};
  if (ProblemSpecP bc_ps = grid_ps->findBlock("BoundaryConditions")) {
    for (ProblemSpecP face = bc_ps->findBlock("Face"); face; face = face->findNextBlock("Face")) {
      std::string side;
      if (!face->getAttribute("side", side)) throw ProblemSetupException("<Face> without side (only side= faces are supported)", __FILE__, __LINE__);
      int f = -1;
      for (int i = 0; i < 6; ++i) if (side == sides[i]) f = i;
      if (f < 0) throw ProblemSetupException("unknown face side " + side, __FILE__, __LINE__);
      for (ProblemSpecP bc = face->findBlock("BCType"); bc; bc = bc->findNextBlock("BCType")) {
        BCEntry e;
        std::string label;
        bc->getAttribute("label", label);
        bc->getAttribute("var", e.var);
        bc->get("value", e.value);
        grid->setBC(f, label, e);
      }
    }
  }
  int index = 0;
  Vector prevDx(0, 0, 0);
  for (ProblemSpecP level_ps = grid_ps->findBlock("Level"); level_ps; level_ps = level_ps->findNextBlock("Level"), ++index) {
    // a level is one or more <Box>es of the same cell size (RMCRT_+Domain*.ups: seven boxes forming a 3-D plus); the level's index space
    // is anchored at the lowest corner of all boxes
    struct BoxSpec { Vector lower, upper; IntVector res, patches; };
    std::vector<BoxSpec> boxes;
    IntVector extra(0, 0, 0);
    for (ProblemSpecP box = level_ps->findBlock("Box"); box; box = box->findNextBlock("Box")) {
      BoxSpec bs;
      box->require("lower", bs.lower);
      box->require("upper", bs.upper);
      box->require("resolution", bs.res);
      box->getWithDefault("patches", bs.patches, IntVector(1, 1, 1));
      IntVector ec(0, 0, 0);
      box->getWithDefault("extraCells", ec, IntVector(0, 0, 0));
      if (!boxes.empty() && ec != extra) throw ProblemSetupException("every <Box> of a level must have the same <extraCells>", __FILE__, __LINE__);
      extra = ec;
      boxes.push_back(bs);
    }
    if (boxes.empty()) throw ProblemSetupException("<Level> without <Box>", __FILE__, __LINE__);
    Vector periodic(0, 0, 0);
    level_ps->getWithDefault("periodic", periodic, Vector(0, 0, 0));
    grid->periodic = IntVector((int)periodic.x(), (int)periodic.y(), (int)periodic.z());
    const BoxSpec& b0 = boxes[0];
    const Vector dx((b0.upper.x() - b0.lower.x()) / b0.res.x(), (b0.upper.y() - b0.lower.y()) / b0.res.y(), (b0.upper.z() - b0.lower.z()) / b0.res.z());
    Vector anchor = b0.lower;
    for (auto& bs : boxes)
      for (int d = 0; d < 3; ++d) {
        anchor[d] = std::min(anchor[d], bs.lower[d]);
        const double bdx = (bs.upper[d] - bs.lower[d]) / bs.res[d];
        if (std::fabs(bdx - dx[d]) > 1e-9 * dx[d]) throw ProblemSetupException("the <Box>es of a level must have the same cell size", __FILE__, __LINE__);
      }
    IntVector rr(1, 1, 1);
    if (index > 0) {
      for (int d = 0; d < 3; ++d) {
        const double ratio = prevDx[d] / dx[d];
        rr[d] = (int)std::lround(ratio);
        if (rr[d] < 1 || std::fabs(ratio - rr[d]) > 1e-9 * ratio) throw ProblemSetupException("level resolution is not a multiple of the coarser level's", __FILE__, __LINE__);
      }
    }
    auto level = std::make_shared<Level>(grid.get(), index, anchor, dx, extra, rr);
    // pass 1: patch interiors.  Patch ids in the order sus assigns them (box by box; x slowest, z fastest); uneven splits like sus'
    // (50 cells in 3 patches: 16, 17, 17 -- inputs/Examples/restartUda_RT.uda/t01000/timestep.xml)
    std::vector<std::pair<IntVector, IntVector>> interiors;
    IntVector ilo(INT_MAX, INT_MAX, INT_MAX), ihi(INT_MIN, INT_MIN, INT_MIN);
    for (auto& bs : boxes) {
      IntVector off;
      for (int d = 0; d < 3; ++d) {
        const double o = (bs.lower[d] - anchor[d]) / dx[d];
        off[d] = (int)std::lround(o);
        if (std::fabs(o - off[d]) > 1e-6) throw ProblemSetupException("a <Box> does not start on a cell boundary of its level", __FILE__, __LINE__);
        if (bs.patches[d] < 1 || bs.patches[d] > bs.res[d]) throw ProblemSetupException("<patches> must be between 1 and the resolution", __FILE__, __LINE__);
      }
      for (int pi = 0; pi < bs.patches.x(); ++pi)
        for (int pj = 0; pj < bs.patches.y(); ++pj)
          for (int pk = 0; pk < bs.patches.z(); ++pk) {
            const IntVector pidx(pi, pj, pk);
            IntVector lo, hi;
            for (int d = 0; d < 3; ++d) {
              lo[d] = off[d] + (bs.res[d] * pidx[d]) / bs.patches[d];
              hi[d] = off[d] + (bs.res[d] * (pidx[d] + 1)) / bs.patches[d];
            }
            interiors.emplace_back(lo, hi);
            ilo = Min(ilo, lo); ihi = Max(ihi, hi);
          }
    }
    level->setInterior(ilo, ihi);
    // pass 2: a patch face carries extra cells iff no other patch lies across it (the edge of the computational domain)
    auto covered = [&](const IntVector& c) {
      for (auto& in : interiors) {
        bool inside = true;
        for (int d = 0; d < 3; ++d) inside = inside && c[d] >= in.first[d] && c[d] < in.second[d];
        if (inside) return true;
      }
      return false;
    };
    int id = index * 100000;
    for (auto& in : interiors) {
      IntVector exLo = in.first, exHi = in.second;
      for (int d = 0; d < 3; ++d)
        for (int side = 0; side < 2; ++side) {
          const int a1 = (d + 1) % 3, a2 = (d + 2) % 3;
          int nCov = 0, nProbe = 0;
          for (int u = 0; u < 3; ++u)
            for (int v = 0; v < 3; ++v) {     // corners, edge mid points and the centre of the layer just outside the face
              IntVector c;
              c[d] = side == 0 ? in.first[d] - 1 : in.second[d];
              c[a1] = u == 0 ? in.first[a1] : (u == 1 ? (in.first[a1] + in.second[a1] - 1) / 2 : in.second[a1] - 1);
              c[a2] = v == 0 ? in.first[a2] : (v == 1 ? (in.first[a2] + in.second[a2] - 1) / 2 : in.second[a2] - 1);
              nCov += covered(c) ? 1 : 0;
              ++nProbe;
            }
          if (nCov != 0 && nCov != nProbe) throw ProblemSetupException("a patch face is only partly covered by neighbouring patches (not supported)", __FILE__, __LINE__);
          if (nCov == 0) { if (side == 0) exLo[d] -= extra[d]; else exHi[d] += extra[d]; }
        }
      level->addPatch(std::unique_ptr<Patch>(new Patch(id++, level.get(), in.first, in.second, exLo, exHi)));
    }
    grid->addLevel(level);
    prevDx = dx;
  }
  if (grid->numLevels() == 0) throw ProblemSetupException("no <Level> in <Grid>", __FILE__, __LINE__);
  return grid;
}


template <class T>
inline void dumpVar(const std::string& dir, const std::string& name, const char* dtype, int comps, DataWarehouse* dw, const VarLabel* label,
             const Level* level, std::ostream& meta, bool& first) {
  constCCVariable<T> v;
  dw->getLevel(v, label, 0, level);
  const std::string file = name + ".bin";
  std::ofstream out(dir + "/" + file, std::ios::binary);
  out.write(reinterpret_cast<const char*>(v.getPointer()), (std::streamsize)(v.cells() * sizeof(T)));
  const IntVector lo = v.getLowIndex(), hi = v.getHighIndex();
  meta << (first ? "" : ",\n") << "  \"" << name << "\": {\"file\": \"" << file << "\", \"dtype\": \"" << dtype << "\", \"components\": " << comps
       << ", \"lo\": [" << lo.x() << "," << lo.y() << "," << lo.z() << "], \"hi\": [" << hi.x() << "," << hi.y() << "," << hi.z() << "]}";
  first = false;
}

}  // namespace Uintah
