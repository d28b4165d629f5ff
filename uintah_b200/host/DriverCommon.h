// DriverCommon.h -- what the stand-alone drivers (rmcrt_test, arches_rmcrt_test) share: reading a .ups file, building the Grid from
// its <Grid> block (levels, patches, extra cells, <BoundaryConditions> side= faces) and dumping DataWarehouse variables as raw arrays.
#pragma once
#include <sys/stat.h>

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
  IntVector prevRes(0, 0, 0);
  for (ProblemSpecP level_ps = grid_ps->findBlock("Level"); level_ps; level_ps = level_ps->findNextBlock("Level"), ++index) {
    ProblemSpecP box = level_ps->findBlock("Box");
    if (!box) throw ProblemSetupException("<Level> without <Box>", __FILE__, __LINE__);
    if (box->findNextBlock("Box")) throw ProblemSetupException("multi-box levels are not supported by this driver", __FILE__, __LINE__);
    Vector lower, upper;
    IntVector res, patches(1, 1, 1), extra(0, 0, 0);
    box->require("lower", lower);
    box->require("upper", upper);
    box->require("resolution", res);
    box->getWithDefault("patches", patches, IntVector(1, 1, 1));
    box->getWithDefault("extraCells", extra, IntVector(0, 0, 0));
    Vector periodic(0, 0, 0);
    level_ps->getWithDefault("periodic", periodic, Vector(0, 0, 0));
    grid->periodic = IntVector((int)periodic.x(), (int)periodic.y(), (int)periodic.z());
    const Vector dx((upper.x() - lower.x()) / res.x(), (upper.y() - lower.y()) / res.y(), (upper.z() - lower.z()) / res.z());
    IntVector rr(1, 1, 1);
    if (index > 0) {
      for (int d = 0; d < 3; ++d) {
        if (res[d] % prevRes[d] != 0) throw ProblemSetupException("level resolution is not a multiple of the coarser level's", __FILE__, __LINE__);
        rr[d] = res[d] / prevRes[d];
      }
    }
    auto level = std::make_shared<Level>(grid.get(), index, lower, dx, extra, rr);
    level->setInterior(IntVector(0, 0, 0), res);
    int id = index * 100000;
    // patch ids in the order sus assigns them: x slowest, z fastest; uneven splits like sus' (50 cells in 3 patches: 16, 17, 17 --
    // inputs/Examples/restartUda_RT.uda/t01000/timestep.xml)
    for (int pi = 0; pi < patches.x(); ++pi)
      for (int pj = 0; pj < patches.y(); ++pj)
        for (int pk = 0; pk < patches.z(); ++pk) {
          const IntVector pidx(pi, pj, pk);
          IntVector lo, hi, exLo, exHi;
          for (int d = 0; d < 3; ++d) {
            if (patches[d] < 1 || patches[d] > res[d]) throw ProblemSetupException("<patches> must be between 1 and the resolution", __FILE__, __LINE__);
            lo[d] = (res[d] * pidx[d]) / patches[d];
            hi[d] = (res[d] * (pidx[d] + 1)) / patches[d];
            exLo[d] = lo[d] - (pidx[d] == 0 ? extra[d] : 0);
            exHi[d] = hi[d] + (pidx[d] == patches[d] - 1 ? extra[d] : 0);
          }
          level->addPatch(std::unique_ptr<Patch>(new Patch(id++, level.get(), lo, hi, exLo, exHi)));
        }
    grid->addLevel(level);
    prevRes = res;
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
