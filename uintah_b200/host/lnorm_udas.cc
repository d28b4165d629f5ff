// lnorm_udas.cc -- L1 / L2 / Linf norms of the difference of two data archives, variable by variable: the tool the reference uses to
// compare RMCRT runs (StandAlone/tools/compute_Lnorm_udas.cc; R_Tester compares against gold-standard udas the same way).
//
//   lnorm_udas <uda1> <uda2> [--out DIR] [--timeTolerance t]
//
// Every cell-centred variable of uda1 that uda2 also holds is examined on every level and every time step (matched by position in
// the time-step list, |time1 - time2| <= timeTolerance, default 1e-5 as in the reference).  The two archives may use different patch
// layouts: each patch of uda1 is compared over its interior cells (patch->getCellIterator) with the same cells gathered from uda2.
//   L1 = sum|d| / n,  L2 = sqrt(sum d^2 / n),  Linf = max|d|        (Norms<T>::printNorms, compute_Lnorm_udas.cc:208-214)
// Stencil7 variables get one set of norms per component (printNormsS7).  With --out the norms are appended to
// DIR/L-<level>/<variable>_<matl> as "time L1 L2 Linf", the layout of the reference's Lnorm directory.
#include <cmath>
#include <cstdlib>
#include <iostream>

#include "DataArchive.h"

using namespace Uintah;

namespace Uintah {
ProblemSpecP ProblemSpec::parseFile(const std::string& path) { return parseString(uda_detail::readFile(path)); }
}

namespace {

struct Norms { double L1 = 0, L2 = 0, Linf = 0; long long n = 0; };

template <class T> inline int comps() { return 1; }
template <> inline int comps<Stencil7>() { return 7; }
template <class T> inline double comp(const T& v, int) { return (double)v; }
template <> inline double comp<Stencil7>(const Stencil7& v, int c) { return (&v.w)[c]; }

template <class T>
void compareFields(std::vector<Norms>& norms, DataArchive& da1, DataArchive& da2, const std::string& var, int matl, const Patch* patch1,
                   const Level* level2, int t1, int t2) {
  CCVariable<T> f1, f2;
  da1.query(f1, var, matl, patch1, t1);
  const IntVector lo = patch1->getCellLowIndex(), hi = patch1->getCellHighIndex();
  da2.queryRegion(f2, var, matl, level2, t2, lo, hi);
  norms.resize(comps<T>());
  for (int k = lo.z(); k < hi.z(); ++k)
    for (int j = lo.y(); j < hi.y(); ++j)
      for (int i = lo.x(); i < hi.x(); ++i) {
        const IntVector c(i, j, k);
        for (int q = 0; q < comps<T>(); ++q) {
          const double d = std::fabs(comp(f1[c], q) - comp(f2[c], q));
          norms[q].L1 += d; norms[q].L2 += d * d; norms[q].Linf = std::max(norms[q].Linf, d); norms[q].n += 1;
        }
      }
}

}  // namespace

int main(int argc, char** argv) {
  std::string u1, u2, outDir;
  double timeTol = 1e-5;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--out" && i + 1 < argc) outDir = argv[++i];
    else if (a == "--timeTolerance" && i + 1 < argc) timeTol = std::atof(argv[++i]);
    else if (u1.empty()) u1 = a;
    else if (u2.empty()) u2 = a;
    else { std::cerr << "unexpected argument " << a << "\n"; return 2; }
  }
  if (u2.empty()) { std::cerr << "usage: lnorm_udas <uda1> <uda2> [--out DIR] [--timeTolerance t]\n"; return 2; }
  try {
    DataArchive da1(u1), da2(u2);
    std::vector<std::string> vars1, vars2;
    std::vector<int> nm1, nm2;
    std::vector<TypeDescription::Type> ty1, ty2;
    da1.queryVariables(vars1, nm1, ty1);
    da2.queryVariables(vars2, nm2, ty2);
    std::vector<int> idx1, idx2;
    std::vector<double> times1, times2;
    da1.queryTimesteps(idx1, times1);
    da2.queryTimesteps(idx2, times2);
    if (vars1.size() != vars2.size()) std::cerr << "\nWARNING: The udas contain a different number of variables.  Now comparing the common set of variables.\n";
    if (!outDir.empty()) mkdir(outDir.c_str(), 0755);
    std::cout.precision(10);
    int compared = 0;
    for (size_t t = 0; t < times1.size() && t < times2.size(); ++t) {
      if (std::fabs(times1[t] - times2[t]) > timeTol) {
        std::cerr << "Output time " << times1[t] << " in " << u1 << " differs from the output time " << times2[t] << " in " << u2 << "\n";
        return 5;     // abort_uncomparable
      }
      GridP g1 = da1.queryGrid((int)t), g2 = da2.queryGrid((int)t);
      if (g1->numLevels() != g2->numLevels()) { std::cerr << "The udas have a different number of levels\n"; return 5; }
      for (size_t v = 0; v < vars1.size(); ++v) {
        const std::string& var = vars1[v];
        bool common = false;
        for (auto& o : vars2) common = common || o == var;
        const std::string type = da1.typeName(var);
        if (!common || type.compare(0, 11, "CCVariable<") != 0 || da2.typeName(var) != type) continue;
        for (int l = 0; l < g1->numLevels(); ++l) {
          const Level* lev1 = g1->getLevel(l).get();
          const Level* lev2 = g2->getLevel(l).get();
          std::vector<Norms> norms;
          bool present = true;
          try {
            for (int p = 0; p < lev1->numPatches(); ++p) {
              const Patch* patch = lev1->getPatch(p);
              if (type == "CCVariable<int>") compareFields<int>(norms, da1, da2, var, 0, patch, lev2, (int)t, (int)t);
              else if (type == "CCVariable<float>") compareFields<float>(norms, da1, da2, var, 0, patch, lev2, (int)t, (int)t);
              else if (type == "CCVariable<double>") compareFields<double>(norms, da1, da2, var, 0, patch, lev2, (int)t, (int)t);
              else if (type == "CCVariable<Stencil7>") compareFields<Stencil7>(norms, da1, da2, var, 0, patch, lev2, (int)t, (int)t);
              else present = false;
            }
          } catch (const InternalError& e) {
            if (std::string(e.what()).find("is not in") == std::string::npos) throw;
            present = false;     // the variable was not saved on this level
          }
          if (!present || norms.empty()) continue;
          ++compared;
          for (size_t q = 0; q < norms.size(); ++q) {
            Norms n = norms[q];
            n.L1 /= (double)n.n;
            n.L2 = std::sqrt(n.L2 / (double)n.n);
            static const char* s7[7] = {".w", ".e", ".s", ".n", ".b", ".t", ".p"};
            const std::string name = var + (norms.size() > 1 ? s7[q] : "");
            std::cout << name << " L-" << l << " time " << times1[t] << " \t norms: L1 " << n.L1 << " L2: " << n.L2 << " Linf: " << n.Linf
                      << " n_cells: " << n.n << std::endl;
            if (!outDir.empty()) {
              const std::string ldir = outDir + "/L-" + std::to_string(l);
              mkdir(ldir.c_str(), 0755);
              std::ofstream out(ldir + "/" + name + "_0", std::ios::app);
              out.setf(std::ios::scientific, std::ios::floatfield);
              out.precision(10);
              out << times1[t] << " " << n.L1 << " " << n.L2 << " " << n.Linf << "\n";
            }
          }
        }
      }
    }
    if (compared == 0) { std::cerr << "lnorm_udas: the archives have no cell-centred variable in common\n"; return 5; }
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "lnorm_udas: caught exception: " << e.what() << std::endl;
    return 1;
  }
}
