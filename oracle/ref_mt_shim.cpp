// C shim over the reference's OWN Core/Math/MersenneTwister.h, included from where it lies
// (-I/root/reference/src).  Built only into oracle/_ref/ (git-ignored).  Used by
// tests/test_oracle_mt.py and tests/golden/make_mt_golden.py to pin the oracle's MT19937 restatement.
#include <Core/Math/MersenneTwister.h>
#include <stdint.h>
extern "C" {
void ref_mt_u32(uint32_t seed, int n, uint32_t* out) { MTRand m; m.seed(seed); for (int i = 0; i < n; ++i) out[i] = (uint32_t)m.randInt(); }
void ref_mt_rand(uint32_t seed, int n, double* out) { MTRand m; m.seed(seed); for (int i = 0; i < n; ++i) out[i] = m.rand(); }
void ref_mt_randDblExc(uint32_t seed, int n, double* out) { MTRand m; m.seed(seed); for (int i = 0; i < n; ++i) out[i] = m.randDblExc(); }
}
