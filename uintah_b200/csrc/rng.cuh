// rng.cuh -- counter-based and reference-compatible random numbers for RMCRT ray generation.
//
//  * Philox4x32-10 (Salmon et al., SC'11), keyed (seed), counter (cell, ray, timestep, stream):
//    the production generator (<randomSeed> true).  Replaces the task-local MTRand stream of
//    Ray.cc:539 / the wall-clock seeded Kokkos XorShift pool (KokkosTools.h:77-101).
//  * MT19937 "first outputs after seed(s)": reproduces MTRand::seed(s) followed by n <= MT_MAXOUT
//    randInt() calls (Core/Math/MersenneTwister.h:217-231,253-258,339-368) without materialising the
//    624-word state -- what the reference does before EVERY ray when <randomSeed> false
//    (RMCRTCommon.cc:504-506, Ray.cc:1371-1373).
#pragma once
#include <cstdint>

namespace rmcrt {

struct u32x4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
    c = u32x4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// streams (low byte of counter word 3)
enum : uint32_t { STREAM_DIR = 1, STREAM_ORIGIN = 2, STREAM_FLUX_DIR = 3, STREAM_FLUX_ORIGIN = 4, STREAM_LHC_DIVQ = 5, STREAM_LHC_FLUX = 6,
                  STREAM_RADIOMETER = 7, STREAM_SCATTER = 8 };

// (0,1) from 32 random bits, MTRand::randDblExc() formula (MersenneTwister.h:190-197)
__host__ __device__ __forceinline__ double u01_open(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }
// [0,1], MTRand::rand() formula (MersenneTwister.h:172-179)
__host__ __device__ __forceinline__ double u01_closed(uint32_t x) { return (double)x * (1.0 / 4294967295.0); }

// Keyed pseudo-random bijection of [0,n): position x of a permutation chosen by `key` (Latin-Hyper-Cube strata when
// <randomSeed> true; stands in for the Fisher-Yates shuffle RMCRTCommon::randVector draws from the task's MTRand stream,
// RMCRTCommon.cc:998-1024).  Add / multiply-by-odd / xor-shift-right are bijections mod 2^b; values >= n are cycle-walked.
__host__ __device__ __forceinline__ uint32_t keyed_permutation(uint32_t x, uint32_t n, const u32x4& key) {
  if (n <= 1u) return 0u;
  int bits = 0;
  for (uint32_t v = n - 1u; v; v >>= 1) ++bits;
  const uint32_t m = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
  const int sh = bits > 1 ? (bits + 1) / 2 : 1;
  do {
    x = (x + key.x) & m;
    x = (x * (key.y | 1u)) & m;
    x ^= x >> sh;
    x = (x * (key.z | 1u)) & m;
    x = (x + key.w) & m;
    x ^= x >> sh;
    x = (x * 0x9E3779B1u) & m;
  } while (x >= n);
  return x;
}

constexpr int MT_MAXOUT = 8;    // 5 for a ray's direction + origin; scattering draws up to 2 more before the next reseed

// out[i] = i-th randInt() after MTRand::seed(seed), i < NOUT <= MT_MAXOUT
template <int NOUT>
__host__ __device__ __forceinline__ void mt19937_first(uint32_t seed, uint32_t out[NOUT]) {
  static_assert(NOUT >= 1 && NOUT <= MT_MAXOUT, "NOUT");
  uint32_t lo[NOUT + 1];   // state[0..NOUT]
  uint32_t hi[NOUT];       // state[397..397+NOUT-1]
  uint32_t s = seed;
  lo[0] = s;
  // initialize(): state[i] = 1812433253 * (state[i-1] ^ (state[i-1] >> 30)) + i
#pragma unroll
  for (int i = 1; i <= NOUT; ++i) { s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i; lo[i] = s; }
  for (int i = NOUT + 1; i < 397; ++i) s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i;
#pragma unroll
  for (int i = 0; i < NOUT; ++i) { s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)(397 + i); hi[i] = s; }
  // reload(): new[i] = state[i+397] ^ twist(state[i], state[i+1]) for i < 227; then tempering of randInt()
#pragma unroll
  for (int i = 0; i < NOUT; ++i) {
    const uint32_t y = (lo[i] & 0x80000000u) | (lo[i + 1] & 0x7fffffffu);
    uint32_t v = hi[i] ^ (y >> 1) ^ ((lo[i + 1] & 1u) ? 0x9908b0dfu : 0u);
    v ^= (v >> 11);
    v ^= (v << 7) & 0x9d2c5680u;
    v ^= (v << 15) & 0xefc60000u;
    out[i] = v ^ (v >> 18);
  }
}

// Random numbers of the scattering events of ONE ray (RAY_SCATTER code of RMCRTCommon::updateSumI, RMCRTCommon.cc:670-676, 790-836).
// The reference draws them from the task's MTRand inside the march: one randDblExc() for the first free path, then per event one for
// the next free path and two for the new direction (findRayDirection, which in the regression mode re-seeds the generator with
// (i+j+k)*iRay+1, iRay = -9, of the cell the event happens in).
//   mode 0  <randomSeed> true : Philox4x32-10, counter (cell or tile id, ray | draw index, timestep, STREAM_SCATTER)
//   mode 1  <randomSeed> false: the MT19937 stream the reference would be on -- the ray's own seed, `mtPos` outputs consumed
//   mode 2  FIXED_RANDOM_NUM  : every draw 0.3
struct ScatRng {
  int mode;
  uint32_t key0, key1, c0, c2, c3, ray;
  uint32_t draw;
  uint32_t mtSeed;
  int mtPos;
  __host__ __device__ __forceinline__ double open01() {
    if (mode == 2) return 0.3;
    if (mode == 1) {
      uint32_t v[MT_MAXOUT];
      mt19937_first<MT_MAXOUT>(mtSeed, v);
      const int q = mtPos < MT_MAXOUT ? mtPos : MT_MAXOUT - 1;
      ++mtPos;
      return u01_open(v[q]);
    }
    const u32x4 b = philox4x32_10(u32x4{c0, ray + ((draw >> 2) << 20), c2, c3}, key0, key1);
    const uint32_t w = (draw & 3u) == 0u ? b.x : ((draw & 3u) == 1u ? b.y : ((draw & 3u) == 2u ? b.z : b.w));
    ++draw;
    return u01_open(w);
  }
  // findRayDirection(mTwister, cur) with the default iRay = -9 (RMCRTCommon.cc:504-506)
  __host__ __device__ __forceinline__ void reseedAt(int i, int j, int k) {
    if (mode == 1) { mtSeed = (uint32_t)((i + j + k) * (-9) + 1); mtPos = 0; }
  }
};

}  // namespace rmcrt
