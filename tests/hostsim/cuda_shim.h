// Minimal host stand-ins for the CUDA intrinsics used by uintah_b200/csrc/{march,rng}.cuh so the SAME march
// source can be compiled with g++ and stepped against the oracle on a machine without a GPU.
// Test infrastructure only (tests/test_hostsim.py); the product never builds this.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
struct float2 { float x, y; };
struct double2 { double x, y; };
template <class T> inline T __ldg(const T* p) { return *p; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline int __double2hiint(double d) { uint64_t u; std::memcpy(&u, &d, 8); return (int)(u >> 32); }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
using std::min;
using std::max;
// fp16 <-> fp32 (IEEE binary16, round to nearest even) for the half-packed words of the blocked march
inline float half_bits_to_float(unsigned short h) {
  const unsigned sign = (unsigned)(h >> 15) << 31, ex = (h >> 10) & 31u, man = h & 1023u;
  unsigned u;
  if (ex == 0) {
    if (man == 0) u = sign;
    else { int e = -1; unsigned m = man; do { ++e; m <<= 1; } while (!(m & 1024u)); u = sign | ((unsigned)(127 - 15 - e) << 23) | ((m & 1023u) << 13); }
  } else if (ex == 31) u = sign | 0x7F800000u | (man << 13);
  else u = sign | ((ex + 112u) << 23) | (man << 13);
  float f; std::memcpy(&f, &u, 4); return f;
}
inline unsigned short float_to_half_bits(float f) {
  unsigned u; std::memcpy(&u, &f, 4);
  const unsigned sign = (u >> 16) & 0x8000u;
  const int ex = (int)((u >> 23) & 255u) - 127 + 15;
  unsigned man = u & 0x7FFFFFu;
  if (((u >> 23) & 255u) == 255u) return (unsigned short)(sign | 0x7C00u | (man ? 0x200u : 0u));
  if (ex >= 31) return (unsigned short)(sign | 0x7C00u);
  if (ex <= 0) {
    if (ex < -10) return (unsigned short)sign;
    man |= 0x800000u;
    const int shift = 14 - ex;
    unsigned r = man >> shift;
    const unsigned rem = man & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (r & 1u))) ++r;
    return (unsigned short)(sign | r);
  }
  unsigned r = ((unsigned)ex << 10) | (man >> 13);
  const unsigned rem = man & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) ++r;
  return (unsigned short)(sign | r);
}
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
