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
