// indep.cuh -- the march of the DEFAULT sampling (rayCoherence 0): every (cell, ray) has its own direction and in-cell origin (Philox keyed
// on (cell, ray, timestep), the reference's independent rays: Ray.cc:539,757-771), so the 32 lanes of a warp gather from 32 unrelated
// places and nothing of the lockstep machinery of march_leg_pair / _quad applies.
//
// What binds such a march on B200 (ncu): on the level's linear records every ray-cell step is one 32-B sector
// request from L1 to L2 -- the SM's miss-request path (one request per clock) is 96 % busy -- issued by warps whose lanes finish at very
// different steps (13.4 of 32 lanes active).  What is done about it:
//
//  * BLOCKED HALF-PACKED RECORDS (BlkGeom).  The level is kept a second time as 2 x 2 x 2 cell blocks of 4-B half-packed records (the
//    Rec4 word of march.cuh): one block = one 32-B sector.  A ray visits two cells per block on average (mean chord of a 2-cube over the
//    mean chord of a cell), so the second visit is an L1 hit on the sector the first one fetched.
//      - march_leg_blk / BlkRay / march_single_level_blk / march_onion2_blk (PRODUCTION, blockedMarch 2 and 3): the PTX leg loop of
//        march.cuh with plain 4-byte loads and coordinate-free blocked addressing (one XOR per step toggles the stepped axis' delta);
//        divq_sorted_kernel (rmcrt_b200.cu) sorts the rays of a tile by expected length so that the lanes of a warp end together.
//      - blk_fetch / IndepRay / indep_step (blockedMarch 1, measured slower, kept for A/B): one 256-bit load per block entered into a
//        per-lane slot of shared memory, and PER-LANE RAY REFILL -- a lane walks through the rays of its cell on its own and starts the
//        next one when one ends (ray set-up batched over >= REFILL waiting lanes).
//
// Per-ray arithmetic is that of march_single_level_fast (offset-form tau, one ex2 per step, integral per leg scaled by fs at the wall,
// reflections without recomputing tMax -- RMCRTCommon.cc:607-904), on the fp16 properties of the blocked copy.
#pragma once
#include "march.cuh"

namespace rmcrt {

struct BlkGeom {
  const uint32_t* blk;   // [block][8]: word (x&1) | (y&1)<<1 | (z&1)<<2 of block (x>>1, y>>1, z>>1), x-fastest block order; 32-B aligned
  int bx, bxy;           // blocks per row, per plane (one guard block of wall records on every side of the level)
  int lo[3];             // cell index of block-local coordinate 0 (= level lo - 2)
};

#if defined(__CUDACC__)
// A lane's slot in shared memory: two 16-B chunks, chunk c of thread t at byte (c * 128 + t) * 16 of the block's slot array (128 threads).
// The position of a cell inside its 2 x 2 x 2 block is kept directly as the byte offset of its word in the slot:
constexpr unsigned OX = 4u, OY = 8u, OZ = 2048u;     // x&1 -> +4 B, y&1 -> +8 B, z&1 -> the second chunk
constexpr int ONEG = 12;                             // bits (OX|OY|OZ) << ONEG of the same word: the direction is negative on that axis

// one 2 x 2 x 2 block (32 B, one sector) -> the calling thread's slot (slot = 32-bit shared-window address of its first chunk)
__device__ __forceinline__ void blk_fetch(const uint32_t* __restrict__ blk, int B, uint32_t slot) {
  uint32_t v0, v1, v2, v3, v4, v5, v6, v7;
  const uint32_t* p = blk + (size_t)B * 8;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3), "=r"(v4), "=r"(v5), "=r"(v6), "=r"(v7) : "l"(p));
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(slot), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
  asm volatile("st.shared.v4.b32 [%0+2048], {%1,%2,%3,%4};" :: "r"(slot), "r"(v4), "r"(v5), "r"(v6), "r"(v7) : "memory");
}
__device__ __forceinline__ uint32_t blk_word(uint32_t slot, unsigned o) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(slot + (o & (OX | OY | OZ))) : "memory");
  return v;
}

// state of one lane's current ray
struct IndepRay {
  float tx, ty, tz;          // tMax, scaled by log2(e)
  float ddx, ddy, ddz;       // tDelta
  int B;                     // block index of the current cell
  int sbx, sby, sbz;         // signed block strides of the three axes (flip on reflection)
  unsigned o;                // OX|OY|OZ bits: position of the current cell inside its block; the same bits << ONEG: direction negative
  float A, k, s, ePrev, leg, tau;
  float fs, sum, nsteps;
};

// step into the neighbour cell across the face of axis `obit` (OX, OY or OZ) in the ray's direction; returns the raw record word
__device__ __forceinline__ uint32_t indep_enter(IndepRay& r, unsigned obit, int stepB, const uint32_t* __restrict__ blk, uint32_t slot) {
  r.o ^= obit;
  // the block changes when a positive step wraps the position bit to 0 or a negative step wraps it to 1
  if (((r.o ^ (r.o >> ONEG)) & obit) == 0u) {
    r.B += stepB;
    blk_fetch(blk, r.B, slot);
  }
  return blk_word(slot, r.o);
}

__device__ __forceinline__ void indep_decode(IndepRay& r, uint32_t word) {
  r.k = fabsf(__half2float(__ushort_as_half((unsigned short)(word >> 16))));
  r.s = __half2float(__ushort_as_half((unsigned short)(word & 0xFFFFu)));
}

// one DDA step of a live ray.  Returns false when the ray has ended (its contribution is then in r.sum, its steps in r.nsteps).
__device__ __forceinline__ bool indep_step(IndepRay& r, const uint32_t* __restrict__ blk, uint32_t slot, float threshold,
                                           bool allowReflect, bool skipZero) {
  // axis with the smallest tMax, strict '<' tie order of RMCRTCommon.cc:696-708 (dda_peek)
  const bool a = r.tx < r.ty;
  const float m = fminf(r.tx, r.ty);
  const bool nz_ = m < r.tz;
  const bool px = a && nz_, py = !a && nz_;
  const float t = fminf(m, r.tz);
  r.tx = px ? r.tx + r.ddx : r.tx;
  r.ty = py ? r.ty + r.ddy : r.ty;
  r.tz = nz_ ? r.tz : r.tz + r.ddz;
  const unsigned obit = px ? OX : (py ? OY : OZ);
  const int stepB = px ? r.sbx : (py ? r.sby : r.sbz);
  uint32_t word = indep_enter(r, obit, stepB, blk, slot);
  // integ_cross
  r.tau = fmaf(r.k, t, r.A);
  const float e = Num<float>::attenuation(r.tau);
  r.leg = fmaf(r.s, r.ePrev - e, r.leg);
  r.ePrev = e;
  r.nsteps += 1.0f;
  indep_decode(r, word);
  r.A = fmaf(-r.k, t, r.tau);
  if ((int)word >= 0) return true;                         // still in a flow cell
  // ---- a non-flow cell was entered: wall emission, threshold test, reflection (RMCRTCommon.cc:846-902; march_single_level_fast)
  r.sum = fmaf(r.leg, r.fs, r.sum);
  r.leg = 0.0f;
  const float eps = r.k > 1.0f ? 1.0f : r.k;
  r.sum = fmaf(eps * r.s, r.ePrev, r.sum);                 // not weighted by fs
  const float intensity = r.ePrev * r.fs;
  if (!allowReflect || !(intensity > threshold)) return false;
  r.fs *= (1.0f - r.k);
  if (skipZero && r.fs == 0.0f) return false;
  // reflect: back into the cell the ray came from, stepping direction of that axis flipped, tMax NOT recomputed
  r.o ^= obit << ONEG;
  if (px) r.sbx = -r.sbx; else if (py) r.sby = -r.sby; else r.sbz = -r.sbz;
  word = indep_enter(r, obit, -stepB, blk, slot);
  indep_decode(r, word);
  r.A = fmaf(-r.k, t, r.tau);
  return true;
}

// start a ray in the origin cell (block B0, position o0): dir never exactly zero (nz<>), rayDx = origin offset from the cell's low corner
__device__ __forceinline__ void indep_start(IndepRay& r, int B0, unsigned o0, float dx, float dy, float dz, float rx, float ry, float rz,
                                            float cx, float cy, float cz, const BlkGeom& g, uint32_t slot) {
  typedef Num<float> N;
  const bool posx = !sign_bit(dx), posy = !sign_bit(dy), posz = !sign_bit(dz);
  const float ix = N::rcp(dx), iy = N::rcp(dy), iz = N::rcp(dz);
  r.tx = ((posx ? cx : 0.0f) - rx) * ix * N::LEN_SCALE;
  r.ty = ((posy ? cy : 0.0f) - ry) * iy * N::LEN_SCALE;
  r.tz = ((posz ? cz : 0.0f) - rz) * iz * N::LEN_SCALE;
  r.ddx = fabsf(ix) * cx * N::LEN_SCALE; r.ddy = fabsf(iy) * cy * N::LEN_SCALE; r.ddz = fabsf(iz) * cz * N::LEN_SCALE;
  r.sbx = posx ? 1 : -1; r.sby = posy ? g.bx : -g.bx; r.sbz = posz ? g.bxy : -g.bxy;
  r.o = o0 | (posx ? 0u : OX << ONEG) | (posy ? 0u : OY << ONEG) | (posz ? 0u : OZ << ONEG);
  r.B = B0;
  blk_fetch(g.blk, B0, slot);
  indep_decode(r, blk_word(slot, o0));
  r.A = 0.0f; r.ePrev = 1.0f; r.leg = 0.0f; r.tau = 0.0f;
  r.fs = 1.0f; r.sum = 0.0f; r.nsteps = 0.0f;
}
#endif   // __CUDACC__ (the slot march)

//------------------------------------------------------------------------------------------------------
// blockedMarch = 2: the one-cell-per-lane PTX leg loop of march.cuh (march_leg) on the blocked copy, with PLAIN 4-byte loads.
// Nothing is cached per lane: the second cell a ray visits inside a 2 x 2 x 2 block (one 32-B sector) is an L1 hit on the sector its
// first visit fetched, so the L1 -> L2 miss requests per ray-cell step halve while the per-step instruction count grows by three.
//
// Blocked addressing without coordinates: the word index of a cell is 8 * block + (x&1 | (y&1)<<1 | (z&1)<<2).  Along one axis the
// index changes alternately by a and b (x: 1 / 7, y: 2 / 8 nbx - 2, z: 4 / 8 nbx nby - 4; negated for a negative direction), so every axis
// carries its CURRENT signed delta, and a step on that axis toggles it with one XOR: (a ^ b) == ((-a) ^ (-b)) for these pairs, a constant
// of the level (6, tgy, tgz).  A reflection steps back by the previous delta d and goes on with d - sign(d) (a + b).
//------------------------------------------------------------------------------------------------------
struct BlkStride {
  int dlx, dly, dlz;     // current signed word-index deltas of the three axes
};
#if defined(__CUDA_ARCH__)
#define RB_PEEK(J, IB)                                      \
  "min.f32 t" #J ", %0, %1, %2;\n\t"                        \
  "setp.neu.f32 n" #J ", t" #J ", %2;\n\t"                  \
  "setp.lt.and.f32 x" #J ", %0, %1, n" #J ";\n\t"           \
  "setp.ge.and.f32 y" #J ", %0, %1, n" #J ";\n\t"           \
  "selp.s32 i" #J ", %13, %15, x" #J ";\n\t"                \
  "@y" #J " mov.s32 i" #J ", %14;\n\t"                      \
  "add.s32 i" #J ", i" #J ", " IB ";\n\t"                   \
  "mad.wide.s32 ad, i" #J ", 4, %16;\n\t" "ld.global.nc.b32 k" #J ", [ad];\n\t"
#define RB_COMMIT(J)                                        \
  "@x" #J " add.f32 %0, %0, %17;\n\t"                       \
  "@x" #J " xor.b32 %13, %13, 6;\n\t"                       \
  "@y" #J " add.f32 %1, %1, %18;\n\t"                       \
  "@y" #J " xor.b32 %14, %14, %20;\n\t"                     \
  "@!n" #J " add.f32 %2, %2, %19;\n\t"                      \
  "@!n" #J " xor.b32 %15, %15, %21;\n\t"
#define RB_INTEG(P, C)                                      \
  "fma.rn.f32 ta, a" #P ", t" #C ", %4;\n\t"                \
  "neg.f32 d, ta;\n\t"                                      \
  "ex2.approx.ftz.f32 e" #C ", d;\n\t"                      \
  "sub.f32 d, e" #P ", e" #C ";\n\t"                        \
  "fma.rn.f32 %8, s" #P ", d, %8;\n\t"                      \
  "add.f32 %9, %9, 0f3F800000;\n\t"                         \
  "mov.b32 {hl, hh}, k" #C ";\n\t" "cvt.f32.f16 s" #C ", hl;\n\t" "cvt.f32.f16 a" #C ", hh;\n\t" "abs.f32 a" #C ", a" #C ";\n\t" \
  "neg.f32 d, a" #C ";\n\t"                                 \
  "fma.rn.f32 %4, d, t" #C ", ta;\n\t"                      \
  "setp.lt.s32 w, k" #C ", 0;\n\t"                          \
  "@w bra EXIT" #C ";\n\t"
#define RB_EXIT(C)                                          \
  "EXIT" #C ":\n\t"                                         \
  "mov.s32 %3, i" #C ";\n\t"                                \
  "mov.f32 %10, t" #C ";\n\t"                               \
  "selp.s32 %11, 1, 2, y" #C ";\n\t"                        \
  "@x" #C " mov.s32 %11, 0;\n\t"                            \
  "mov.f32 %5, a" #C ";\n\t"                                \
  "mov.f32 %6, s" #C ";\n\t"                                \
  "mov.f32 %7, e" #C ";\n\t"                                \
  "bra DONE;\n\t"
// same three-slot rotation as march_leg; the speculative request of step n+1 uses the deltas as they are after the commit of step n
__device__ __forceinline__ LegOut march_leg_blk(Dda& d, Integ& q, BlkStride& b, const uint32_t* __restrict__ blk, float ddx, float ddy,
                                                float ddz, int tgy, int tgz) {
  LegOut o;
  asm volatile(
      "{\n\t"
      ".reg .pred x0, y0, n0, x1, y1, n1, x2, y2, n2, w;\n\t"
      ".reg .f32 d, ta, t0, t1, t2, s0, s1, s2, e0, e1, e2, a0, a1, a2;\n\t"
      ".reg .b32 k0, k1, k2;\n\t"
      ".reg .b16 hl, hh;\n\t"
      ".reg .s32 i0, i1, i2;\n\t"
      ".reg .s64 ad;\n\t"
      "mov.f32 a2, %5;\n\t"
      "mov.f32 s2, %6;\n\t"
      "mov.f32 e2, %7;\n\t"
      RB_PEEK(0, "%3")
      "LOOP:\n\t"
      RB_COMMIT(0) RB_PEEK(1, "i0") RB_INTEG(2, 0)
      RB_COMMIT(1) RB_PEEK(2, "i1") RB_INTEG(0, 1)
      RB_COMMIT(2) RB_PEEK(0, "i2") RB_INTEG(1, 2)
      "bra LOOP;\n\t"
      RB_EXIT(0) RB_EXIT(1) RB_EXIT(2)
      "DONE:\n\t"
      "mov.f32 %12, ta;\n\t"
      "}"
      : "+f"(d.tx), "+f"(d.ty), "+f"(d.tz), "+r"(d.idx), "+f"(q.A), "+f"(q.k), "+f"(q.s), "+f"(q.ePrev), "+f"(q.leg),
        "+f"(q.nsteps), "=f"(o.t), "=r"(o.axis), "=f"(q.tau), "+r"(b.dlx), "+r"(b.dly), "+r"(b.dlz)
      : "l"(blk), "f"(ddx), "f"(ddy), "f"(ddz), "r"(tgy), "r"(tgz)
      : "memory");
  return o;
}
#undef RB_PEEK
#undef RB_COMMIT
#undef RB_INTEG
#undef RB_EXIT
#else
// the same leg statement by statement in C++ (what tests/hostsim steps against the oracle; the device loop additionally requests the
// word of step n + 1 before the wall test of step n)
__device__ __forceinline__ void blk_decode(Integ& q, uint32_t word);
__device__ __forceinline__ LegOut march_leg_blk(Dda& d, Integ& q, BlkStride& b, const uint32_t* __restrict__ blk, float ddx, float ddy,
                                                float ddz, int tgy, int tgz) {
  for (;;) {
    const bool a = d.tx < d.ty;
    const float m = fminf(d.tx, d.ty);
    const bool nz_ = m < d.tz;                 // false: z is the stepping axis (strict '<' tie order of RMCRTCommon.cc:696-708)
    const bool px = a && nz_, py = !a && nz_;
    const float t = fminf(m, d.tz);
    d.idx += px ? b.dlx : (py ? b.dly : b.dlz);
    if (px) { d.tx += ddx; b.dlx ^= 6; } else if (py) { d.ty += ddy; b.dly ^= tgy; } else { d.tz += ddz; b.dlz ^= tgz; }
    const uint32_t word = blk[d.idx];
    q.tau = fmaf(q.k, t, q.A);
    const float e = Num<float>::attenuation(q.tau);
    q.leg = fmaf(q.s, q.ePrev - e, q.leg);
    q.ePrev = e;
    q.nsteps += 1.0f;
    blk_decode(q, word);
    q.A = fmaf(-q.k, t, q.tau);
    if ((int)word < 0) return LegOut{t, px ? 0 : (py ? 1 : 2)};
  }
}
#endif

// deltas of a ray standing on position bits `par` with the given direction signs: positive direction: a on an even position, b = S - a on
// an odd one; negative: -a on an odd position, -b on an even one
__device__ __forceinline__ BlkStride blk_stride_init(unsigned par, bool posx, bool posy, bool posz, int sy8, int sz8) {
  BlkStride b;
  b.dlx = posx ? ((par & 1u) ? 7 : 1) : ((par & 1u) ? -1 : -7);
  b.dly = posy ? ((par & 2u) ? sy8 - 2 : 2) : ((par & 2u) ? -2 : 2 - sy8);
  b.dlz = posz ? ((par & 4u) ? sz8 - 4 : 4) : ((par & 4u) ? -4 : 4 - sz8);
  return b;
}

__device__ __forceinline__ void blk_decode(Integ& q, uint32_t word) {
  q.k = fabsf(half_bits_to_float((unsigned short)(word >> 16)));
  q.s = half_bits_to_float((unsigned short)(word & 0xFFFFu));
}

// One ray of the single-level march on the blocked copy, as a resumable state: blk_ray_start + blk_ray_legs(all legs) is
// march_single_level_fast (RMCRTCommon.cc:607-904); divq_sorted_kernel marches the first leg, parks the rays that go on after their wall
// (48 B of state each) and finishes them in a second, re-sorted pass.
struct BlkRay {
  Dda d;                 // tMax (scaled by log2 e), word index of the current cell
  BlkStride b;
  Integ q;
  float sum, fs;
};
// w0 = word index of the origin cell, par = its position bits (x&1 | (y&1)<<1 | (z&1)<<2), sy8 = 8 * blocks per row, sz8 = 8 * blocks per
// plane; tDelta returned in ddx, ddy, ddz
__device__ __forceinline__ void blk_ray_start(BlkRay& r, const uint32_t* __restrict__ blk, int w0, unsigned par, int sy8, int sz8, float dx,
                                              float dy, float dz, float rx, float ry, float rz, float cx, float cy, float cz, float& ddx,
                                              float& ddy, float& ddz) {
  typedef Num<float> N;
  const bool posx = !sign_bit(dx), posy = !sign_bit(dy), posz = !sign_bit(dz);
  r.b = blk_stride_init(par, posx, posy, posz, sy8, sz8);
  const float ix = N::rcp(dx), iy = N::rcp(dy), iz = N::rcp(dz);
  r.d.tx = ((posx ? cx : 0.0f) - rx) * ix * N::LEN_SCALE;
  r.d.ty = ((posy ? cy : 0.0f) - ry) * iy * N::LEN_SCALE;
  r.d.tz = ((posz ? cz : 0.0f) - rz) * iz * N::LEN_SCALE;
  ddx = fabsf(ix) * cx * N::LEN_SCALE; ddy = fabsf(iy) * cy * N::LEN_SCALE; ddz = fabsf(iz) * cz * N::LEN_SCALE;
  r.d.idx = w0;
  blk_decode(r.q, __ldg(blk + w0));
  r.q.A = 0.0f; r.q.ePrev = 1.0f; r.q.nsteps = 0.0f; r.q.tau = 0.0f; r.q.leg = 0.0f;
  r.sum = 0.0f; r.fs = 1.0f;
}
// march up to maxLegs legs; true = the ray has ended (its result is r.sum, its steps r.q.nsteps), false = it stands reflected in the cell
// it came from, ready for its next leg
__device__ __forceinline__ bool blk_ray_legs(BlkRay& r, const uint32_t* __restrict__ blk, float ddx, float ddy, float ddz, int sy8, int sz8,
                                             float threshold, bool allowReflect, bool skipZero, int maxLegs) {
  const int tgy = 2 ^ (sy8 - 2), tgz = 4 ^ (sz8 - 4);
  Integ& q = r.q;
  for (int leg = 1;; ++leg) {
    q.leg = 0.0f;
    const LegOut w = march_leg_blk(r.d, q, r.b, blk, ddx, ddy, ddz, tgy, tgz);
    r.sum = fmaf(q.leg, r.fs, r.sum);
    const float eps = q.k > 1.0f ? 1.0f : q.k;                 // 846-850
    r.sum = fmaf(eps * q.s, q.ePrev, r.sum);                   // 854: NOT multiplied by fs
    const float intensity = q.ePrev * r.fs;                    // 856
    if (!allowReflect || !(intensity > threshold)) return true;      // 884, 898, 683
    r.fs *= (1.0f - q.k);
    if (skipZero && r.fs == 0.0f) return true;
    // RMCRTCommon::reflect (581-602): back into the cell the ray came from (the delta of that step is the current one toggled back),
    // stepping direction of that axis flipped; tMax is NOT recomputed (reference behaviour)
    int dprev;
    if (w.axis == 0)      { dprev = r.b.dlx ^ 6;   r.b.dlx = dprev > 0 ? dprev - 8 : dprev + 8; }
    else if (w.axis == 1) { dprev = r.b.dly ^ tgy; r.b.dly = dprev > 0 ? dprev - sy8 : dprev + sy8; }
    else                  { dprev = r.b.dlz ^ tgz; r.b.dlz = dprev > 0 ? dprev - sz8 : dprev + sz8; }
    r.d.idx -= dprev;
    blk_decode(q, __ldg(blk + r.d.idx));
    q.A = fmaf(-q.k, w.t, q.tau);
    if (leg >= maxLegs) return false;
  }
}

// the whole ray.  Result in the scaled units of the half-packed records.
__device__ __forceinline__ float march_single_level_blk(const uint32_t* __restrict__ blk, int w0, unsigned par, int sy8, int sz8,
                                                        float dx, float dy, float dz, float rx, float ry, float rz,
                                                        float cx, float cy, float cz, float threshold, bool allowReflect,
                                                        bool skipZero, unsigned& nSteps) {
  if (!(1.0f > threshold)) return 0.0f;
  BlkRay r;
  float ddx, ddy, ddz;
  blk_ray_start(r, blk, w0, par, sy8, sz8, dx, dy, dz, rx, ry, rz, cx, cy, cz, ddx, ddy, ddz);
  blk_ray_legs(r, blk, ddx, ddy, ddz, sy8, sz8, threshold, allowReflect, skipZero, 0x7fffffff);
  nSteps += (unsigned)r.q.nsteps;
  return r.sum;
}

//------------------------------------------------------------------------------------------------------
// Data Onion on blocked copies, any number of levels: Ray::updateSumI_ML (Ray.cc:2349-2627) with march_leg_blk -- march_onion2_fast
// (march.cuh) generalised.  Every level above 0 is a blocked half-packed box: the fine level's region of interest (Ray.cc:2489), a coarser
// level's region (2490), each wrapped in marker cells {wall flag, abskg = inf} wherever the cell next to the box is an interior cell of the
// level (where the box ends at the domain boundary the level's own extra cells are copied, so a ray meets its wall on that level).  Entering
// a marker ends the leg like a wall and identifies the level jump of Ray.cc:2492-2527: the cell is decoded from the word index, mapped to
// the next coarser level (mapCellToCoarser), tMax of the stepping axis re-anchored as the reference does and the march goes on
// in that level's box; level 0 is the blocked copy of the whole level.  (The reference tests the stepped axis only; a box test is the same
// as long as the ray is inside the box before the step, which the host guarantees: every region contains the mapped region above it.)
//------------------------------------------------------------------------------------------------------
constexpr int ONION_MAXL = 4;
struct OnionBlk {
  const uint32_t* blk[ONION_MAXL];   // per level: blocked box incl. marker shell and guard blocks (level 0: the whole level)
  int bx[ONION_MAXL], bxy[ONION_MAXL];   // blocks per row / plane
  int lo[ONION_MAXL][3];             // cell index (on that level) of block-local coordinate 0
};

template <int MAXL>
__device__ __forceinline__ float march_onion_blk(const OnionGeom<MAXL>& G, const OnionBlk& ob, int ci, int cj, int ck, int w0,
                                                 float dx, float dy, float dz, float rx, float ry, float rz, float threshold,
                                                 bool allowReflect, bool skipZero, unsigned& nSteps) {
  typedef Num<float> N;
  if (!(1.0f > threshold)) return 0.0f;
  const int F = G.maxLevels - 1;
  int L = F;
  const float LN2 = 0.6931471805599453f;
  const bool posx = !sign_bit(dx), posy = !sign_bit(dy), posz = !sign_bit(dz);
  int sy8 = 8 * ob.bx[L], sz8 = 8 * ob.bxy[L];
  int tgy = 2 ^ (sy8 - 2), tgz = 4 ^ (sz8 - 4);
  BlkStride b = blk_stride_init((unsigned)(w0 & 7), posx, posy, posz, sy8, sz8);
  // sign[] of the reference: 1/0, after a reflection along that axis -1/-0 (RMCRTCommon.cc:599)
  float sgx = posx ? 1.0f : 0.0f, sgy = posy ? 1.0f : 0.0f, sgz = posz ? 1.0f : 0.0f;
  const float ix = N::rcp(dx), iy = N::rcp(dy), iz = N::rcp(dz);     // never refreshed on reflection (Ray.cc:2395)
  const float fx = (float)G.lev[F].dx[0], fy = (float)G.lev[F].dx[1], fz = (float)G.lev[F].dx[2];
  Dda d;
  d.tx = (sgx * fx - rx) * ix * N::LEN_SCALE;
  d.ty = (sgy * fy - ry) * iy * N::LEN_SCALE;
  d.tz = (sgz * fz - rz) * iz * N::LEN_SCALE;
  float ddx = fabsf(ix) * fx * N::LEN_SCALE, ddy = fabsf(iy) * fy * N::LEN_SCALE, ddz = fabsf(iz) * fz * N::LEN_SCALE;
  d.idx = w0;
  const uint32_t* cur = ob.blk[L];
  Integ q;
  blk_decode(q, __ldg(cur + w0));
  q.A = 0.0f; q.ePrev = 1.0f; q.nsteps = 0.0f; q.tau = 0.0f; q.leg = 0.0f;
  // position along each axis relative to the origin cell's low corner: l(t) = lb + (t - tb) ln2 * dir   (t is scaled by log2 e)
  float lbx = rx, lby = ry, lbz = rz, tbx = 0.0f, tby = 0.0f, tbz = 0.0f, dvx = dx, dvy = dy, dvz = dz;
  float sum = 0.0f, fs = 1.0f;
  for (;;) {
    q.leg = 0.0f;
    const LegOut w = march_leg_blk(d, q, b, cur, ddx, ddy, ddz, tgy, tgz);
    sum = fmaf(q.leg, fs, sum);
    const int ax = w.axis;
    if (L > 0 && q.k > 3.0e38f) {                    // stepped into the marker shell: leave this level's box (Ray.cc:2489-2527)
      const LevelGeom& gp = G.lev[L];                // the level being left
      const LevelGeom& gc = G.lev[L - 1];
      int B = d.idx >> 3;
      const int par = d.idx & 7;
      const int bk = B / ob.bxy[L]; B -= bk * ob.bxy[L];
      const int bj = B / ob.bx[L];
      const int bi = B - bj * ob.bx[L];
      const int c0 = ob.lo[L][0] + 2 * bi + (par & 1), c1 = ob.lo[L][1] + 2 * bj + ((par >> 1) & 1), c2 = ob.lo[L][2] + 2 * bk + (par >> 2);
      // re-anchor tMax of the stepping axis from the ray position and the centre of the cell entered, on the level being left
      // (Ray.cc:2486,2534-2550), relative to the origin cell's low corner
      const float pA = sel3(ax, (float)gp.dx[0], (float)gp.dx[1], (float)gp.dx[2]);
      const float ccRel = (sel3(ax, (float)gp.off[0], (float)gp.off[1], (float)gp.off[2]) + pA * ((float)sel3(ax, c0, c1, c2) + 0.5f)) -
                          sel3(ax, fx, fy, fz) * (float)sel3(ax, ci, cj, ck);
      const float lA = fmaf((w.t - sel3(ax, tbx, tby, tbz)) * LN2, sel3(ax, dvx, dvy, dvz), sel3(ax, lbx, lby, lbz));
      const float cdx = sel3(ax, (float)gc.dx[0], (float)gc.dx[1], (float)gc.dx[2]);
      const float rayDxLevel = lA - (ccRel - 0.5f * cdx);
      const float tTmp = (sel3(ax, sgx, sgy, sgz) * cdx - rayDxLevel) * sel3(ax, ix, iy, iz) * N::LEN_SCALE;
      if (ax == 0) d.tx = w.t + tTmp; else if (ax == 1) d.ty = w.t + tTmp; else d.tz = w.t + tTmp;
      // the coarser cell, blocked strides and tDelta of its level (the stepping directions are those of the current deltas)
      --L;
      const int k0 = min(max(map_to_coarser(c0, gp.rr[0]) - gc.lo[0], 0), gc.n[0] - 1) + gc.lo[0] - ob.lo[L][0];
      const int k1 = min(max(map_to_coarser(c1, gp.rr[1]) - gc.lo[1], 0), gc.n[1] - 1) + gc.lo[1] - ob.lo[L][1];
      const int k2 = min(max(map_to_coarser(c2, gp.rr[2]) - gc.lo[2], 0), gc.n[2] - 1) + gc.lo[2] - ob.lo[L][2];
      sy8 = 8 * ob.bx[L]; sz8 = 8 * ob.bxy[L];
      tgy = 2 ^ (sy8 - 2); tgz = 4 ^ (sz8 - 4);
      const unsigned parC = (unsigned)((k0 & 1) | ((k1 & 1) << 1) | ((k2 & 1) << 2));
      b = blk_stride_init(parC, b.dlx > 0, b.dly > 0, b.dlz > 0, sy8, sz8);
      ddx = fabsf(ix) * (float)gc.dx[0] * N::LEN_SCALE; ddy = fabsf(iy) * (float)gc.dx[1] * N::LEN_SCALE;
      ddz = fabsf(iz) * (float)gc.dx[2] * N::LEN_SCALE;
      d.idx = 8 * ((k0 >> 1) + ob.bx[L] * (k1 >> 1) + ob.bxy[L] * (k2 >> 1)) + (int)parC;
      cur = ob.blk[L];
      const uint32_t word = __ldg(cur + d.idx);
      blk_decode(q, word);
      q.A = fmaf(-q.k, w.t, q.tau);
      if ((int)word >= 0) continue;                  // keep marching on the coarser level
    }
    // wall (on whatever level): (q.k, q.s) = record of the non-flow cell entered
    const float eps = q.k > 1.0f ? 1.0f : q.k;                 // Ray.cc:2594-2598
    sum = fmaf(eps * q.s, q.ePrev, sum);                       // 2602: NOT multiplied by fs
    const float intensity = q.ePrev * fs;
    if (!allowReflect || !(intensity > threshold)) break;      // 2607-2609, 2621
    fs *= (1.0f - q.k);
    if (skipZero && fs == 0.0f) break;
    // reflect (RMCRTCommon.cc:581-602): previous cell on the CURRENT level, flip the stepping axis
    const float lHit = fmaf((w.t - sel3(ax, tbx, tby, tbz)) * LN2, sel3(ax, dvx, dvy, dvz), sel3(ax, lbx, lby, lbz));
    int dprev;
    if (ax == 0)      { dprev = b.dlx ^ 6;   b.dlx = dprev > 0 ? dprev - 8 : dprev + 8;     sgx = -sgx; dvx = -dvx; lbx = lHit; tbx = w.t; }
    else if (ax == 1) { dprev = b.dly ^ tgy; b.dly = dprev > 0 ? dprev - sy8 : dprev + sy8; sgy = -sgy; dvy = -dvy; lby = lHit; tby = w.t; }
    else              { dprev = b.dlz ^ tgz; b.dlz = dprev > 0 ? dprev - sz8 : dprev + sz8; sgz = -sgz; dvz = -dvz; lbz = lHit; tbz = w.t; }
    d.idx -= dprev;
    blk_decode(q, __ldg(cur + d.idx));
    q.A = fmaf(-q.k, w.t, q.tau);
  }
  nSteps += (unsigned)q.nsteps;
  return sum;
}

}  // namespace rmcrt
