// march.cuh -- Amanatides-Woo DDA ray march with the emission/absorption integral of RMCRT.
//
// Behavioural contract (what is reproduced, quirks included -- SURVEY.md 3.5):
//   single level : RMCRTCommon::updateSumI<T>   RMCRT/RMCRTCommon.cc:607-904
//   multi level  : Ray::updateSumI_ML<T>        RMCRT/Ray.cc:2349-2627
//   reflection   : RMCRTCommon::reflect         RMCRT/RMCRTCommon.cc:581-602
//
// B200 design: a cell is ONE packed record {abskg with the sign bit = "not a flow cell", sigmaT4/pi}
// (8 B for the FP32 production march, 16 B for the FP64 parity march) so every DDA step issues a single
// vectorised read-only load (LDG.E.64/.128.CONSTANT) that yields both the stop flag of the cell just entered
// and the properties integrated on the next step.  The FP32 march keeps all lengths pre-multiplied by
// log2(e) so the attenuation is one ex2.approx per step.
//
// Leg loops (PTX on the device, the same algorithm in C++ for tests/hostsim):
//   march_leg       one ray                                  ~22 SASS instructions per ray-cell step
//   march_leg_pair  two translates of one ray (one DDA)      ~15
//   march_leg_quad  four translates of one ray (one DDA)     ~10.5   <- production single-level and Data-Onion ROI march
// Translates = rays of neighbouring cells with the same direction and in-cell origin: identical tMax sequence, records a constant
// stride apart.  A joint leg ends at the first step where any of its rays enters a non-flow cell; the rays then finish one by one with
// march_leg (march_single_level_quad / _dual, march_onion2_quad).
#pragma once
#include <cstdint>
#include <cmath>
#include <cfloat>
#include "rng.cuh"
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

namespace rmcrt {

struct Rec8  { float  k, s; };   // k: abskg, sign bit set <=> cellType != flowCell ; s: sigmaT4/pi
struct Rec16 { double k, s; };
// half-packed record (4 B): high half = fp16 abskg with the wall flag in ITS sign bit (bit 31 of the word), low half = fp16 of
// sigmaT4/pi * HALF_S_SCALE.  The scale is a fixed power of two (exact), chosen so that 18 K < T < 3270 K stays in fp16's
// normal range; the march integrates in scaled units and the per-ray result is multiplied by 1/HALF_S_SCALE once.
struct Rec4  { uint32_t w; };
constexpr float HALF_S_SCALE = 1.0f / 32.0f;
template <class RecT> struct RecInfo;
template <> struct RecInfo<Rec8>  { typedef Rec8  decoded; typedef float  real; static constexpr float  unscale = 1.0f; };
template <> struct RecInfo<Rec16> { typedef Rec16 decoded; typedef double real; static constexpr double unscale = 1.0;  };
template <> struct RecInfo<Rec4>  { typedef Rec8  decoded; typedef float  real; static constexpr float  unscale = 1.0f / HALF_S_SCALE; };

__device__ __forceinline__ Rec8 ldrec(const Rec8* p) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return Rec8{v.x, v.y};
}
__device__ __forceinline__ Rec16 ldrec(const Rec16* p) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return Rec16{v.x, v.y};
}
#if defined(__CUDACC__)
__device__ __forceinline__ Rec8 ldrec(const Rec4* p) {       // decoded: k signed, s still scaled by HALF_S_SCALE
  const uint32_t w = __ldg(reinterpret_cast<const unsigned int*>(p));
  return Rec8{__half2float(__ushort_as_half((unsigned short)(w >> 16))), __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu)))};
}
__device__ __forceinline__ Rec4 encode_rec4(float kSigned, float s) {
  const uint32_t hk = __half_as_ushort(__float2half_rn(kSigned)), hs = __half_as_ushort(__float2half_rn(s * HALF_S_SCALE));
  return Rec4{(hk << 16) | hs};
}
#endif
// fp16 bits -> float (the half-packed words; on the host, tests/hostsim/cuda_shim.h supplies the conversion)
#if defined(__CUDACC__)
__device__ __forceinline__ float half_bits_to_float(unsigned short h) { return __half2float(__ushort_as_half(h)); }
#endif
__device__ __forceinline__ bool sign_bit(float v)  { return (__float_as_uint(v) >> 31) != 0u; }
__device__ __forceinline__ bool sign_bit(double v) { return (__double2hiint(v) < 0); }
// a record whose abskg carries the sign bit is "not a flow cell" (cellType != d_flowCell)
__device__ __forceinline__ bool is_wall(float k)  { return sign_bit(k); }
__device__ __forceinline__ bool is_wall(double k) { return sign_bit(k); }

template <class Real> struct Num;
template <> struct Num<float> {
  // all lengths carry a factor log2(e): exp(-tau) == ex2(-tau2)
  static constexpr float LEN_SCALE = 1.4426950408889634f;
  static __device__ __forceinline__ float attenuation(float tau2) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-tau2));
    return r;
#else
    return exp2f(-tau2);   // host build of this header (tests/hostsim) only
#endif
  }
  static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
  static __device__ __forceinline__ float add(float a, float b) { return a + b; }
  static __device__ __forceinline__ float rcp(float a) { return __frcp_rn(a); }
  static __device__ __forceinline__ float fuzz(float ds) { return ds; }
  static __device__ __forceinline__ float wall_product(float eps, float s, bool) { return eps * s; }
};
template <> struct Num<double> {
  static constexpr double LEN_SCALE = 1.0;
  static __device__ __forceinline__ double attenuation(double tau) { return exp(-tau); }
  // no FMA contraction: keeps the operation sequence of the reference's x86-64 build
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double rcp(double a) { return __ddiv_rn(1.0, a); }
  // RMCRTCommon.cc:717-720 (FUZZ = 1e-12); single level only
  static __device__ __forceinline__ double fuzz(double ds) { return (ds > -1e-12 && ds < 1e-12) ? ds + 1e-12 : ds; }
  // `wallEmissivity * sigmaT4OverPi[cur]` is a T*T product in the reference (RMCRTCommon.cc:846-854; NOT in Ray::updateSumI_ML, whose wallEmissivity is a double): evaluated in
  // float when the properties are float, before the double `* intensity`
  static __device__ __forceinline__ double wall_product(double eps, double s, bool floatProps) {
    return floatProps ? (double)((float)eps * (float)s) : __dmul_rn(eps, s);
  }
};

template <class Real> __device__ __forceinline__ Real absr(Real v);
template <> __device__ __forceinline__ float  absr(float v)  { return fabsf(v); }
template <> __device__ __forceinline__ double absr(double v) { return fabs(v); }

struct PathSink {       // debug only
  int32_t* buf;         // [cap][4]
  int cap;
  int n;
};

//------------------------------------------------------------------------------------------------------
// single level
//------------------------------------------------------------------------------------------------------
// A box of level records a block has staged in shared memory (origin tile + halo): north_star's "shared-memory staging of the per-patch
// origin block", built for the A/B measurement of DESIGN.md section 3 (rmcrt_params.originBrickHalo; default off).
struct Brick {
  const Rec8* sm;        // [n[2]][n[1]][n[0]], x fastest
  int lo[3], n[3];       // the cell window it covers
};

// recs  : level records, x-fastest, strides (1, sy, sz)
// idx   : record index of the origin cell
// d*    : ray direction;  r* : rayDx = distance of the ray origin from the cell's low corner (RMCRTCommon.cc:643-646)
// returns the ray's contribution to sumI
template <class Real, class RecT, bool RECORD, bool BRICK = false>
__device__ __forceinline__ Real march_single_level(const RecT* __restrict__ recs, long long idx, int sy, int sz,
                                                   Real dx, Real dy, Real dz, Real rx, Real ry, Real rz,
                                                   Real cx, Real cy, Real cz,   // cell size
                                                   Real threshold, bool allowReflect, bool skipZero,
                                                   unsigned& nSteps, PathSink* path, int pi, int pj, int pk,
                                                   Real maxLen = Real(INFINITY),     // d_maxRayLength * LEN_SCALE
                                                   bool floatProps = false,          // FP64 march of a float deck
                                                   const Brick* brick = nullptr) {   // BRICK: records inside this box come from shared memory
  typedef Num<Real> N;
  // the record of cell (i, j, k) = recs[index]: from the staged brick when the cell lies inside it
  auto fetch = [&](long long index, int i, int j, int k) -> typename RecInfo<RecT>::decoded {
    if constexpr (BRICK) {
      const unsigned x = (unsigned)(i - brick->lo[0]), y = (unsigned)(j - brick->lo[1]), z = (unsigned)(k - brick->lo[2]);
      if (x < (unsigned)brick->n[0] && y < (unsigned)brick->n[1] && z < (unsigned)brick->n[2])
        return brick->sm[x + (unsigned)brick->n[0] * (y + (unsigned)brick->n[1] * z)];
    }
    return ldrec(recs + index);
  };
  if (!(Real(1) > threshold) || !(Real(0) < maxLen)) return Real(0);   // while (intensity > d_threshold && rayLength < d_maxRayLength), RMCRTCommon.cc:680

  // raySignStep (RMCRTCommon.cc:482-494): +1 / -1 by the sign BIT of the direction
  const bool posx = !sign_bit(dx), posy = !sign_bit(dy), posz = !sign_bit(dz);
  int stx = posx ? 1 : -1, sty = posy ? sy : -sy, stz = posz ? sz : -sz;
  const Real ix = N::rcp(dx), iy = N::rcp(dy), iz = N::rcp(dz);
  // tMax (649-652) and tDelta (655)
  Real tx = N::mul(N::mul(N::add(posx ? cx : Real(0), -rx), ix), N::LEN_SCALE);
  Real ty = N::mul(N::mul(N::add(posy ? cy : Real(0), -ry), iy), N::LEN_SCALE);
  Real tz = N::mul(N::mul(N::add(posz ? cz : Real(0), -rz), iz), N::LEN_SCALE);
  const Real ddx = N::mul(N::mul(absr(ix), cx), N::LEN_SCALE);
  const Real ddy = N::mul(N::mul(absr(iy), cy), N::LEN_SCALE);
  const Real ddz = N::mul(N::mul(absr(iz), cz), N::LEN_SCALE);

  Real tPrev = 0, tau = 0, ePrev = 1, sum = 0, fs = 1;
  Real len = 0;                    // rayLength (723), only compared when the ROI is boundedRayLength
  typename RecInfo<RecT>::decoded r = fetch(idx, pi, pj, pk);
  Real k = absr(r.k), s = r.s;
  unsigned steps = 0;
  int ci = pi, cj = pj, ck = pk;   // only maintained when RECORD

  for (;;) {
    long long idxPrev;
    bool px, py;
    Real e;
    do {
      idxPrev = idx;
      // axis with the smallest tMax; strict '<' so ties go Y-before-... exactly as RMCRTCommon.cc:696-708
      const bool xy = tx < ty;
      px = xy && (tx < tz);
      py = !xy && (ty < tz);
      const Real tNext = px ? tx : (py ? ty : tz);
      idx += px ? stx : (py ? sty : stz);
      if (RECORD || BRICK) {
        if (px) ci += (stx > 0 ? 1 : -1); else if (py) cj += (sty > 0 ? 1 : -1); else ck += (stz > 0 ? 1 : -1);
      }
      r = fetch(idx, ci, cj, ck);            // the one load of this step
      Real ds = N::add(tNext, -tPrev);
      tPrev = tNext;
      tx = px ? N::add(tx, ddx) : tx;
      ty = py ? N::add(ty, ddy) : ty;
      tz = (px || py) ? tz : N::add(tz, ddz);
      ds = N::fuzz(ds);
      len = N::add(len, ds);                 // rayLength += disMin   (723)
      tau = N::add(tau, N::mul(k, ds));      // optical_thickness += abskg[prev]*disMin   (731)
      ++steps;
      e = N::attenuation(tau);               // 782
      sum = N::add(sum, N::mul(N::mul(s, N::add(ePrev, -e)), fs));   // 786
      ePrev = e;
      k = absr(r.k);
      s = r.s;
      if (RECORD && path && path->n < path->cap) {
        int32_t* q = path->buf + 4 * path->n; q[0] = ci; q[1] = cj; q[2] = ck; q[3] = 0;
      }
      if (RECORD && path) path->n++;
    } while (!is_wall(r.k) && len < maxLen); // in_domain = (celltype[cur] == d_flowCell) (729) && rayLength < d_maxRayLength (684)

    // wall: r = record of the non-flow cell entered -- or, when the ray ran out of length, of the FLOW cell it stopped in:
    // the reference then adds that cell's abskg*sigmaT4/pi like a wall emission (846-854) and ends the ray (680)
    const Real eps = k > Real(1) ? Real(1) : k;          // 846-850
    Real intensity = e;                                   // exp(-optical_thickness), same argument as the last step
    sum = N::add(sum, N::mul(N::wall_product(eps, s, floatProps), intensity)); // 854: NOT multiplied by fs
    intensity = N::mul(intensity, fs);                    // 856
    if (!allowReflect || !(intensity > threshold) || !(len < maxLen)) break; // 884, 898, 680
    // RMCRTCommon::reflect (581-602)
    fs = N::mul(fs, N::add(Real(1), -k));
    idx = idxPrev;
    if (px) stx = -stx; else if (py) sty = -sty; else stz = -stz;
    if (RECORD || BRICK) { if (px) ci += (stx > 0 ? 1 : -1); else if (py) cj += (sty > 0 ? 1 : -1); else ck += (stz > 0 ? 1 : -1); }
    if (skipZero && fs == Real(0)) break;                 // exact: every later term carries fs==0 or a zero wall emission
    r = fetch(idx, ci, cj, ck);
    k = absr(r.k);
    s = r.s;
  }
  nSteps += steps;
  return sum;
}

//------------------------------------------------------------------------------------------------------
// single level with isotropic scattering: RMCRTCommon::updateSumI<T> as compiled with RAY_SCATTER (--enable-ray-scatter),
// RMCRTCommon.cc:670-676 and 790-836 on top of the march above.  Absorption still uses abskg only; a scattering event re-draws the
// direction, puts the ray back into the previous cell if the stepping axis flipped, and rebuilds tMax / tDelta from the ray's
// ABSOLUTE location in the (new) current cell -- which is why this variant tracks cell indices and the location like the reference
// does.  Lengths are physical here (tMax carries LEN_SCALE only where it multiplies abskg).  Generic template, not the PTX fast path.
//   org : absolute ray origin;  anchor/cdx : Level::getCellPosition data;  sigmaScat > 0
template <class Real, class RecT>
__device__ __forceinline__ Real march_single_level_scatter(const RecT* __restrict__ recs, long long idx, int sy, int sz, int ci, int cj, int ck,
                                                           const double dir0[3], const double org[3], const double anchor[3],
                                                           const double cdx[3], Real threshold, bool allowReflect, unsigned& nSteps,
                                                           Real maxLen, bool floatProps, double sigmaScat, ScatRng& rng) {
  typedef Num<Real> N;
  typedef Num<double> D;
  if (!(Real(1) > threshold) || !(Real(0) < maxLen)) return Real(0);
  const Real LS = N::LEN_SCALE;
  double dir[3] = {dir0[0], dir0[1], dir0[2]};
  double loc[3] = {org[0], org[1], org[2]};                       // ray_location (667)
  int c[3] = {ci, cj, ck};
  int st[3];
  double sgn[3];
  Real t[3], dd[3];
  // tMax / tDelta from the location inside cell c (643-655 at the start, 812-822 after an event); geometry in FP64 like the reference
  auto setup = [&]() {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const bool pos = !sign_bit(dir[d]);
      st[d] = pos ? 1 : -1;
      sgn[d] = pos ? 1.0 : 0.0;
      const double inv = D::rcp(dir[d]);
      const double cc = D::add(D::add(anchor[d], D::mul(cdx[d], (double)c[d])), D::mul(cdx[d], 0.5));   // Level.cc:473-477
      const double rayDx = D::add(loc[d], -D::add(cc, -D::mul(0.5, cdx[d])));
      t[d] = (Real)D::mul(D::add(D::mul(sgn[d], cdx[d]), -rayDx), inv) * LS;
      dd[d] = (Real)D::mul(fabs(inv), cdx[d]) * LS;
    }
  };
  setup();
  const double scatCoeff = sigmaScat > 1e-99 ? sigmaScat : 1e-99;
  double scatLength = -log(rng.open01()) / scatCoeff;            // 675
  double lenScatter = 0.0;
  Real tPrev = 0, tau = 0, ePrev = 1, sum = 0, fs = 1, len = 0;
  typename RecInfo<RecT>::decoded r = ldrec(recs + idx);
  Real k = absr(r.k), s = r.s;
  unsigned steps = 0;
  for (;;) {
    long long idxPrev;
    int ax, cPrev;
    Real e;
    do {
      idxPrev = idx;
      ax = (t[0] < t[1]) ? ((t[0] < t[2]) ? 0 : 2) : ((t[1] < t[2]) ? 1 : 2);        // 696-708
      const Real tNext = ax == 0 ? t[0] : (ax == 1 ? t[1] : t[2]);
      cPrev = ax == 0 ? c[0] : (ax == 1 ? c[1] : c[2]);
      if (ax == 0) { c[0] += st[0]; idx += st[0]; } else if (ax == 1) { c[1] += st[1]; idx += st[1] * (long long)sy; } else { c[2] += st[2]; idx += st[2] * (long long)sz; }
      r = ldrec(recs + idx);
      Real ds = N::add(tNext, -tPrev);
      tPrev = tNext;
      if (ax == 0) t[0] = N::add(t[0], dd[0]); else if (ax == 1) t[1] = N::add(t[1], dd[1]); else t[2] = N::add(t[2], dd[2]);
      ds = N::fuzz(ds);
      len = N::add(len, ds);
      const double dsPhys = (double)ds / (double)LS;
      lenScatter = D::add(lenScatter, dsPhys);                     // 724
#pragma unroll
      for (int d = 0; d < 3; ++d) loc[d] = D::add(loc[d], D::mul(dsPhys, dir[d]));    // 726-728
      tau = N::add(tau, N::mul(k, ds));
      ++steps;
      e = N::attenuation(tau);
      sum = N::add(sum, N::mul(N::mul(s, N::add(ePrev, -e)), fs));
      ePrev = e;
      k = absr(r.k);
      s = r.s;
      if (lenScatter > scatLength && !is_wall(r.k)) {              // 790: a scattering event inside the flow cell just entered
        scatLength = -log(rng.open01()) / scatCoeff;
        rng.reseedAt(c[0], c[1], c[2]);
        const double pm1 = 2.0 * rng.open01() - 1.0 + DBL_EPSILON;  // findRayDirection (517-530)
        const double rr = sqrt(1.0 - pm1 * pm1);
        const double theta = 2.0 * M_PI * rng.open01();
        double sn, cs;
        sincos(theta, &sn, &cs);
        const int stepOld = ax == 0 ? st[0] : (ax == 1 ? st[1] : st[2]);
        dir[0] = rr * cs; dir[1] = rr * sn; dir[2] = pm1;
        const bool posA = !sign_bit(ax == 0 ? dir[0] : (ax == 1 ? dir[1] : dir[2]));
        if ((posA ? 1 : -1) * stepOld < 0) {                       // 803-807: back scattering, the ray returns to prevCell
          if (ax == 0) c[0] = cPrev; else if (ax == 1) c[1] = cPrev; else c[2] = cPrev;
          idx = idxPrev;
          r = ldrec(recs + idx);
          k = absr(r.k);
          s = r.s;
        }
        setup();
        tPrev = 0;                                                 // 833
        lenScatter = 0.0;                                          // 834
      }
    } while (!is_wall(r.k) && len < maxLen);
    const Real eps = k > Real(1) ? Real(1) : k;
    Real intensity = e;
    sum = N::add(sum, N::mul(N::wall_product(eps, s, floatProps), intensity));
    intensity = N::mul(intensity, fs);
    if (!allowReflect || !(intensity > threshold) || !(len < maxLen)) break;
    fs = N::mul(fs, N::add(Real(1), -k));                          // reflect (581-602)
    idx = idxPrev;
    if (ax == 0) { c[0] = cPrev; st[0] = -st[0]; sgn[0] = -sgn[0]; dir[0] = -dir[0]; }
    else if (ax == 1) { c[1] = cPrev; st[1] = -st[1]; sgn[1] = -sgn[1]; dir[1] = -dir[1]; }
    else { c[2] = cPrev; st[2] = -st[2]; sgn[2] = -sgn[2]; dir[2] = -dir[2]; }
    r = ldrec(recs + idx);
    k = absr(r.k);
    s = r.s;
  }
  nSteps += steps;
  return sum;
}

//------------------------------------------------------------------------------------------------------
// single level, production FP32 fast path.  Same ray, same cell sequence and the same integral as
// march_single_level<float, Rec8, false>, arranged for the B200 issue model (ALU and FMA pipes each take one
// warp instruction every 2 cycles per SM sub-partition, so the step is written to balance them):
//   * 32-bit record index (one IMAD.WIDE per address instead of a 64-bit add + two LEAs);
//   * tau in "offset form": inside a cell tau(t) = A + k t, so a step is one FFMA for tau and one FFMA that
//     re-bases A with the next cell's k -- no tPrev, no separate ds;
//   * the per-leg integral is accumulated without fs and scaled once per wall hit (fs is constant on a leg);
//   * the step is unrolled by two so the loop-carried (ePrev, k, s) need no register moves;
//   * the record of the NEXT step is requested before the wall test of the current one (see below);
//   * executed steps are counted in an FP32 register (exact below 2^24 per ray; FADD runs on the FMA pipe).
// The axis choice keeps the reference's strict-'<' tie order (RMCRTCommon.cc:696-708).
//------------------------------------------------------------------------------------------------------
// DDA state of one ray.  A step is split into PEEK (which axis steps next, at which t, into which record -- no state
// change) and COMMIT (advance tMax of that axis and the index), so the record of step n+1 can be requested before the wall
// test of step n resolves while the state still is exactly "after step n" if that test fires.
struct Dda {
  float tx, ty, tz;      // tMax, scaled by log2(e)
  int idx;               // record index of the current cell
};
struct Peek {
  float t;               // == px ? tx : py ? ty : tz
  bool px, py, pz;       // stepping axis; strict '<' tie order of RMCRTCommon.cc:696-708
  int idx;               // record index of the cell the step enters
};
__device__ __forceinline__ Peek dda_peek(const Dda& d, int stx, int sty, int stz) {
  Peek p;
  const bool a = d.tx < d.ty;
  const float m = fminf(d.tx, d.ty);
  const bool nz_ = m < d.tz;                 // false: z is the stepping axis
  p.px = a && nz_;
  p.py = !a && nz_;
  p.pz = !nz_;
  p.t = fminf(m, d.tz);
  p.idx = d.idx + (p.px ? stx : (p.py ? sty : stz));
  return p;
}
__device__ __forceinline__ void dda_commit(Dda& d, const Peek& p, float ddx, float ddy, float ddz) {
  d.idx = p.idx;
  d.tx = p.px ? d.tx + ddx : d.tx;      // predicated in-place adds, no branches
  d.ty = p.py ? d.ty + ddy : d.ty;
  d.tz = p.pz ? d.tz + ddz : d.tz;
}

// emission/absorption integral state along one ray
struct Integ {
  float A;               // tau(t) = A + k t inside the current cell
  float k, s;            // properties of the current cell
  float ePrev;           // exp(-tau) on entry to the current cell
  float leg;             // sum over the leg of s (ePrev - e)
  float nsteps;
  float tau;             // optical thickness at the last cell face crossed
};
// cross the face at t into the cell whose record is r
__device__ __forceinline__ void integ_cross(Integ& q, float t, const Rec8 r) {
  q.tau = fmaf(q.k, t, q.A);                          // optical thickness at the cell exit (RMCRTCommon.cc:731)
  const float e = Num<float>::attenuation(q.tau);     // 782
  q.leg = fmaf(q.s, q.ePrev - e, q.leg);              // 786 (fs applied per leg)
  q.ePrev = e;
  q.nsteps += 1.0f;
  q.k = fabsf(r.k);
  q.s = r.s;
  q.A = fmaf(-q.k, t, q.tau);                         // re-base for the cell just entered
}

// One LEG of a ray: march from the current cell until a non-flow cell is entered.
//
// The loads are software-pipelined: the DDA does not depend on the loaded data, so the record of step n+1 is requested
// before the wall test of step n resolves (two records in flight per ray), while the DDA state is only committed after
// that test -- at a wall the state is exactly "after step n".  The speculative request past a wall reads a neighbour of
// a non-flow cell: inside the level, or inside the guard records rmcrt_define_level puts around it.
//
// On the device the loop is written in PTX: three steps per iteration so that the three cells in play (being left, being
// entered, in flight) rotate through fixed registers without moves, predicated in-place tMax updates, and no
// re-association by the compiler; ~21 SASS instructions per ray-cell step.  The C++ version below it is the same
// algorithm statement by statement and is what the host-side simulation (tests/hostsim) steps against the oracle.
struct LegOut {
  float t;        // tMax of the step that entered the wall
  int axis;       // stepping axis of that step: 0 x, 1 y, 2 z
};

#if defined(__CUDA_ARCH__)
// record-type specific pieces: request the record of slot J / decode slot C into a<C> = |abskg|, s<C> (k<C> keeps the raw word
// whose bit 31 is the wall flag for both layouts)
// Address arithmetic, measured alternatives (sm_100a SASS, session 8): `mad.wide.s32 ad, i, 8, base` becomes LEA + LEA.HI.X (2 ALU
// ops, what ships); a REGISTER record size (%20, unused) makes it IMAD.WIDE + IADD3 + IMAD.X (3 ops); 32-bit low-word addressing with a
// constant high word costs two MOVs per step to rebuild the register pair the load clobbers (no gain at the 32-register budget);
// a 32-bit offset from a uniform base has no single-instruction LDG form (IADD3 + IMAD.X).
// Step counter, measured alternative: dropping the per-step `add.f32 nsteps` and deriving the count per leg from the DDA state
// ((tMax_after - tMax_before) / tDelta per axis) removes one instruction per step but was 0.8 % SLOWER at 128^3 (three more values
// live across the leg at the 32-register budget: 236 B of spills instead of 180 B) and off by 3 steps in 5.7e8; not adopted.
#define RM_LD8(J)  "mad.wide.s32 ad, i" #J ", 8, %13;\n\t" "ld.global.nc.v2.f32 {k" #J ", s" #J "}, [ad];\n\t"
#define RM_LD4(J)  "mad.wide.s32 ad, i" #J ", 4, %13;\n\t" "ld.global.nc.b32 k" #J ", [ad];\n\t"
#define RM_DEC8(C) "abs.f32 a" #C ", k" #C ";\n\t"
#define RM_DEC4(C) "mov.b32 {hl, hh}, k" #C ";\n\t" "cvt.f32.f16 s" #C ", hl;\n\t" "cvt.f32.f16 a" #C ", hh;\n\t" "abs.f32 a" #C ", a" #C ";\n\t"
#define RM_PEEK(J, IB, LD)                                  \
  "min.f32 t" #J ", %0, %1, %2;\n\t"                        \
  "setp.neu.f32 n" #J ", t" #J ", %2;\n\t"                  \
  "setp.lt.and.f32 x" #J ", %0, %1, n" #J ";\n\t"           \
  "setp.ge.and.f32 y" #J ", %0, %1, n" #J ";\n\t"           \
  "selp.s32 i" #J ", %17, %19, x" #J ";\n\t"                \
  "@y" #J " mov.s32 i" #J ", %18;\n\t"                      \
  "add.s32 i" #J ", i" #J ", " IB ";\n\t"                   \
  LD(J)
#define RM_COMMIT(J)                                        \
  "@x" #J " add.f32 %0, %0, %14;\n\t"                       \
  "@y" #J " add.f32 %1, %1, %15;\n\t"                       \
  "@!n" #J " add.f32 %2, %2, %16;\n\t"
// leave the cell in slot P at t<C>, enter the cell in slot C
#define RM_INTEG(P, C, DEC)                                 \
  "fma.rn.f32 ta, a" #P ", t" #C ", %4;\n\t"                \
  "neg.f32 d, ta;\n\t"                                      \
  "ex2.approx.ftz.f32 e" #C ", d;\n\t"                      \
  "sub.f32 d, e" #P ", e" #C ";\n\t"                        \
  "fma.rn.f32 %8, s" #P ", d, %8;\n\t"                      \
  "add.f32 %9, %9, 0f3F800000;\n\t"                         \
  DEC(C)                                                    \
  "neg.f32 d, a" #C ";\n\t"                                 \
  "fma.rn.f32 %4, d, t" #C ", ta;\n\t"                      \
  "setp.lt.s32 w, k" #C ", 0;\n\t"                          \
  "@w bra EXIT" #C ";\n\t"
#define RM_EXIT(C)                                          \
  "EXIT" #C ":\n\t"                                         \
  "mov.s32 %3, i" #C ";\n\t"                                \
  "mov.f32 %10, t" #C ";\n\t"                               \
  "selp.s32 %11, 1, 2, y" #C ";\n\t"                        \
  "@x" #C " mov.s32 %11, 0;\n\t"                            \
  "mov.f32 %5, a" #C ";\n\t"                                \
  "mov.f32 %6, s" #C ";\n\t"                                \
  "mov.f32 %7, e" #C ";\n\t"                                \
  "bra DONE;\n\t"
#define RM_ASM(LD, DEC)                                                                                              \
  asm volatile(                                                                                                      \
      "{\n\t"                                                                                                        \
      ".reg .pred x0, y0, n0, x1, y1, n1, x2, y2, n2, w;\n\t"                                                        \
      ".reg .f32 d, ta, t0, t1, t2, s0, s1, s2, e0, e1, e2, a0, a1, a2;\n\t"                                      \
      ".reg .b32 k0, k1, k2;\n\t"                                                                                    \
      ".reg .b16 hl, hh;\n\t"                                                                                        \
      ".reg .s32 i0, i1, i2;\n\t"                                                                                    \
      ".reg .s64 ad;\n\t"                                                                                            \
      /* slot 2 = the cell the leg starts in */                                                                      \
      "mov.f32 a2, %5;\n\t"                                                                                          \
      "mov.f32 s2, %6;\n\t"                                                                                          \
      "mov.f32 e2, %7;\n\t"                                                                                          \
      RM_PEEK(0, "%3", LD)                                                                                           \
      "LOOP:\n\t"                                                                                                    \
      RM_COMMIT(0) RM_PEEK(1, "i0", LD) RM_INTEG(2, 0, DEC)                                                          \
      RM_COMMIT(1) RM_PEEK(2, "i1", LD) RM_INTEG(0, 1, DEC)                                                          \
      RM_COMMIT(2) RM_PEEK(0, "i2", LD) RM_INTEG(1, 2, DEC)                                                          \
      "bra LOOP;\n\t"                                                                                                \
      RM_EXIT(0) RM_EXIT(1) RM_EXIT(2)                                                                               \
      "DONE:\n\t"                                                                                                    \
      "mov.f32 %12, ta;\n\t"                                                                                         \
      "}"                                                                                                            \
      : "+f"(d.tx), "+f"(d.ty), "+f"(d.tz), "+r"(d.idx), "+f"(q.A), "+f"(q.k), "+f"(q.s), "+f"(q.ePrev), "+f"(q.leg),  \
        "+f"(q.nsteps), "=f"(o.t), "=r"(o.axis), "=f"(q.tau)                                                         \
      : "l"(recs), "f"(ddx), "f"(ddy), "f"(ddz), "r"(stx), "r"(sty), "r"(stz), "r"(recBytes)                         \
      : "memory")

__device__ __forceinline__ LegOut march_leg(Dda& d, Integ& q, const Rec8* __restrict__ recs, float ddx, float ddy, float ddz,
                                            int stx, int sty, int stz, int recBytes) {
  LegOut o;
  RM_ASM(RM_LD8, RM_DEC8);
  return o;
}
__device__ __forceinline__ LegOut march_leg(Dda& d, Integ& q, const Rec4* __restrict__ recs, float ddx, float ddy, float ddz,
                                            int stx, int sty, int stz, int recBytes) {
  LegOut o;
  RM_ASM(RM_LD4, RM_DEC4);
  return o;
}
#undef RM_ASM
#undef RM_PEEK
#undef RM_COMMIT
#undef RM_INTEG
#undef RM_EXIT
#undef RM_LD8
#undef RM_LD4
#undef RM_DEC8
#undef RM_DEC4
#else
__device__ __forceinline__ LegOut march_leg(Dda& d, Integ& q, const Rec8* __restrict__ recs, float ddx, float ddy, float ddz,
                                            int stx, int sty, int stz, int /*recBytes*/) {
  Peek p1 = dda_peek(d, stx, sty, stz), p2;
  Rec8 r1 = ldrec(recs + p1.idx), r2;
  for (;;) {
    dda_commit(d, p1, ddx, ddy, ddz);
    p2 = dda_peek(d, stx, sty, stz);          // speculative: issued before r1's wall test
    r2 = ldrec(recs + p2.idx);
    integ_cross(q, p1.t, r1);
    if (is_wall(r1.k)) return LegOut{p1.t, p1.px ? 0 : (p1.py ? 1 : 2)};
    dda_commit(d, p2, ddx, ddy, ddz);
    p1 = dda_peek(d, stx, sty, stz);
    r1 = ldrec(recs + p1.idx);
    integ_cross(q, p2.t, r2);
    if (is_wall(r2.k)) return LegOut{p2.t, p2.px ? 0 : (p2.py ? 1 : 2)};
  }
}
#endif

template <class RecT>
__device__ __forceinline__ float march_single_level_fast(const RecT* __restrict__ recs, int idx0, int sy, int sz,
                                                         float dx, float dy, float dz, float rx, float ry, float rz,
                                                         float cx, float cy, float cz, float threshold, bool allowReflect,
                                                         bool skipZero, unsigned& nSteps, int recBytes) {
  typedef Num<float> N;
  if (!(1.0f > threshold)) return 0.0f;
  const bool posx = !sign_bit(dx), posy = !sign_bit(dy), posz = !sign_bit(dz);
  int stx = posx ? 1 : -1, sty = posy ? sy : -sy, stz = posz ? sz : -sz;
  const float ix = N::rcp(dx), iy = N::rcp(dy), iz = N::rcp(dz);
  Dda d;
  d.tx = ((posx ? cx : 0.0f) - rx) * ix * N::LEN_SCALE;
  d.ty = ((posy ? cy : 0.0f) - ry) * iy * N::LEN_SCALE;
  d.tz = ((posz ? cz : 0.0f) - rz) * iz * N::LEN_SCALE;
  const float ddx = fabsf(ix) * cx * N::LEN_SCALE, ddy = fabsf(iy) * cy * N::LEN_SCALE, ddz = fabsf(iz) * cz * N::LEN_SCALE;
  d.idx = idx0;
  Integ q;
  {
    const Rec8 r = ldrec(recs + idx0);
    q.k = fabsf(r.k); q.s = r.s;
  }
  q.A = 0.0f; q.ePrev = 1.0f; q.nsteps = 0.0f; q.tau = 0.0f;
  float sum = 0.0f, fs = 1.0f;
  for (;;) {
    q.leg = 0.0f;
    const LegOut w = march_leg(d, q, recs, ddx, ddy, ddz, stx, sty, stz, recBytes);
    sum = fmaf(q.leg, fs, sum);
    // wall: (q.k, q.s) = record of the non-flow cell entered, q.ePrev = exp(-tau) at its face
    const float eps = q.k > 1.0f ? 1.0f : q.k;                 // 846-850
    sum = fmaf(eps * q.s, q.ePrev, sum);                       // 854: NOT multiplied by fs
    const float intensity = q.ePrev * fs;                      // 856
    if (!allowReflect || !(intensity > threshold)) break;      // 884, 898, 683
    // RMCRTCommon::reflect (581-602): back to the cell the ray came from, flip the stepping axis
    fs *= (1.0f - q.k);
    if (skipZero && fs == 0.0f) break;
    d.idx -= w.axis == 0 ? stx : (w.axis == 1 ? sty : stz);
    if (w.axis == 0) stx = -stx; else if (w.axis == 1) sty = -sty; else stz = -stz;
    const Rec8 r = ldrec(recs + d.idx);
    q.k = fabsf(r.k); q.s = r.s;
    q.A = fmaf(-q.k, w.t, q.tau);                                // tMax is NOT recomputed on reflection (reference behaviour)
  }
  nSteps += (unsigned)q.nsteps;
  return sum;
}

//------------------------------------------------------------------------------------------------------
// single level, TWO cells per lane (rayCoherence == 2, cellsPerLane == 2).  With a shared direction AND a shared in-cell origin the
// rays of neighbouring cells are exact translates: tMax, the stepping axis and tDelta are identical, and the record a ray enters is
// its neighbour's plus a constant stride.  One DDA therefore drives the rays of cell A and of cell B = A + (0,1,0) of the same lane:
// 12 DDA instructions are shared by two loads / integrals per step.  The joint leg ends at the first step where either ray enters a
// non-flow cell; from there each ray finishes on its own with the single-ray leg loop (they have different histories once one of them
// met a wall).  Same arithmetic per ray as march_single_level_fast, so the result of each ray is identical to the single-cell kernel's.
//------------------------------------------------------------------------------------------------------
struct RayState {
  Dda d;
  Integ q;
  float sum, fs;
  int stx, sty, stz;
};
// the ray just entered a non-flow cell (q.k, q.s = its record) at w: wall emission, threshold test, reflection (RMCRTCommon.cc:846-902).
// Returns true when the ray goes on (reflected into the cell it came from).
template <class RecT>
__device__ __forceinline__ bool ray_wall(RayState& r, const LegOut& w, const RecT* __restrict__ recs, float threshold, bool allowReflect,
                                         bool skipZero) {
  Integ& q = r.q;
  r.sum = fmaf(q.leg, r.fs, r.sum);
  q.leg = 0.0f;
  const float eps = q.k > 1.0f ? 1.0f : q.k;
  r.sum = fmaf(eps * q.s, q.ePrev, r.sum);
  const float intensity = q.ePrev * r.fs;
  if (!allowReflect || !(intensity > threshold)) return false;
  r.fs *= (1.0f - q.k);
  if (skipZero && r.fs == 0.0f) return false;
  r.d.idx -= w.axis == 0 ? r.stx : (w.axis == 1 ? r.sty : r.stz);
  if (w.axis == 0) r.stx = -r.stx; else if (w.axis == 1) r.sty = -r.sty; else r.stz = -r.stz;
  const Rec8 rec = ldrec(recs + r.d.idx);
  q.k = fabsf(rec.k); q.s = rec.s;
  q.A = fmaf(-q.k, w.t, q.tau);
  return true;
}
template <class RecT>
__device__ __forceinline__ void ray_finish(RayState& r, const RecT* __restrict__ recs, float ddx, float ddy, float ddz, float threshold,
                                           bool allowReflect, bool skipZero, int recBytes) {
  for (;;) {
    const LegOut w = march_leg(r.d, r.q, recs, ddx, ddy, ddz, r.stx, r.sty, r.stz, recBytes);
    if (!ray_wall(r, w, recs, threshold, allowReflect, skipZero)) break;
  }
}

// The joint leg of two translate rays, PTX on the device like march_leg: three-slot rotation, the records of step n+1 of BOTH rays
// requested before the wall tests of step n, one DDA (peek / commit) per step for the pair.  Ends at the first step where either ray
// enters a non-flow cell.  Returns axis | 4*(ray A hit a wall) | 8*(ray B hit a wall) in LegOut::axis.
#if defined(__CUDA_ARCH__)
// (%20 = byte distance of ray B's record from ray A's: the second address is one 64-bit add, not a second multiply-add)
#define RJ_LD8(J)                                                                                   \
  "mad.wide.s32 ad, i" #J ", 8, %19;\n\t" "ld.global.nc.v2.f32 {ka" #J ", sa" #J "}, [ad];\n\t"     \
  "add.s64 ad, ad, %20;\n\t" "ld.global.nc.v2.f32 {kb" #J ", sb" #J "}, [ad];\n\t"
#define RJ_LD4(J)                                                                                   \
  "mad.wide.s32 ad, i" #J ", 4, %19;\n\t" "ld.global.nc.b32 ka" #J ", [ad];\n\t"                    \
  "add.s64 ad, ad, %20;\n\t" "ld.global.nc.b32 kb" #J ", [ad];\n\t"
#define RJ_DEC8(C) "abs.f32 aa" #C ", ka" #C ";\n\t" "abs.f32 ab" #C ", kb" #C ";\n\t"
#define RJ_DEC4(C)                                                                                  \
  "mov.b32 {hl, hh}, ka" #C ";\n\t" "cvt.f32.f16 sa" #C ", hl;\n\t" "cvt.f32.f16 aa" #C ", hh;\n\t" "abs.f32 aa" #C ", aa" #C ";\n\t" \
  "mov.b32 {hl, hh}, kb" #C ";\n\t" "cvt.f32.f16 sb" #C ", hl;\n\t" "cvt.f32.f16 ab" #C ", hh;\n\t" "abs.f32 ab" #C ", ab" #C ";\n\t"
#define RJ_PEEK(J, IB, LD)                                  \
  "min.f32 t" #J ", %0, %1, %2;\n\t"                        \
  "setp.neu.f32 n" #J ", t" #J ", %2;\n\t"                  \
  "setp.lt.and.f32 x" #J ", %0, %1, n" #J ";\n\t"           \
  "setp.ge.and.f32 y" #J ", %0, %1, n" #J ";\n\t"           \
  "selp.s32 i" #J ", %24, %26, x" #J ";\n\t"                \
  "@y" #J " mov.s32 i" #J ", %25;\n\t"                      \
  "add.s32 i" #J ", i" #J ", " IB ";\n\t"                   \
  LD(J)
#define RJ_COMMIT(J)                                        \
  "@x" #J " add.f32 %0, %0, %21;\n\t"                       \
  "@y" #J " add.f32 %1, %1, %22;\n\t"                       \
  "@!n" #J " add.f32 %2, %2, %23;\n\t"
#define RJ_INTEG(P, C, DEC)                                 \
  "fma.rn.f32 taa, aa" #P ", t" #C ", %4;\n\t"              \
  "fma.rn.f32 tab, ab" #P ", t" #C ", %9;\n\t"              \
  "neg.f32 d, taa;\n\t"                                     \
  "ex2.approx.ftz.f32 ea" #C ", d;\n\t"                     \
  "neg.f32 d, tab;\n\t"                                     \
  "ex2.approx.ftz.f32 eb" #C ", d;\n\t"                     \
  "sub.f32 d, ea" #P ", ea" #C ";\n\t"                      \
  "fma.rn.f32 %8, sa" #P ", d, %8;\n\t"                     \
  "sub.f32 d, eb" #P ", eb" #C ";\n\t"                      \
  "fma.rn.f32 %13, sb" #P ", d, %13;\n\t"                   \
  DEC(C)                                                    \
  "neg.f32 d, aa" #C ";\n\t"                                \
  "fma.rn.f32 %4, d, t" #C ", taa;\n\t"                     \
  "neg.f32 d, ab" #C ";\n\t"                                \
  "fma.rn.f32 %9, d, t" #C ", tab;\n\t"                     \
  "lop3.or.b32 wz|w, ka" #C ", kb" #C ", 0x80000000, 0xA8, pf;\n\t" \
  "@w bra EXIT" #C ";\n\t"
// (steps are counted once per loop iteration of three; an exit adds the steps of the iteration it leaves: NS = C + 1 as a float)
#define RJ_EXIT(L, C, NS)                                   \
  "EXIT" #L ":\n\t"                                         \
  "add.f32 %14, %14, " NS ";\n\t"                           \
  "mov.s32 %3, i" #C ";\n\t"                                \
  "mov.f32 %15, t" #C ";\n\t"                               \
  "selp.s32 %16, 1, 2, y" #C ";\n\t"                        \
  "@x" #C " mov.s32 %16, 0;\n\t"                            \
  "setp.lt.s32 w, ka" #C ", 0;\n\t"                         \
  "@w add.s32 %16, %16, 4;\n\t"                             \
  "setp.lt.s32 w, kb" #C ", 0;\n\t"                         \
  "@w add.s32 %16, %16, 8;\n\t"                             \
  "mov.f32 %5, aa" #C ";\n\t"                               \
  "mov.f32 %6, sa" #C ";\n\t"                               \
  "mov.f32 %7, ea" #C ";\n\t"                               \
  "mov.f32 %10, ab" #C ";\n\t"                              \
  "mov.f32 %11, sb" #C ";\n\t"                              \
  "mov.f32 %12, eb" #C ";\n\t"                              \
  "bra DONE;\n\t"
#define RJ_REGS(SLOTS)                                                                                               \
      ".reg .pred x0, y0, n0, x1, y1, n1, x2, y2, n2, w, pf;\n\t"                                                    \
      ".reg .f32 d, taa, tab, t0, t1, t2, sa0, sa1, sa2, sb0, sb1, sb2;\n\t"                                         \
      ".reg .f32 ea0, ea1, ea2, eb0, eb1, eb2, aa0, aa1, aa2, ab0, ab1, ab2;\n\t"                                    \
      ".reg .b32 ka0, ka1, ka2, kb0, kb1, kb2, wz;\n\t"                                                              \
      ".reg .b16 hl, hh;\n\t"                                                                                        \
      ".reg .s32 i0, i1, i2;\n\t"                                                                                    \
      ".reg .s64 ad;\n\t"
#define RJ_OPERANDS                                                                                                  \
      : "+f"(d.tx), "+f"(d.ty), "+f"(d.tz), "+r"(d.idx), "+f"(qa.A), "+f"(qa.k), "+f"(qa.s), "+f"(qa.ePrev), "+f"(qa.leg), \
        "+f"(qb.A), "+f"(qb.k), "+f"(qb.s), "+f"(qb.ePrev), "+f"(qb.leg), "+f"(nsteps), "=f"(o.t), "=r"(o.axis), "=f"(qa.tau), "=f"(qb.tau) \
      : "l"(recs), "l"(offB), "f"(ddx), "f"(ddy), "f"(ddz), "r"(stx), "r"(sty), "r"(stz)                              \
      : "memory"
#define RJ_ASM(LD, DEC)                                                                                              \
  asm volatile(                                                                                                      \
      "{\n\t"                                                                                                        \
      RJ_REGS(3)                                                                                                     \
      /* slot 2 = the cells the leg starts in */                                                                     \
      "mov.f32 aa2, %5;\n\t" "mov.f32 sa2, %6;\n\t" "mov.f32 ea2, %7;\n\t"                                           \
      "mov.f32 ab2, %10;\n\t" "mov.f32 sb2, %11;\n\t" "mov.f32 eb2, %12;\n\t"                                        \
      "mov.f32 taa, %4;\n\t" "mov.f32 tab, %9;\n\t" "setp.ne.s32 pf, 0, 0;\n\t"                                      \
      RJ_PEEK(0, "%3", LD)                                                                                           \
      "LOOP:\n\t"                                                                                                    \
      RJ_COMMIT(0) RJ_PEEK(1, "i0", LD) RJ_INTEG(2, 0, DEC)                                                          \
      RJ_COMMIT(1) RJ_PEEK(2, "i1", LD) RJ_INTEG(0, 1, DEC)                                                          \
      RJ_COMMIT(2) RJ_PEEK(0, "i2", LD) RJ_INTEG(1, 2, DEC)                                                          \
      "add.f32 %14, %14, 0f40400000;\n\t"                                                                            \
      "bra LOOP;\n\t"                                                                                                \
      RJ_EXIT(0, 0, "0f3F800000") RJ_EXIT(1, 1, "0f40000000") RJ_EXIT(2, 2, "0f40400000")                                                                            \
      "DONE:\n\t"                                                                                                    \
      "mov.f32 %17, taa;\n\t"                                                                                        \
      "mov.f32 %18, tab;\n\t"                                                                                        \
      "}"                                                                                                            \
      RJ_OPERANDS)
// (Measured and dropped in session 9: a four-slot rotation with the requests of steps n+1 AND n+2 in flight at the wall test of step n --
// bit-identical results, but ptxas turned the predicated commits into FFMA + FSEL: +12 % instructions, -7 % throughput,
// profiles/r01s9_ab_pair_depth.log.)

__device__ __forceinline__ LegOut march_leg_pair(Dda& d, Integ& qa, Integ& qb, float& nsteps, const Rec8* __restrict__ recs,
                                                 const Rec8* __restrict__ recsB, float ddx, float ddy, float ddz, int stx, int sty, int stz) {
  LegOut o;
  const long long offB = (long long)(reinterpret_cast<const char*>(recsB) - reinterpret_cast<const char*>(recs));
  RJ_ASM(RJ_LD8, RJ_DEC8);
  return o;
}
__device__ __forceinline__ LegOut march_leg_pair(Dda& d, Integ& qa, Integ& qb, float& nsteps, const Rec4* __restrict__ recs,
                                                 const Rec4* __restrict__ recsB, float ddx, float ddy, float ddz, int stx, int sty, int stz) {
  LegOut o;
  const long long offB = (long long)(reinterpret_cast<const char*>(recsB) - reinterpret_cast<const char*>(recs));
  RJ_ASM(RJ_LD4, RJ_DEC4);
  return o;
}
#undef RJ_ASM
#undef RJ_PEEK
#undef RJ_COMMIT
#undef RJ_INTEG
#undef RJ_EXIT
#undef RJ_REGS
#undef RJ_OPERANDS
#undef RJ_LD8
#undef RJ_LD4
#undef RJ_DEC8
#undef RJ_DEC4
#else
// host build (tests/hostsim): the same algorithm statement by statement
template <class RecT>
__device__ __forceinline__ LegOut march_leg_pair(Dda& d, Integ& qa, Integ& qb, float& nsteps, const RecT* __restrict__ recs,
                                                 const RecT* __restrict__ recsB, float ddx, float ddy, float ddz, int stx, int sty, int stz) {
  Peek p1 = dda_peek(d, stx, sty, stz), p2;
  Rec8 a1 = ldrec(recs + p1.idx), b1 = ldrec(recsB + p1.idx), a2, b2;
  const float n0 = qa.nsteps, n1 = qb.nsteps;
  LegOut o;
  for (;;) {
    dda_commit(d, p1, ddx, ddy, ddz);
    p2 = dda_peek(d, stx, sty, stz);
    a2 = ldrec(recs + p2.idx); b2 = ldrec(recsB + p2.idx);
    integ_cross(qa, p1.t, a1); integ_cross(qb, p1.t, b1);
    if (is_wall(a1.k) || is_wall(b1.k)) { o = LegOut{p1.t, (p1.px ? 0 : (p1.py ? 1 : 2)) | (is_wall(a1.k) ? 4 : 0) | (is_wall(b1.k) ? 8 : 0)}; break; }
    dda_commit(d, p2, ddx, ddy, ddz);
    p1 = dda_peek(d, stx, sty, stz);
    a1 = ldrec(recs + p1.idx); b1 = ldrec(recsB + p1.idx);
    integ_cross(qa, p2.t, a2); integ_cross(qb, p2.t, b2);
    if (is_wall(a2.k) || is_wall(b2.k)) { o = LegOut{p2.t, (p2.px ? 0 : (p2.py ? 1 : 2)) | (is_wall(a2.k) ? 4 : 0) | (is_wall(b2.k) ? 8 : 0)}; break; }
  }
  nsteps += qa.nsteps - n0;
  qa.nsteps = n0; qb.nsteps = n1;
  return o;
}
#endif

// offB = record offset of cell B from cell A (one row, or one plane).
// REJOIN: when both rays enter a non-flow cell at the SAME step -- a planar wall not normal to the pair's axis, e.g. the x- and y-
// mirrors of RMCRT_1L_perf.ups for a pair stacked in z -- and both are reflected, the reflected rays are again exact translates
// (RMCRTCommon::reflect flips the same axis for both and tMax is never recomputed): the joint leg goes on instead of handing over to
// two single marches.  Same per-ray arithmetic either way.
template <bool REJOIN, class RecT>
__device__ __forceinline__ void march_single_level_dual(const RecT* __restrict__ recs, int idxA, int offB, int sy, int sz, float dx, float dy,
                                                        float dz, float rx, float ry, float rz, float cx, float cy, float cz, float threshold,
                                                        bool allowReflect, bool skipZero, unsigned& stepsA, unsigned& stepsB, int recBytes,
                                                        float& sumA, float& sumB) {
  typedef Num<float> N;
  sumA = sumB = 0.0f;
  if (!(1.0f > threshold)) return;
  const bool posx = !sign_bit(dx), posy = !sign_bit(dy), posz = !sign_bit(dz);
  int stx = posx ? 1 : -1, sty = posy ? sy : -sy, stz = posz ? sz : -sz;
  const float ix = N::rcp(dx), iy = N::rcp(dy), iz = N::rcp(dz);
  Dda d;
  d.tx = ((posx ? cx : 0.0f) - rx) * ix * N::LEN_SCALE;
  d.ty = ((posy ? cy : 0.0f) - ry) * iy * N::LEN_SCALE;
  d.tz = ((posz ? cz : 0.0f) - rz) * iz * N::LEN_SCALE;
  const float ddx = fabsf(ix) * cx * N::LEN_SCALE, ddy = fabsf(iy) * cy * N::LEN_SCALE, ddz = fabsf(iz) * cz * N::LEN_SCALE;
  d.idx = idxA;
  const RecT* __restrict__ recsB = recs + offB;
  Integ qa, qb;
  {
    const Rec8 ra = ldrec(recs + idxA), rb = ldrec(recsB + idxA);
    qa.k = fabsf(ra.k); qa.s = ra.s;
    qb.k = fabsf(rb.k); qb.s = rb.s;
  }
  qa.A = qb.A = 0.0f; qa.ePrev = qb.ePrev = 1.0f; qa.nsteps = qb.nsteps = 0.0f; qa.tau = qb.tau = 0.0f; qa.leg = qb.leg = 0.0f;
  float joint = 0.0f, fsA = 1.0f, fsB = 1.0f;
  bool goA = true, goB = true;     // the ray still marches
  bool wa, wb;                     // the ray has a pending wall (lw) that has not been dealt with yet
  LegOut lw;
  for (;;) {
    // ---- joint leg (march_leg_pair): both rays take the same steps until either enters a non-flow cell
    const LegOut jw = march_leg_pair(d, qa, qb, joint, recs, recsB, ddx, ddy, ddz, stx, sty, stz);
    wa = (jw.axis & 4) != 0; wb = (jw.axis & 8) != 0;
    lw = LegOut{jw.t, jw.axis & 3};
    if (!REJOIN || !(wa && wb) || !allowReflect) break;
    // a wall for both at once: wall emission and threshold test per ray (RMCRTCommon.cc:846-898), as in ray_wall
    {
      sumA = fmaf(qa.leg, fsA, sumA); qa.leg = 0.0f;
      const float eps = qa.k > 1.0f ? 1.0f : qa.k;
      sumA = fmaf(eps * qa.s, qa.ePrev, sumA);
      goA = qa.ePrev * fsA > threshold;
      if (goA) { fsA *= (1.0f - qa.k); if (skipZero && fsA == 0.0f) goA = false; }
    }
    {
      sumB = fmaf(qb.leg, fsB, sumB); qb.leg = 0.0f;
      const float eps = qb.k > 1.0f ? 1.0f : qb.k;
      sumB = fmaf(eps * qb.s, qb.ePrev, sumB);
      goB = qb.ePrev * fsB > threshold;
      if (goB) { fsB *= (1.0f - qb.k); if (skipZero && fsB == 0.0f) goB = false; }
    }
    wa = wb = false;               // the wall has been dealt with for both
    if (!goA && !goB) break;
    // RMCRTCommon::reflect (581-602), once for the shared DDA: back to the cells the rays came from, flip the stepping axis
    d.idx -= lw.axis == 0 ? stx : (lw.axis == 1 ? sty : stz);
    if (lw.axis == 0) stx = -stx; else if (lw.axis == 1) sty = -sty; else stz = -stz;
    {
      const Rec8 ra = ldrec(recs + d.idx), rb = ldrec(recsB + d.idx);
      qa.k = fabsf(ra.k); qa.s = ra.s; qa.A = fmaf(-qa.k, lw.t, qa.tau);
      qb.k = fabsf(rb.k); qb.s = rb.s; qb.A = fmaf(-qb.k, lw.t, qb.tau);
    }
    if (!(goA && goB)) break;      // one ray ended here: the other finishes on its own (already reflected)
  }
  qa.nsteps += joint; qb.nsteps += joint;
  // ---- each ray that still marches on its own from here
  if (goA) {
    RayState A{d, qa, sumA, fsA, stx, sty, stz};
    if (!wa || ray_wall(A, lw, recs, threshold, allowReflect, skipZero)) ray_finish(A, recs, ddx, ddy, ddz, threshold, allowReflect, skipZero, recBytes);
    sumA = A.sum; qa.nsteps = A.q.nsteps;
  }
  if (goB) {
    RayState B{d, qb, sumB, fsB, stx, sty, stz};
    if (!wb || ray_wall(B, lw, recsB, threshold, allowReflect, skipZero)) ray_finish(B, recsB, ddx, ddy, ddz, threshold, allowReflect, skipZero, recBytes);
    sumB = B.sum; qb.nsteps = B.q.nsteps;
  }
  stepsA += (unsigned)qa.nsteps;
  stepsB += (unsigned)qb.nsteps;
}

//------------------------------------------------------------------------------------------------------
// single level, FOUR cells per lane (cellsPerLane == 4): the 2 x 2 block of cells (i,j,k), (i,j+1,k), (i,j,k+1), (i,j+1,k+1) of a lane
// traces four exact translates of one ray, so the 16 instructions of a step that do not depend on the ray (DDA peek / commit, index,
// first address, step counter, loop control) are paid once per FOUR ray-cell steps: ~11 SASS instructions per ray-cell step instead of
// ~15 for the pair.  Same structure as the pair: a joint leg (march_leg_quad, PTX, three-slot rotation, the four records of step n+1 in
// flight) until any of the four rays enters a non-flow cell, then every ray finishes on its own with the single-ray leg loop.  Per-ray
// arithmetic is that of march_single_level_fast, statement by statement.
//------------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
// operands: %0-%2 tMax, %3 idx; ray a %4 A %5 k %6 s %7 ePrev %8 leg; ray b %9-%13; ray c %14-%18; ray d %19-%23; %24 nsteps;
// %25 o.t, %26 o.axis, %27-%30 tau a..d; %31 recs, %32-%34 record offsets of rays b, c, d; %35-%37 tDelta; %38-%40 index strides
#define RQ_LD8(J)                                                                                   \
  "mad.wide.s32 pa, i" #J ", 8, %31;\n\t"  "ld.global.nc.v2.f32 {ka" #J ", sa" #J "}, [pa];\n\t"    \
  "add.s64 pb, pa, %32;\n\t"             "ld.global.nc.v2.f32 {kb" #J ", sb" #J "}, [pb];\n\t"    \
  "add.s64 pb, pa, %33;\n\t"             "ld.global.nc.v2.f32 {kc" #J ", sc" #J "}, [pb];\n\t"    \
  "add.s64 pb, pa, %34;\n\t"             "ld.global.nc.v2.f32 {kd" #J ", sd" #J "}, [pb];\n\t"
#define RQ_LD4(J)                                                                                   \
  "mad.wide.s32 pa, i" #J ", 4, %31;\n\t"  "ld.global.nc.b32 ka" #J ", [pa];\n\t"                   \
  "add.s64 pb, pa, %32;\n\t"             "ld.global.nc.b32 kb" #J ", [pb];\n\t"                   \
  "add.s64 pb, pa, %33;\n\t"             "ld.global.nc.b32 kc" #J ", [pb];\n\t"                   \
  "add.s64 pb, pa, %34;\n\t"             "ld.global.nc.b32 kd" #J ", [pb];\n\t"
#define RQ_DEC8_1(R, C) "abs.f32 a" #R #C ", k" #R #C ";\n\t"
#define RQ_DEC4_1(R, C) "mov.b32 {hl, hh}, k" #R #C ";\n\t" "cvt.f32.f16 s" #R #C ", hl;\n\t" "cvt.f32.f16 a" #R #C ", hh;\n\t" "abs.f32 a" #R #C ", a" #R #C ";\n\t"
#define RQ_PEEK(J, IB, LD)                                  \
  "min.f32 t" #J ", %0, %1, %2;\n\t"                        \
  "setp.neu.f32 n" #J ", t" #J ", %2;\n\t"                  \
  "setp.lt.and.f32 x" #J ", %0, %1, n" #J ";\n\t"           \
  "setp.ge.and.f32 y" #J ", %0, %1, n" #J ";\n\t"           \
  "selp.s32 i" #J ", %38, %40, x" #J ";\n\t"                \
  "@y" #J " mov.s32 i" #J ", %39;\n\t"                      \
  "add.s32 i" #J ", i" #J ", " IB ";\n\t"                   \
  LD(J)
#define RQ_COMMIT(J)                                        \
  "@x" #J " add.f32 %0, %0, %35;\n\t"                       \
  "@y" #J " add.f32 %1, %1, %36;\n\t"                       \
  "@!n" #J " add.f32 %2, %2, %37;\n\t"
// one ray: leave the cell in slot P after the path dt inside it (LEG = operand of the leg integral).  The quad leg starts at the ray
// origin (t = 0), so it can carry tau itself instead of the offset form of the single-ray legs: tau += |k_P| dt with ONE dt = t<C> - t<P>
// for the four rays -- one FFMA per ray and step instead of two.  (Same quantity, other FP32 rounding: a cell's divQ differs from the
// single-ray kernels' by round-off, ~1e-6 relative.)
#define RQ_INT1(R, P, C, LEG)                               \
  "fma.rn.f32 t" #R ", a" #R #P ", dt, t" #R ";\n\t"        \
  "neg.f32 d, t" #R ";\n\t"                                 \
  "ex2.approx.ftz.f32 e" #R #C ", d;\n\t"                   \
  "sub.f32 d, e" #R #P ", e" #R #C ";\n\t"                  \
  "fma.rn.f32 " LEG ", s" #R #P ", d, " LEG ";\n\t"
#define RQ_INTEG(P, C, DEC1)                                \
  "sub.f32 dt, t" #C ", t" #P ";\n\t"                       \
  RQ_INT1(a, P, C, "%8") RQ_INT1(b, P, C, "%13") RQ_INT1(c, P, C, "%18") RQ_INT1(d, P, C, "%23") \
  DEC1(a, C) DEC1(b, C) DEC1(c, C) DEC1(d, C)               \
  "lop3.b32 wz, ka" #C ", kb" #C ", kc" #C ", 0xFE;\n\t"      \
  "lop3.or.b32 wy|w, wz, kd" #C ", 0x80000000, 0xA8, pf;\n\t" \
  "@w bra EXIT" #C ";\n\t"
#define RQ_OUT1(R, C, A, K, S, E)                           \
  "neg.f32 d, a" #R #C ";\n\t"                              \
  "fma.rn.f32 " A ", d, t" #C ", t" #R ";\n\t"              \
  "mov.f32 " K ", a" #R #C ";\n\t"                          \
  "mov.f32 " S ", s" #R #C ";\n\t"                          \
  "mov.f32 " E ", e" #R #C ";\n\t"
#define RQ_FLAG1(R, C, BIT)                                 \
  "setp.lt.s32 w, k" #R #C ", 0;\n\t"                       \
  "@w add.s32 %26, %26, " #BIT ";\n\t"
// (steps are counted once per loop iteration of three; an exit adds the steps of the iteration it leaves: NS = C + 1 as a float)
#define RQ_EXIT(C, NS)                                      \
  "EXIT" #C ":\n\t"                                         \
  "add.f32 %24, %24, " NS ";\n\t"                           \
  "mov.s32 %3, i" #C ";\n\t"                                \
  "mov.f32 %25, t" #C ";\n\t"                               \
  "selp.s32 %26, 1, 2, y" #C ";\n\t"                        \
  "@x" #C " mov.s32 %26, 0;\n\t"                            \
  RQ_FLAG1(a, C, 4) RQ_FLAG1(b, C, 8) RQ_FLAG1(c, C, 16) RQ_FLAG1(d, C, 32) \
  RQ_OUT1(a, C, "%4", "%5", "%6", "%7") RQ_OUT1(b, C, "%9", "%10", "%11", "%12") RQ_OUT1(c, C, "%14", "%15", "%16", "%17") RQ_OUT1(d, C, "%19", "%20", "%21", "%22") \
  "bra DONE;\n\t"
#define RQ_START1(R, K, S, E) "mov.f32 a" #R "2, " K ";\n\t" "mov.f32 s" #R "2, " S ";\n\t" "mov.f32 e" #R "2, " E ";\n\t"
#define RQ_ASM(LD, DEC1)                                                                                             \
  asm volatile(                                                                                                      \
      "{\n\t"                                                                                                        \
      ".reg .pred x0, y0, n0, x1, y1, n1, x2, y2, n2, w, pf;\n\t"                                                    \
      ".reg .f32 d, dt, ta, tb, tc, td, t0, t1, t2;\n\t"                                                                 \
      ".reg .f32 sa0, sa1, sa2, sb0, sb1, sb2, sc0, sc1, sc2, sd0, sd1, sd2;\n\t"                                    \
      ".reg .f32 ea0, ea1, ea2, eb0, eb1, eb2, ec0, ec1, ec2, ed0, ed1, ed2;\n\t"                                    \
      ".reg .f32 aa0, aa1, aa2, ab0, ab1, ab2, ac0, ac1, ac2, ad0, ad1, ad2;\n\t"                                    \
      ".reg .b32 ka0, ka1, ka2, kb0, kb1, kb2, kc0, kc1, kc2, kd0, kd1, kd2, wz, wy;\n\t"                            \
      ".reg .b16 hl, hh;\n\t"                                                                                        \
      ".reg .s32 i0, i1, i2;\n\t"                                                                                    \
      ".reg .s64 pa, pb;\n\t"                                                                                        \
      /* slot 2 = the cells the leg starts in */                                                                     \
      RQ_START1(a, "%5", "%6", "%7") RQ_START1(b, "%10", "%11", "%12") RQ_START1(c, "%15", "%16", "%17") RQ_START1(d, "%20", "%21", "%22") \
      "mov.f32 ta, %4;\n\t" "mov.f32 tb, %9;\n\t" "mov.f32 tc, %14;\n\t" "mov.f32 td, %19;\n\t"                      \
      "setp.ne.s32 pf, 0, 0;\n\t" "mov.f32 t2, 0f00000000;\n\t" /* the leg starts at t = 0 with tau = A */                                                                                   \
      RQ_PEEK(0, "%3", LD)                                                                                           \
      "LOOP:\n\t"                                                                                                    \
      RQ_COMMIT(0) RQ_PEEK(1, "i0", LD) RQ_INTEG(2, 0, DEC1)                                                         \
      RQ_COMMIT(1) RQ_PEEK(2, "i1", LD) RQ_INTEG(0, 1, DEC1)                                                         \
      RQ_COMMIT(2) RQ_PEEK(0, "i2", LD) RQ_INTEG(1, 2, DEC1)                                                         \
      "add.f32 %24, %24, 0f40400000;\n\t"                                                                            \
      "bra LOOP;\n\t"                                                                                                \
      RQ_EXIT(0, "0f3F800000") RQ_EXIT(1, "0f40000000") RQ_EXIT(2, "0f40400000")                                                                             \
      "DONE:\n\t"                                                                                                    \
      "mov.f32 %27, ta;\n\t" "mov.f32 %28, tb;\n\t" "mov.f32 %29, tc;\n\t" "mov.f32 %30, td;\n\t"                    \
      "}"                                                                                                            \
      : "+f"(d.tx), "+f"(d.ty), "+f"(d.tz), "+r"(d.idx),                                                             \
        "+f"(q[0].A), "+f"(q[0].k), "+f"(q[0].s), "+f"(q[0].ePrev), "+f"(q[0].leg),                                  \
        "+f"(q[1].A), "+f"(q[1].k), "+f"(q[1].s), "+f"(q[1].ePrev), "+f"(q[1].leg),                                  \
        "+f"(q[2].A), "+f"(q[2].k), "+f"(q[2].s), "+f"(q[2].ePrev), "+f"(q[2].leg),                                  \
        "+f"(q[3].A), "+f"(q[3].k), "+f"(q[3].s), "+f"(q[3].ePrev), "+f"(q[3].leg),                                  \
        "+f"(nsteps), "=f"(o.t), "=r"(o.axis), "=f"(q[0].tau), "=f"(q[1].tau), "=f"(q[2].tau), "=f"(q[3].tau)         \
      : "l"(recs), "l"(bB), "l"(bC), "l"(bD), "f"(ddx), "f"(ddy), "f"(ddz), "r"(stx), "r"(sty), "r"(stz)       \
      : "memory")

// Returns axis | 4 << r for every ray r that entered a non-flow cell in LegOut::axis.
__device__ __forceinline__ LegOut march_leg_quad(Dda& d, Integ (&q)[4], float& nsteps, const Rec8* __restrict__ recs, int offB, int offC,
                                                 int offD, float ddx, float ddy, float ddz, int stx, int sty, int stz) {
  LegOut o;
  const long long bB = (long long)offB * 8, bC = (long long)offC * 8, bD = (long long)offD * 8;     // byte offsets of rays b, c, d
  RQ_ASM(RQ_LD8, RQ_DEC8_1);
  return o;
}
__device__ __forceinline__ LegOut march_leg_quad(Dda& d, Integ (&q)[4], float& nsteps, const Rec4* __restrict__ recs, int offB, int offC,
                                                 int offD, float ddx, float ddy, float ddz, int stx, int sty, int stz) {
  LegOut o;
  const long long bB = (long long)offB * 4, bC = (long long)offC * 4, bD = (long long)offD * 4;
  RQ_ASM(RQ_LD4, RQ_DEC4_1);
  return o;
}
#undef RQ_ASM
#undef RQ_START1
#undef RQ_EXIT
#undef RQ_FLAG1
#undef RQ_OUT1
#undef RQ_INTEG
#undef RQ_INT1
#undef RQ_COMMIT
#undef RQ_PEEK
#undef RQ_DEC4_1
#undef RQ_DEC8_1
#undef RQ_LD4
#undef RQ_LD8
#else
// host build (tests/hostsim): the same algorithm statement by statement
template <class RecT>
__device__ __forceinline__ LegOut march_leg_quad(Dda& d, Integ (&q)[4], float& nsteps, const RecT* __restrict__ recs, int offB, int offC,
                                                 int offD, float ddx, float ddy, float ddz, int stx, int sty, int stz) {
  const int off[4] = {0, offB, offC, offD};
  Peek p = dda_peek(d, stx, sty, stz);
  Rec8 rec[4];
  float tau[4], tPrev = 0.0f;                 // the leg starts at the ray origin: t = 0, tau = A
  for (int r = 0; r < 4; ++r) { rec[r] = ldrec(recs + off[r] + p.idx); tau[r] = q[r].A; }
  LegOut o;
  for (;;) {
    dda_commit(d, p, ddx, ddy, ddz);
    const Peek pn = dda_peek(d, stx, sty, stz);
    Rec8 nxt[4];
    for (int r = 0; r < 4; ++r) nxt[r] = ldrec(recs + off[r] + pn.idx);      // speculative, like the device loop
    const float dt = p.t - tPrev;
    int walls = 0;
    for (int r = 0; r < 4; ++r) {
      tau[r] = fmaf(q[r].k, dt, tau[r]);
      const float e = Num<float>::attenuation(tau[r]);
      q[r].leg = fmaf(q[r].s, q[r].ePrev - e, q[r].leg);
      q[r].ePrev = e;
      q[r].k = fabsf(rec[r].k); q[r].s = rec[r].s;
      if (is_wall(rec[r].k)) walls |= 4 << r;
    }
    nsteps += 1.0f;
    tPrev = p.t;
    if (walls) { o = LegOut{p.t, (p.px ? 0 : (p.py ? 1 : 2)) | walls}; break; }
    p = pn;
    for (int r = 0; r < 4; ++r) rec[r] = nxt[r];
  }
  for (int r = 0; r < 4; ++r) { q[r].tau = tau[r]; q[r].A = fmaf(-q[r].k, o.t, tau[r]); }      // offset form for the legs that follow
  return o;
}
#endif

// cells: A = idxA, B = A + one row, C = A + one plane, D = A + one row + one plane
template <class RecT>
__device__ __forceinline__ void march_single_level_quad(const RecT* __restrict__ recs, int idxA, int sy, int sz, float dx, float dy, float dz,
                                                        float rx, float ry, float rz, float cx, float cy, float cz, float threshold,
                                                        bool allowReflect, bool skipZero, unsigned& steps, int recBytes, float (&sum)[4]) {
  typedef Num<float> N;
  sum[0] = sum[1] = sum[2] = sum[3] = 0.0f;
  if (!(1.0f > threshold)) return;
  const bool posx = !sign_bit(dx), posy = !sign_bit(dy), posz = !sign_bit(dz);
  const int stx = posx ? 1 : -1, sty = posy ? sy : -sy, stz = posz ? sz : -sz;
  const float ix = N::rcp(dx), iy = N::rcp(dy), iz = N::rcp(dz);
  Dda d;
  d.tx = ((posx ? cx : 0.0f) - rx) * ix * N::LEN_SCALE;
  d.ty = ((posy ? cy : 0.0f) - ry) * iy * N::LEN_SCALE;
  d.tz = ((posz ? cz : 0.0f) - rz) * iz * N::LEN_SCALE;
  const float ddx = fabsf(ix) * cx * N::LEN_SCALE, ddy = fabsf(iy) * cy * N::LEN_SCALE, ddz = fabsf(iz) * cz * N::LEN_SCALE;
  d.idx = idxA;
  const int off[4] = {0, sy, sz, sy + sz};
  Integ q[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const Rec8 rec = ldrec(recs + off[r] + idxA);
    q[r].k = fabsf(rec.k); q[r].s = rec.s;
    q[r].A = 0.0f; q[r].ePrev = 1.0f; q[r].nsteps = 0.0f; q[r].tau = 0.0f; q[r].leg = 0.0f;
  }
  float joint = 0.0f;
  const LegOut jw = march_leg_quad(d, q, joint, recs, off[1], off[2], off[3], ddx, ddy, ddz, stx, sty, stz);
  const LegOut lw{jw.t, jw.axis & 3};
  // ---- each ray on its own from here
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    q[r].nsteps += joint;
    RayState R{d, q[r], 0.0f, 1.0f, stx, sty, stz};
    const RecT* __restrict__ rr = recs + off[r];
    if (!(jw.axis & (4 << r)) || ray_wall(R, lw, rr, threshold, allowReflect, skipZero))
      ray_finish(R, rr, ddx, ddy, ddz, threshold, allowReflect, skipZero, recBytes);
    sum[r] = R.sum;
    steps += (unsigned)R.q.nsteps;
  }
}

//------------------------------------------------------------------------------------------------------
// multi level (Data Onion)
//------------------------------------------------------------------------------------------------------
struct LevelGeom {
  const void* rec;
  const void* recFull;   // half-packed levels: float {abskg|wall, sigmaT4/pi} copy read by the divQ epilogue only
  int lo[3];        // window low index (records cover [lo, lo+n))
  int n[3];
  int intLo[3], intHi[3];
  int rr[3];        // refinement ratio to the next coarser level
  int regLo[3], regHi[3];   // Ray::computeExtents region on this level
  double dx[3];
  double anchor[3]; // Level::m_anchor
  double off[3];    // anchor[L] - anchor[fine]
};

template <int MAXL>
struct OnionGeom {
  LevelGeom lev[MAXL];
  int maxLevels;
  int roiLo[3], roiHi[3];
};

// Level::mapCellToCoarser (Level.cc:1285-1307): truncating division + (int)fmod(idx, rr) for negative idx
__device__ __forceinline__ int map_to_coarser(int i, int rr) {
  int q = i / rr;
  if (i < 0 && rr > 1) q += (i % rr);   // C '%' truncates like fmod for ints
  return q;
}

//------------------------------------------------------------------------------------------------------
// 2-level Data Onion, production FP32 fast path (same behaviour as march_multi_level<float, RecT, MAXL, false> for
// maxLevels == 2, i.e. BASELINE config C4 / RMCRT_bm1_DO).
//
// The reference copies the fine level's region of interest per patch (getRegion(ROI), Ray.cc:1060-1075).  So do we, on the
// device, into a compact box with a ONE-CELL SHELL whose records carry the wall flag and abskg = inf.  Inside the ROI the
// ray then runs the very same PTX leg loop as the single-level solve (no per-step ROI test, no per-step position update);
// entering the shell ends the leg like a wall does, and the "wall" being infinite identifies it as the ROI exit of
// Ray.cc:2489-2527: map the cell to the coarse level, re-anchor tMax of the stepping axis exactly as the reference does
// (SURVEY Q10), switch strides / tDelta to the coarse level and keep marching there with the same leg loop (level 0's
// region is the whole level, so only walls end it).  The ray position the re-anchoring needs is reconstructed from the
// crossing time instead of being accumulated every step.
//------------------------------------------------------------------------------------------------------
struct RoiBox {
  const void* rec;   // records of the shell-inclusive box [lo, lo+n), x-fastest
  int lo[3], n[3];
};

__device__ __forceinline__ float sel3(int ax, float a, float b, float c) { return ax == 0 ? a : (ax == 1 ? b : c); }
__device__ __forceinline__ int sel3(int ax, int a, int b, int c) { return ax == 0 ? a : (ax == 1 ? b : c); }

template <class RecT, int MAXL>
__device__ __forceinline__ float march_onion2_fast(const OnionGeom<MAXL>& G, const RoiBox& roi, int ci, int cj, int ck,
                                                   float dx, float dy, float dz, float rx, float ry, float rz, float threshold,
                                                   bool allowReflect, bool skipZero, unsigned& nSteps, int recBytes,
                                                   const Dda* d0 = nullptr, const Integ* q0 = nullptr, const LegOut* w0 = nullptr) {
  // d0 / q0 / w0: continue a ray that march_onion2_quad marched through the ROI together with three translates -- the DDA and integral
  // state after their joint leg, and (w0) the non-flow / shell cell this ray entered at its last step, if it did
  typedef Num<float> N;
  if (!(1.0f > threshold)) return 0.0f;
  const LevelGeom& gf = G.lev[1];
  const LevelGeom& gc = G.lev[0];
  const float LN2 = 0.6931471805599453f;
  const bool posx = !sign_bit(dx), posy = !sign_bit(dy), posz = !sign_bit(dz);
  int sy = roi.n[0], sz = roi.n[0] * roi.n[1];
  int stx = posx ? 1 : -1, sty = posy ? sy : -sy, stz = posz ? sz : -sz;
  // sign[] of the reference: 1/0, after a reflection along that axis -1/-0 (RMCRTCommon.cc:599)
  float sgx = posx ? 1.0f : 0.0f, sgy = posy ? 1.0f : 0.0f, sgz = posz ? 1.0f : 0.0f;
  const float ix = N::rcp(dx), iy = N::rcp(dy), iz = N::rcp(dz);     // never refreshed on reflection (Ray.cc:2395)
  const float fx = (float)gf.dx[0], fy = (float)gf.dx[1], fz = (float)gf.dx[2];
  Dda d;
  d.tx = (sgx * fx - rx) * ix * N::LEN_SCALE;
  d.ty = (sgy * fy - ry) * iy * N::LEN_SCALE;
  d.tz = (sgz * fz - rz) * iz * N::LEN_SCALE;
  float ddx = fabsf(ix) * fx * N::LEN_SCALE, ddy = fabsf(iy) * fy * N::LEN_SCALE, ddz = fabsf(iz) * fz * N::LEN_SCALE;
  d.idx = (ci - roi.lo[0]) + sy * (cj - roi.lo[1]) + sz * (ck - roi.lo[2]);
  const RecT* cur = static_cast<const RecT*>(roi.rec);
  Integ q;
  {
    const Rec8 r = ldrec(cur + d.idx);
    q.k = fabsf(r.k); q.s = r.s;
  }
  q.A = 0.0f; q.ePrev = 1.0f; q.nsteps = 0.0f; q.tau = 0.0f; q.leg = 0.0f;
  if (d0) { d = *d0; q = *q0; }
  bool first = d0 != nullptr;
  // position along each axis relative to the origin cell's low corner: l(t) = lb + (t - tb) ln2 * dir   (t is scaled by log2 e)
  float lbx = rx, lby = ry, lbz = rz, tbx = 0.0f, tby = 0.0f, tbz = 0.0f, dvx = dx, dvy = dy, dvz = dz;
  bool onFine = true;
  float sum = 0.0f, fs = 1.0f;
  for (;;) {
    if (!first) q.leg = 0.0f;                          // (a continued ray keeps the integral of its joint leg)
    const LegOut w = (first && w0) ? *w0 : march_leg(d, q, cur, ddx, ddy, ddz, stx, sty, stz, recBytes);
    first = false;
    sum = fmaf(q.leg, fs, sum);
    const int ax = w.axis;
    if (onFine && q.k > 3.0e38f) {                   // stepped into the shell: leave the ROI (Ray.cc:2489-2527)
      int b = d.idx;
      const int bk = b / sz; b -= bk * sz;
      const int bj = b / sy;
      const int bi = b - bj * sy;
      const int c0 = roi.lo[0] + bi, c1 = roi.lo[1] + bj, c2 = roi.lo[2] + bk;     // fine cell just entered
      // re-anchor tMax of the stepping axis from the ray position and the centre of that fine cell (Ray.cc:2486,2534-2550)
      const float fA = sel3(ax, fx, fy, fz);
      const float ccRel = fA * ((float)sel3(ax, c0, c1, c2) + 0.5f) - fA * (float)sel3(ax, ci, cj, ck);
      const float lA = fmaf((w.t - sel3(ax, tbx, tby, tbz)) * LN2, sel3(ax, dvx, dvy, dvz), sel3(ax, lbx, lby, lbz));
      const float cdx = (float)sel3(ax, (float)gc.dx[0], (float)gc.dx[1], (float)gc.dx[2]);
      const float rayDxLevel = lA - (ccRel - 0.5f * cdx);
      const float tTmp = (sel3(ax, sgx, sgy, sgz) * cdx - rayDxLevel) * sel3(ax, ix, iy, iz) * N::LEN_SCALE;
      if (ax == 0) d.tx = w.t + tTmp; else if (ax == 1) d.ty = w.t + tTmp; else d.tz = w.t + tTmp;
      // coarse cell, strides and tDelta of level 0
      const int k0 = min(max(map_to_coarser(c0, gf.rr[0]) - gc.lo[0], 0), gc.n[0] - 1);
      const int k1 = min(max(map_to_coarser(c1, gf.rr[1]) - gc.lo[1], 0), gc.n[1] - 1);
      const int k2 = min(max(map_to_coarser(c2, gf.rr[2]) - gc.lo[2], 0), gc.n[2] - 1);
      sy = gc.n[0]; sz = gc.n[0] * gc.n[1];
      stx = stx > 0 ? 1 : -1; sty = sty > 0 ? sy : -sy; stz = stz > 0 ? sz : -sz;
      ddx = fabsf(ix) * (float)gc.dx[0] * N::LEN_SCALE; ddy = fabsf(iy) * (float)gc.dx[1] * N::LEN_SCALE;
      ddz = fabsf(iz) * (float)gc.dx[2] * N::LEN_SCALE;
      d.idx = k0 + sy * k1 + sz * k2;
      cur = static_cast<const RecT*>(gc.rec);
      onFine = false;
      const Rec8 r = ldrec(cur + d.idx);
      q.k = fabsf(r.k); q.s = r.s;
      q.A = fmaf(-q.k, w.t, q.tau);
      if (!is_wall(r.k)) continue;                   // keep marching on the coarse level
    }
    // wall (fine or coarse): (q.k, q.s) = record of the non-flow cell entered
    const float eps = q.k > 1.0f ? 1.0f : q.k;                 // Ray.cc:2594-2598
    sum = fmaf(eps * q.s, q.ePrev, sum);                       // 2602: NOT multiplied by fs
    const float intensity = q.ePrev * fs;
    if (!allowReflect || !(intensity > threshold)) break;      // 2607-2609, 2621
    fs *= (1.0f - q.k);
    if (skipZero && fs == 0.0f) break;
    // reflect (RMCRTCommon.cc:581-602): previous cell on the CURRENT level, flip the stepping axis
    d.idx -= sel3(ax, stx, sty, stz);
    const float lHit = fmaf((w.t - sel3(ax, tbx, tby, tbz)) * LN2, sel3(ax, dvx, dvy, dvz), sel3(ax, lbx, lby, lbz));
    if (ax == 0) { stx = -stx; sgx = -sgx; dvx = -dvx; lbx = lHit; tbx = w.t; }
    else if (ax == 1) { sty = -sty; sgy = -sgy; dvy = -dvy; lby = lHit; tby = w.t; }
    else { stz = -stz; sgz = -sgz; dvz = -dvz; lbz = lHit; tbz = w.t; }
    const Rec8 r = ldrec(cur + d.idx);
    q.k = fabsf(r.k); q.s = r.s;
    q.A = fmaf(-q.k, w.t, q.tau);
  }
  nSteps += (unsigned)q.nsteps;
  return sum;
}

// 2-level Data Onion, four cells per lane: the rays of the 2 x 2 cells (ci, cj..cj+1, ck..ck+1) are exact translates inside the ROI copy
// (one array, one cell size), so ONE DDA marches all four there (march_leg_quad) until any of them enters a non-flow cell or the shell;
// from that step on every ray continues on its own through march_onion2_fast (ROI exit, coarse level, reflections).
template <class RecT, int MAXL>
__device__ __forceinline__ void march_onion2_quad(const OnionGeom<MAXL>& G, const RoiBox& roi, int ci, int cj, int ck, float dx, float dy,
                                                  float dz, float rx, float ry, float rz, float threshold, bool allowReflect, bool skipZero,
                                                  unsigned& steps, int recBytes, float (&sum)[4]) {
  typedef Num<float> N;
  sum[0] = sum[1] = sum[2] = sum[3] = 0.0f;
  if (!(1.0f > threshold)) return;
  const LevelGeom& gf = G.lev[1];
  const bool posx = !sign_bit(dx), posy = !sign_bit(dy), posz = !sign_bit(dz);
  const int sy = roi.n[0], sz = roi.n[0] * roi.n[1];
  const int stx = posx ? 1 : -1, sty = posy ? sy : -sy, stz = posz ? sz : -sz;
  const float sgx = posx ? 1.0f : 0.0f, sgy = posy ? 1.0f : 0.0f, sgz = posz ? 1.0f : 0.0f;
  const float ix = N::rcp(dx), iy = N::rcp(dy), iz = N::rcp(dz);
  const float fx = (float)gf.dx[0], fy = (float)gf.dx[1], fz = (float)gf.dx[2];
  Dda d;
  d.tx = (sgx * fx - rx) * ix * N::LEN_SCALE;
  d.ty = (sgy * fy - ry) * iy * N::LEN_SCALE;
  d.tz = (sgz * fz - rz) * iz * N::LEN_SCALE;
  const float ddx = fabsf(ix) * fx * N::LEN_SCALE, ddy = fabsf(iy) * fy * N::LEN_SCALE, ddz = fabsf(iz) * fz * N::LEN_SCALE;
  d.idx = (ci - roi.lo[0]) + sy * (cj - roi.lo[1]) + sz * (ck - roi.lo[2]);
  const RecT* cur = static_cast<const RecT*>(roi.rec);
  const int off[4] = {0, sy, sz, sy + sz};
  Integ q[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const Rec8 rec = ldrec(cur + off[r] + d.idx);
    q[r].k = fabsf(rec.k); q[r].s = rec.s;
    q[r].A = 0.0f; q[r].ePrev = 1.0f; q[r].nsteps = 0.0f; q[r].tau = 0.0f; q[r].leg = 0.0f;
  }
  float joint = 0.0f;
  const LegOut jw = march_leg_quad(d, q, joint, cur, off[1], off[2], off[3], ddx, ddy, ddz, stx, sty, stz);
  const LegOut lw{jw.t, jw.axis & 3};
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    q[r].nsteps += joint;
    Dda dr = d;
    dr.idx += off[r];
    sum[r] = march_onion2_fast<RecT, MAXL>(G, roi, ci, cj + (r & 1), ck + (r >> 1), dx, dy, dz, rx, ry, rz, threshold, allowReflect, skipZero,
                                           steps, recBytes, &dr, &q[r], (jw.axis & (4 << r)) ? &lw : nullptr);
  }
}


template <class Real, class RecT, int MAXL, bool RECORD>
__device__ __forceinline__ Real march_multi_level(const OnionGeom<MAXL>& G, int ci, int cj, int ck,
                                                  Real dx, Real dy, Real dz, Real rx, Real ry, Real rz,
                                                  Real threshold, bool allowReflect, bool skipZero,
                                                  unsigned& nSteps, PathSink* path, bool floatProps = false) {
  typedef Num<Real> N;
  if (!(Real(1) > threshold)) return Real(0);
  int L = G.maxLevels - 1;
  const int F = L;
  const int oi = ci, oj = cj, ok = ck;   // origin cell on the fine level

  const bool posx = !sign_bit(dx), posy = !sign_bit(dy), posz = !sign_bit(dz);
  int stx = posx ? 1 : -1, sty = posy ? 1 : -1, stz = posz ? 1 : -1;
  // sign[] of the reference: 1/0, and after a reflection along that axis -1/-0 (sign *= -1, RMCRTCommon.cc:599)
  Real sgx = posx ? Real(1) : Real(0), sgy = posy ? Real(1) : Real(0), sgz = posz ? Real(1) : Real(0);
  const Real ix = N::rcp(dx), iy = N::rcp(dy), iz = N::rcp(dz);   // never refreshed on reflection (Ray.cc:2395)
  const Real aix = absr(ix), aiy = absr(iy), aiz = absr(iz);

  Real cx = (Real)G.lev[L].dx[0], cy = (Real)G.lev[L].dx[1], cz = (Real)G.lev[L].dx[2];
  Real tx = N::mul(N::mul(N::add(N::mul(sgx, cx), -rx), ix), N::LEN_SCALE);
  Real ty = N::mul(N::mul(N::add(N::mul(sgy, cy), -ry), iy), N::LEN_SCALE);
  Real tz = N::mul(N::mul(N::add(N::mul(sgz, cz), -rz), iz), N::LEN_SCALE);
  Real ddx = N::mul(N::mul(aix, cx), N::LEN_SCALE), ddy = N::mul(N::mul(aiy, cy), N::LEN_SCALE),
       ddz = N::mul(N::mul(aiz, cz), N::LEN_SCALE);

  // ray location relative to the origin cell's low corner, in true length units
  Real lx = rx, ly = ry, lz = rz;
  Real dirx = dx, diry = dy, dirz = dz;
  Real tPrev = 0, tau = 0, ePrev = 1, sum = 0, fs = 1;
  bool onFine = true;
  unsigned steps = 0;

  auto recIndex = [&](int l, int i, int j, int kk) -> long long {
    const LevelGeom& g = G.lev[l];
    // clamp: only reachable through the reference's own out-of-range quirk
    int a = min(max(i - g.lo[0], 0), g.n[0] - 1), b = min(max(j - g.lo[1], 0), g.n[1] - 1),
        c = min(max(kk - g.lo[2], 0), g.n[2] - 1);
    return (long long)a + (long long)g.n[0] * ((long long)b + (long long)g.n[1] * (long long)c);
  };

  const RecT* recs = static_cast<const RecT*>(G.lev[L].rec);
  typename RecInfo<RecT>::decoded r = ldrec(recs + recIndex(L, ci, cj, ck));
  Real k = absr(r.k), s = r.s;

  for (;;) {
    int pi, pj, pk;
    bool px, py, stop;
    Real e;
    do {
      pi = ci; pj = cj; pk = ck;
      const bool xy = tx < ty;
      px = xy && (tx < tz);
      py = !xy && (ty < tz);
      const bool pz = !(px || py);
      const Real tNext = px ? tx : (py ? ty : tz);
      ci += px ? stx : 0; cj += py ? sty : 0; ck += pz ? stz : 0;
      const int ax = px ? 0 : (py ? 1 : 2);
      const int cA = px ? ci : (py ? cj : ck);

      const LevelGeom& g = G.lev[L];
      // domain_BB.inside(level->getCellPosition(cur)) (Ray.cc:2486-2487): only the stepped axis can change the
      // answer, and a cell centre is inside the interior box exactly when its index is an interior index.
      bool inDomain = (cA >= g.intLo[ax]) && (cA < g.intHi[ax]);
      const bool outsideROI = !(G.roiLo[ax] <= cA && G.roiHi[ax] > cA);          // 2489
      const bool outsideReg = !(g.regLo[ax] <= cA && g.regHi[ax] > cA);          // 2490
      const bool jumpFine = onFine && outsideROI && inDomain;                     // 2492
      const bool jumpCoarse = (!onFine) && outsideReg && (L > 0) && inDomain;     // 2493
      const bool jump = jumpFine || jumpCoarse;
      // CC_pos[ax] of the cell entered, on the PRE-jump level (Ray.cc:2486), relative to the origin corner
      const int oA = px ? oi : (py ? oj : ok);
      const Real cA_dx = px ? cx : (py ? cy : cz);
      const Real ccRel = N::add(N::add((Real)g.off[ax], N::mul(cA_dx, N::add((Real)cA, Real(0.5)))),
                                -N::mul((Real)G.lev[F].dx[ax], (Real)oA));
      if (jump) {                                                                 // 2505-2527
        ci = map_to_coarser(ci, g.rr[0]); cj = map_to_coarser(cj, g.rr[1]); ck = map_to_coarser(ck, g.rr[2]);
        L = L - 1;
        onFine = false;
        cx = (Real)G.lev[L].dx[0]; cy = (Real)G.lev[L].dx[1]; cz = (Real)G.lev[L].dx[2];
        ddx = N::mul(N::mul(aix, cx), N::LEN_SCALE); ddy = N::mul(N::mul(aiy, cy), N::LEN_SCALE);
        ddz = N::mul(N::mul(aiz, cz), N::LEN_SCALE);
        recs = static_cast<const RecT*>(G.lev[L].rec);
      }
      r = ldrec(recs + recIndex(L, ci, cj, ck));

      const Real dsS = N::add(tNext, -tPrev);            // distanceTraveled (scaled), 2534
      tPrev = tNext;
      const Real ds = (N::LEN_SCALE == Real(1)) ? dsS : N::mul(dsS, Real(0.6931471805599453));
      lx = N::add(lx, N::mul(ds, dirx)); ly = N::add(ly, N::mul(ds, diry)); lz = N::add(lz, N::mul(ds, dirz));  // 2539-2541
      if (!jump) {
        tx = px ? N::add(tx, ddx) : tx; ty = py ? N::add(ty, ddy) : ty; tz = pz ? N::add(tz, ddz) : tz;   // 2537
      } else {                                           // 2545-2550: only tMax[ax] is re-anchored
        const Real lA = px ? lx : (py ? ly : lz);
        const Real cdx = px ? cx : (py ? cy : cz);       // coarse Dx[L][ax]
        const Real sg = px ? sgx : (py ? sgy : sgz);
        const Real inv = px ? ix : (py ? iy : iz);
        const Real rayDxLevel = N::add(lA, -N::add(ccRel, -N::mul(Real(0.5), cdx)));
        const Real tTmp = N::mul(N::mul(N::add(N::mul(sg, cdx), -rayDxLevel), inv), N::LEN_SCALE);
        tx = px ? N::add(tx, tTmp) : tx; ty = py ? N::add(ty, tTmp) : ty; tz = pz ? N::add(tz, tTmp) : tz;
      }
      inDomain = inDomain && !is_wall(r.k);              // 2554
      tau = N::add(tau, N::mul(k, dsS));                 // 2557
      ++steps;
      e = N::attenuation(tau);
      sum = N::add(sum, N::mul(N::mul(s, N::add(ePrev, -e)), fs));   // 2588
      ePrev = e;
      k = absr(r.k);
      s = r.s;
      if (RECORD && path && path->n < path->cap) {
        int32_t* q = path->buf + 4 * path->n; q[0] = ci; q[1] = cj; q[2] = ck; q[3] = L;
      }
      if (RECORD && path) path->n++;
      stop = !inDomain;
    } while (!stop);

    const Real eps = k > Real(1) ? Real(1) : k;          // 2594-2598
    Real intensity = e;
    // Ray.cc:2595,2602: here `double wallEmissivity = abskg[L][cur]`, so the product with sigmaT4OverPi is evaluated in double even for
    // T = float (only the single-level updateSumI keeps `T wallEmissivity`, RMCRTCommon.cc:846-854)
    sum = N::add(sum, N::mul(N::wall_product(eps, s, false), intensity));
    intensity = N::mul(intensity, fs);
    if (!allowReflect || !(intensity > threshold)) break; // 2607-2609, 2621
    fs = N::mul(fs, N::add(Real(1), -k));                // reflect(); cur = prevCell with L unchanged
    ci = pi; cj = pj; ck = pk;
    if (px) { stx = -stx; sgx = -sgx; dirx = -dirx; }
    else if (py) { sty = -sty; sgy = -sgy; diry = -diry; }
    else { stz = -stz; sgz = -sgz; dirz = -dirz; }
    if (skipZero && fs == Real(0)) break;
    r = ldrec(recs + recIndex(L, ci, cj, ck));
    k = absr(r.k);
    s = r.s;
  }
  nSteps += steps;
  return sum;
}

}  // namespace rmcrt
