// dvf_math2.cuh -- the per-pixel arithmetic of dvf_math.cuh on PAIRS of pixels, using Blackwell's packed
// fp32x2 instructions (FFMA2 / FMUL2 / FADD2, sm_100+): one issue slot performs the operation for two
// pixels, each half rounded exactly like the scalar IEEE instruction (round-to-nearest-even), so every
// result is bit-identical to the scalar chain.  The kernel is issue-bound, not FLOP-bound: halving the
// floating-point instruction count is what packing buys.
//
// Convention: a float2 holds (pixel A, pixel B).  CTA-uniform operands (P, K^-1, per-level constants) are
// scalars broadcast with dup(): the SASS operand form `R.F32` feeds one register to both lanes for free, and
// negations fold into operand modifiers.
#pragma once
#include "dvf_math.cuh"

namespace dvf {

typedef float2 f2;

__device__ __forceinline__ f2 dup(float x) { return make_float2(x, x); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 neg2(f2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __fadd2_rn(a, neg2(b)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }

__device__ __forceinline__ f2 dot3_2(f2 m0, f2 m1, f2 m2, f2 c0, f2 c1, f2 c2) {
  return fma2(m2, c2, fma2(m1, c1, mul2(m0, c0)));
}

// refined reciprocal and exact quotient, two lanes at once (same op sequence as the scalar versions)
__device__ __forceinline__ f2 rcp_refined2(f2 b, f2 nb /* = -b */) {
  f2 r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(b.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(b.y));
  const f2 t = fma2(nb, r, dup(1.0f));
  return fma2(r, t, r);
}
__device__ __forceinline__ f2 div_by2(f2 a, f2 nb /* = -b */, f2 r) {   // one correction step: see dvf_math.cuh: div_by
  const f2 q = mul2(a, r);
  return fma2(fma2(nb, q, a), r, q);
}

struct Geo2 {          // per-level constants, both halves equal
  f2 nW1, nH1;         // -(W-1), -(H-1)
  f2 rW1, rH1;         // correctly rounded reciprocals
  f2 halfW, halfH;
  f2 offs;
};
__device__ __forceinline__ Geo2 make_geo2(const Geo& g) {
  Geo2 o;
  o.nW1 = dup(-g.fW1);
  o.nH1 = dup(-g.fH1);
  o.rW1 = dup(g.rW1);
  o.rH1 = dup(g.rH1);
  o.halfW = dup(g.halfW);
  o.halfH = dup(g.halfH);
  o.offs = dup(g.offs);
  return o;
}

struct Cam2 {
  f2 ray[3];
  f2 cam[3];
};

// M: K^-1 (9 scalars, broadcast to both lanes for free), d = (depth A, depth B), fi/fj = rows / columns
__device__ __forceinline__ void pixel_to_cam2(const float* __restrict__ M, f2 d, f2 fi, f2 fj, Cam2& o) {
  const f2 one = dup(1.0f);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    o.ray[k] = dot3_2(dup(M[k * 3 + 0]), dup(M[k * 3 + 1]), dup(M[k * 3 + 2]), fj, fi, one);
    o.cam[k] = mul2(o.ray[k], d);
  }
}

struct Proj2 {
  f2 qz, nZ, rZ, u, v, xn, yn;   // nZ = -clamp(qz, 1e-3)
  // zeros padding: a coordinate that was overwritten (|x| > 1 -> 2, gradient killed) is recognised later by its value:
  // it is exactly 2, which no coordinate that was kept (|x| <= 1 or NaN) can be -- no flag registers are carried
};

// hot path only (callers guard the operand range and patch lanes with the scalar exact routine)
template <bool kZeros>
__device__ __forceinline__ void project2(const float* __restrict__ P /*3x4, broadcast*/, const Cam2& c, const Geo2& g,
                                         Proj2& o) {
  const f2 X = add2(dot3_2(dup(P[0]), dup(P[1]), dup(P[2]), c.cam[0], c.cam[1], c.cam[2]), dup(P[3]));
  const f2 Y = add2(dot3_2(dup(P[4]), dup(P[5]), dup(P[6]), c.cam[0], c.cam[1], c.cam[2]), dup(P[7]));
  const f2 Zr = add2(dot3_2(dup(P[8]), dup(P[9]), dup(P[10]), c.cam[0], c.cam[1], c.cam[2]), dup(P[11]));
  o.qz = Zr;
  const f2 Z = make_float2(fmaxf(Zr.x, kMinDepthZ), fmaxf(Zr.y, kMinDepthZ));
  o.nZ = neg2(Z);
  o.rZ = rcp_refined2(Z, o.nZ);
  o.u = div_by2(X, o.nZ, o.rZ);
  o.v = div_by2(Y, o.nZ, o.rZ);
  const f2 mone = dup(-1.0f);
  o.xn = add2(div_by2(add2(o.u, o.u), g.nW1, g.rW1), mone);
  o.yn = add2(div_by2(add2(o.v, o.v), g.nH1, g.rH1), mone);
  if (kZeros) {
    o.xn = make_float2(fabsf(o.xn.x) > 1.0f ? 2.0f : o.xn.x, fabsf(o.xn.y) > 1.0f ? 2.0f : o.xn.y);
    o.yn = make_float2(fabsf(o.yn.x) > 1.0f ? 2.0f : o.yn.x, fabsf(o.yn.y) > 1.0f ? 2.0f : o.yn.y);
  }
}

struct Loc2 {
  int x0A, y0A, x0B, y0B;
  f2 w, e, n, s;
  f2 gmx, gmy;
  bool nwA, neA, swA, seA, nwB, neB, swB, seB;
};

template <bool kZeros>
__device__ __forceinline__ void locate2(f2 xn, f2 yn, int H, int W, const Geo& gs, const Geo2& g, Loc2& L) {
  const f2 one = dup(1.0f);
  f2 ix = fma2(add2(xn, one), g.halfW, g.offs);
  f2 iy = fma2(add2(yn, one), g.halfH, g.offs);
  L.gmx = g.halfW;
  L.gmy = g.halfH;
  if (!kZeros) {  // border padding: clip_coordinates(_set_grad)
    if (ix.x <= 0.0f) { ix.x = 0.0f; L.gmx.x = 0.0f; } else if (ix.x >= gs.fW1) { ix.x = gs.fW1; L.gmx.x = 0.0f; }
    if (ix.y <= 0.0f) { ix.y = 0.0f; L.gmx.y = 0.0f; } else if (ix.y >= gs.fW1) { ix.y = gs.fW1; L.gmx.y = 0.0f; }
    if (iy.x <= 0.0f) { iy.x = 0.0f; L.gmy.x = 0.0f; } else if (iy.x >= gs.fH1) { iy.x = gs.fH1; L.gmy.x = 0.0f; }
    if (iy.y <= 0.0f) { iy.y = 0.0f; L.gmy.y = 0.0f; } else if (iy.y >= gs.fH1) { iy.y = gs.fH1; L.gmy.y = 0.0f; }
  }
  const f2 fx = make_float2(floorf(ix.x), floorf(ix.y)), fy = make_float2(floorf(iy.x), floorf(iy.y));
  L.w = sub2(ix, fx);
  L.e = sub2(one, L.w);
  L.n = sub2(iy, fy);
  L.s = sub2(one, L.n);
  L.x0A = __float2int_rz(fx.x);
  L.x0B = __float2int_rz(fx.y);
  L.y0A = __float2int_rz(fy.x);
  L.y0B = __float2int_rz(fy.y);
  {
    const bool y0 = (unsigned)L.y0A < (unsigned)H, y1 = (unsigned)(L.y0A + 1) < (unsigned)H;
    const bool x0 = (unsigned)L.x0A < (unsigned)W, x1 = (unsigned)(L.x0A + 1) < (unsigned)W;
    L.nwA = x0 && y0; L.neA = x1 && y0; L.swA = x0 && y1; L.seA = x1 && y1;
  }
  {
    const bool y0 = (unsigned)L.y0B < (unsigned)H, y1 = (unsigned)(L.y0B + 1) < (unsigned)H;
    const bool x0 = (unsigned)L.x0B < (unsigned)W, x1 = (unsigned)(L.x0B + 1) < (unsigned)W;
    L.nwB = x0 && y0; L.neB = x1 && y0; L.swB = x0 && y1; L.seB = x1 && y1;
  }
}

__device__ __forceinline__ f2 bilerp2(f2 vnw, f2 vne, f2 vsw, f2 vse, f2 wnw, f2 wne, f2 wsw, f2 wse) {
  return fma2(vse, wse, fma2(vsw, wsw, fma2(vne, wne, mul2(vnw, wnw))));
}
__device__ __forceinline__ void bilerp_grad2(f2 vnw, f2 vne, f2 vsw, f2 vse, const Loc2& L, f2 g, f2& gx, f2& gy) {
  const f2 nnw = neg2(vnw);
  gx = fma2(fma2(add2(vse, neg2(vsw)), L.n, mul2(add2(vne, nnw), L.s)), g, gx);
  gy = fma2(fma2(add2(vse, neg2(vne)), L.w, mul2(add2(vsw, nnw), L.e)), g, gy);
}

struct ChainGrad2 {
  f2 gq[3];
  f2 gdepth;
};

// hot path only
template <bool kZeros>
__device__ __forceinline__ void chain_backward2(const float* __restrict__ P, const Cam2& c, const Proj2& p, const Loc2& L,
                                                f2 gx, f2 gy, const Geo2& g, ChainGrad2& o) {
  f2 gxn = mul2(gx, L.gmx), gyn = mul2(gy, L.gmy);
  if (kZeros) {
    gxn = make_float2(p.xn.x == 2.0f ? 0.0f : gxn.x, p.xn.y == 2.0f ? 0.0f : gxn.y);
    gyn = make_float2(p.yn.x == 2.0f ? 0.0f : gyn.x, p.yn.y == 2.0f ? 0.0f : gyn.y);
  }
  const f2 two = dup(2.0f);
  const f2 gu = mul2(div_by2(gxn, g.nW1, g.rW1), two);
  const f2 gv = mul2(div_by2(gyn, g.nH1, g.rH1), two);
  const f2 gq0 = div_by2(gu, p.nZ, p.rZ);
  const f2 gq1 = div_by2(gv, p.nZ, p.rZ);
  const f2 uz = div_by2(p.u, p.nZ, p.rZ);
  const f2 vz = div_by2(p.v, p.nZ, p.rZ);
  // ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (it does not for the scalar forms): wherever the
  // reference rounds a product and then adds, the sum is done with scalar adds on the halves
  const f2 pu = mul2(neg2(gu), uz), pv = mul2(neg2(gv), vz);
  const f2 gZ = make_float2(add(pu.x, pv.x), add(pu.y, pv.y));
  const f2 gq2 = make_float2(p.qz.x >= kMinDepthZ ? gZ.x : 0.0f, p.qz.y >= kMinDepthZ ? gZ.y : 0.0f);
  o.gq[0] = gq0;
  o.gq[1] = gq1;
  o.gq[2] = gq2;
  const f2 m0 = mul2(dot3_2(dup(P[0]), dup(P[4]), dup(P[8]), gq0, gq1, gq2), c.ray[0]);
  const f2 m1 = mul2(dot3_2(dup(P[1]), dup(P[5]), dup(P[9]), gq0, gq1, gq2), c.ray[1]);
  const f2 m2 = mul2(dot3_2(dup(P[2]), dup(P[6]), dup(P[10]), gq0, gq1, gq2), c.ray[2]);
  o.gdepth = make_float2(add(add(m0.x, m1.x), m2.x), add(add(m0.y, m1.y), m2.y));
}

}  // namespace dvf
