// ORCA (RVO2) velocity solve for ONE agent against its neighbours — the arithmetic the
// reference obtains from the external rvo2 library through crowd_nav/policy/orca.py:64-117
// (Agent::computeNeighbors / computeNewVelocity / linearProgram1-3 of RVO2 v2.0.x).
//
// One thread per human.  All arithmetic is IEEE fp32 in the RVO2 operation order; this
// translation unit MUST be compiled with --fmad=false (nvcc) / -ffp-contract=off (g++ test
// harness) so no multiply-add is contracted: that is what makes the result bit-identical to
// the CPU oracle (oracle/rvo2_ref.cpp).  Division and sqrt are the IEEE-rounded ones.
//
// ORCA lines live in a caller-provided store (shared memory on the GPU, [line][thread]
// layout => conflict-free 16-byte accesses); the rarely used LP3 projection lines live in a
// per-thread local array.
#pragma once
#include "cn_common.cuh"
#include "cn_rng.cuh"   // CnCoop and the warp collectives

#define CN_RVO_EPS 0.00001f

struct CnF2 {
  float x, y;
};
CN_HD CnF2 f2(float x, float y) { CnF2 r; r.x = x; r.y = y; return r; }
CN_HD CnF2 f2add(CnF2 a, CnF2 b) { return f2(a.x + b.x, a.y + b.y); }
CN_HD CnF2 f2sub(CnF2 a, CnF2 b) { return f2(a.x - b.x, a.y - b.y); }
CN_HD CnF2 f2neg(CnF2 a) { return f2(-a.x, -a.y); }
CN_HD float f2dot(CnF2 a, CnF2 b) { return a.x * b.x + a.y * b.y; }
CN_HD CnF2 f2scale(float s, CnF2 a) { return f2(s * a.x, s * a.y); }
// RVO2 Vector2::operator/(float): multiply by the reciprocal
CN_HD CnF2 f2div(CnF2 a, float s) { const float inv = 1.0f / s; return f2(a.x * inv, a.y * inv); }
CN_HD float f2abssq(CnF2 a) { return f2dot(a, a); }
CN_HD float f2abs(CnF2 a) { return sqrtf(f2dot(a, a)); }
CN_HD float f2det(CnF2 a, CnF2 b) { return a.x * b.y - a.y * b.x; }
CN_HD CnF2 f2normalize(CnF2 a) { return f2div(a, f2abs(a)); }
// std::min / std::max semantics (matter only for NaN, kept for faithfulness)
CN_HD float cn_minf(float a, float b) { return (b < a) ? b : a; }
CN_HD float cn_maxf(float a, float b) { return (a < b) ? b : a; }

struct CnLine {
  CnF2 point, dir;
};

// Two-tier strided line store.  The first `cap` lines of a thread live in shared memory
// (element k at base[k * stride]: [line][thread] layout, conflict-free 16-byte accesses); lines
// beyond `cap` (only reached when an agent has more than `cap` neighbours inside neighborDist)
// spill to a per-thread overflow array.  Keeping cap < H-1 is what lets all environments of a
// launch be resident in ONE wave (shared memory is the occupancy limiter of the step kernel).
struct CnLineStore {
  float4* base;
  int stride;
  int cap;
  float4* ovf;     // per-thread overflow storage for k >= cap
  CN_HD CnLine get(int k) const {
    const float4 v = (k < cap) ? base[k * stride] : ovf[k - cap];
    CnLine l; l.point = f2(v.x, v.y); l.dir = f2(v.z, v.w); return l;
  }
  CN_HD void set(int k, const CnLine& l) {
    float4 v; v.x = l.point.x; v.y = l.point.y; v.z = l.dir.x; v.w = l.dir.y;
    if (k < cap) base[k * stride] = v; else ovf[k - cap] = v;
  }
};

template <class Lines>
CN_HD bool cn_lp1(const Lines& lines, int lineNo, float radius, CnF2 optVelocity, bool directionOpt,
                  CnF2& result) {
  const CnLine ln = lines.get(lineNo);
  const float dotProduct = f2dot(ln.point, ln.dir);
  const float discriminant = dotProduct * dotProduct + radius * radius - f2abssq(ln.point);
  if (discriminant < 0.0f) return false;
  const float sqrtDiscriminant = sqrtf(discriminant);
  float tLeft = -dotProduct - sqrtDiscriminant;
  float tRight = -dotProduct + sqrtDiscriminant;
  for (int i = 0; i < lineNo; ++i) {
    const CnLine li = lines.get(i);
    const float denominator = f2det(ln.dir, li.dir);
    const float numerator = f2det(li.dir, f2sub(ln.point, li.point));
    if (fabsf(denominator) <= CN_RVO_EPS) {
      if (numerator < 0.0f) return false;
      continue;
    }
    const float t = numerator / denominator;
    if (denominator >= 0.0f) {
      tRight = cn_minf(tRight, t);
    } else {
      tLeft = cn_maxf(tLeft, t);
    }
    if (tLeft > tRight) return false;
  }
  if (directionOpt) {
    if (f2dot(optVelocity, ln.dir) > 0.0f) {
      result = f2add(ln.point, f2scale(tRight, ln.dir));
    } else {
      result = f2add(ln.point, f2scale(tLeft, ln.dir));
    }
  } else {
    const float t = f2dot(ln.dir, f2sub(optVelocity, ln.point));
    if (t < tLeft) {
      result = f2add(ln.point, f2scale(tLeft, ln.dir));
    } else if (t > tRight) {
      result = f2add(ln.point, f2scale(tRight, ln.dir));
    } else {
      result = f2add(ln.point, f2scale(t, ln.dir));
    }
  }
  return true;
}

template <class Lines>
CN_HD int cn_lp2(const Lines& lines, int numLines, float radius, CnF2 optVelocity, bool directionOpt,
                 CnF2& result) {
  if (directionOpt) {
    result = f2scale(radius, optVelocity);   // optVelocity * radius (commutative)
  } else if (f2abssq(optVelocity) > radius * radius) {
    result = f2scale(radius, f2normalize(optVelocity));
  } else {
    result = optVelocity;
  }
  for (int i = 0; i < numLines; ++i) {
    const CnLine li = lines.get(i);
    if (f2det(li.dir, f2sub(li.point, result)) > 0.0f) {
      const CnF2 tempResult = result;
      if (!cn_lp1(lines, i, radius, optVelocity, directionOpt, result)) {
        result = tempResult;
        return i;
      }
    }
  }
  return numLines;
}

template <int MAXH>
struct CnLocalLines {
  float4 v[MAXH];
  CN_HD CnLine get(int k) const {
    CnLine l; l.point = f2(v[k].x, v[k].y); l.dir = f2(v[k].z, v[k].w); return l;
  }
  CN_HD void set(int k, const CnLine& l) {
    v[k].x = l.point.x; v[k].y = l.point.y; v[k].z = l.dir.x; v[k].w = l.dir.y;
  }
};

template <int MAXH, class Lines>
CN_HD_NOINLINE void cn_lp3(const Lines& lines, int numLines, int beginLine, float radius, CnF2& result) {
  CnLocalLines<MAXH> proj;
  float distance = 0.0f;
  for (int i = beginLine; i < numLines; ++i) {
    const CnLine li = lines.get(i);
    if (f2det(li.dir, f2sub(li.point, result)) > distance) {
      int np = 0;
      for (int j = 0; j < i; ++j) {
        const CnLine lj = lines.get(j);
        CnLine line;
        const float determinant = f2det(li.dir, lj.dir);
        if (fabsf(determinant) <= CN_RVO_EPS) {
          if (f2dot(li.dir, lj.dir) > 0.0f) continue;
          line.point = f2scale(0.5f, f2add(li.point, lj.point));
        } else {
          line.point = f2add(li.point,
                             f2scale(f2det(lj.dir, f2sub(li.point, lj.point)) / determinant, li.dir));
        }
        line.dir = f2normalize(f2sub(lj.dir, li.dir));
        proj.set(np++, line);
      }
      const CnF2 tempResult = result;
      if (cn_lp2(proj, np, radius, f2(-li.dir.y, li.dir.x), true, result) < np) {
        result = tempResult;
      }
      distance = f2det(li.dir, f2sub(li.point, result));
    }
  }
}

// One ORCA half-plane (Agent::computeNewVelocity body for one neighbour).
CN_HD CnLine cn_orca_line(CnF2 pos, CnF2 vel, float radius, CnF2 opos, CnF2 ovel, float oradius,
                          float invTimeHorizon, float timeStep) {
  const CnF2 relativePosition = f2sub(opos, pos);
  const CnF2 relativeVelocity = f2sub(vel, ovel);
  const float distSq = f2abssq(relativePosition);
  const float combinedRadius = radius + oradius;
  const float combinedRadiusSq = combinedRadius * combinedRadius;
  CnLine line;
  CnF2 u;
  if (distSq > combinedRadiusSq) {
    const CnF2 w = f2sub(relativeVelocity, f2scale(invTimeHorizon, relativePosition));
    const float wLengthSq = f2abssq(w);
    const float dotProduct1 = f2dot(w, relativePosition);
    if (dotProduct1 < 0.0f && dotProduct1 * dotProduct1 > combinedRadiusSq * wLengthSq) {
      const float wLength = sqrtf(wLengthSq);
      const CnF2 unitW = f2div(w, wLength);
      line.dir = f2(unitW.y, -unitW.x);
      u = f2scale(combinedRadius * invTimeHorizon - wLength, unitW);
    } else {
      const float leg = sqrtf(distSq - combinedRadiusSq);
      if (f2det(relativePosition, w) > 0.0f) {
        line.dir = f2div(f2(relativePosition.x * leg - relativePosition.y * combinedRadius,
                            relativePosition.x * combinedRadius + relativePosition.y * leg),
                         distSq);
      } else {
        line.dir = f2div(f2neg(f2(relativePosition.x * leg + relativePosition.y * combinedRadius,
                                  -relativePosition.x * combinedRadius + relativePosition.y * leg)),
                         distSq);
      }
      const float dotProduct2 = f2dot(relativeVelocity, line.dir);
      u = f2sub(f2scale(dotProduct2, line.dir), relativeVelocity);
    }
  } else {
    const float invTimeStep = 1.0f / timeStep;
    const CnF2 w = f2sub(relativeVelocity, f2scale(invTimeStep, relativePosition));
    const float wLength = f2abs(w);
    const CnF2 unitW = f2div(w, wLength);
    line.dir = f2(unitW.y, -unitW.x);
    u = f2scale(combinedRadius * invTimeStep - wLength, unitW);
  }
  line.point = f2add(vel, f2scale(0.5f, u));
  return line;
}


// ==========================================================================================
// Warp-cooperative LP.  In steady state ~30 % of the human-steps end in linearProgram3 and the
// serial per-thread solve leaves 1-2 lanes of a warp active for most of the kernel (profiles/).
// Here the 32 lanes of a warp solve ONE human's LP at a time: the loops over "previous lines"
// (linearProgram1's interval clipping, linearProgram3's projected-line construction) are split
// across lanes and reduced with exact min / max / ballot, so the result is bit-identical to the
// sequential RVO2 order:  tLeft only grows and tRight only shrinks, hence "some prefix has
// tLeft > tRight" <=> "the final interval is empty", and a parallel-infeasible line fails the LP
// wherever it sits.  Every function below is called by all lanes with warp-uniform arguments.

template <class Lines>
CN_HD bool cn_lp1_coop(const CnCoop& co, const Lines& lines, int lineNo, float radius, CnF2 optVelocity,
                       bool directionOpt, CnF2& result) {
  const CnLine ln = lines.get(lineNo);
  const float dotProduct = f2dot(ln.point, ln.dir);
  const float discriminant = dotProduct * dotProduct + radius * radius - f2abssq(ln.point);
  if (discriminant < 0.0f) return false;
  const float sqrtDiscriminant = sqrtf(discriminant);
  float tLeft = -dotProduct - sqrtDiscriminant;
  float tRight = -dotProduct + sqrtDiscriminant;
  float tl = -INFINITY, tr = INFINITY;
  bool bad = false;
  for (int i = co.lane; i < lineNo; i += co.nlanes) {
    const CnLine li = lines.get(i);
    const float denominator = f2det(ln.dir, li.dir);
    const float numerator = f2det(li.dir, f2sub(ln.point, li.point));
    if (fabsf(denominator) <= CN_RVO_EPS) {
      if (numerator < 0.0f) bad = true;
      continue;
    }
    const float t = numerator / denominator;
    if (denominator >= 0.0f) tr = cn_minf(tr, t);
    else tl = cn_maxf(tl, t);
  }
  bad = cn_any(co, bad);
  tr = cn_warp_min(co, tr);
  tl = cn_warp_max(co, tl);
  if (bad) return false;
  tRight = cn_minf(tRight, tr);
  tLeft = cn_maxf(tLeft, tl);
  if (tLeft > tRight) return false;
  if (directionOpt) {
    if (f2dot(optVelocity, ln.dir) > 0.0f) result = f2add(ln.point, f2scale(tRight, ln.dir));
    else result = f2add(ln.point, f2scale(tLeft, ln.dir));
  } else {
    const float t = f2dot(ln.dir, f2sub(optVelocity, ln.point));
    if (t < tLeft) result = f2add(ln.point, f2scale(tLeft, ln.dir));
    else if (t > tRight) result = f2add(ln.point, f2scale(tRight, ln.dir));
    else result = f2add(ln.point, f2scale(t, ln.dir));
  }
  return true;
}

// linearProgram2 with a warp-uniform result (used on the projected lines inside linearProgram3)
template <class Lines>
CN_HD int cn_lp2_coop(const CnCoop& co, const Lines& lines, int numLines, float radius, CnF2 optVelocity,
                      bool directionOpt, CnF2& result) {
  if (directionOpt) result = f2scale(radius, optVelocity);
  else if (f2abssq(optVelocity) > radius * radius) result = f2scale(radius, f2normalize(optVelocity));
  else result = optVelocity;
  for (int i = 0; i < numLines; ++i) {
    const CnLine li = lines.get(i);
    if (f2det(li.dir, f2sub(li.point, result)) > 0.0f) {
      const CnF2 tempResult = result;
      if (!cn_lp1_coop(co, lines, i, radius, optVelocity, directionOpt, result)) {
        result = tempResult;
        return i;
      }
    }
  }
  return numLines;
}

// linearProgram3: `proj` is a warp-shared scratch store for the projected lines
template <class Lines>
CN_HD void cn_lp3_coop(const CnCoop& co, const Lines& lines, int numLines, int beginLine, float radius, CnF2& result,
                       CnLineStore proj) {
  float distance = 0.0f;
  for (int i = beginLine; i < numLines; ++i) {
    const CnLine li = lines.get(i);
    if (f2det(li.dir, f2sub(li.point, result)) > distance) {
      int np = 0;
      for (int base = 0; base < i; base += co.nlanes) {
        const int j = base + co.lane;
        bool keep = false;
        CnLine line;
        line.point = f2(0.0f, 0.0f); line.dir = f2(0.0f, 0.0f);
        if (j < i) {
          const CnLine lj = lines.get(j);
          const float determinant = f2det(li.dir, lj.dir);
          if (fabsf(determinant) <= CN_RVO_EPS) {
            if (!(f2dot(li.dir, lj.dir) > 0.0f)) {
              line.point = f2scale(0.5f, f2add(li.point, lj.point));
              keep = true;
            }
          } else {
            line.point = f2add(li.point, f2scale(f2det(lj.dir, f2sub(li.point, lj.point)) / determinant, li.dir));
            keep = true;
          }
          if (keep) line.dir = f2normalize(f2sub(lj.dir, li.dir));
        }
        const uint32_t m = cn_ballot(co, keep);
        if (keep) proj.set(np + cn_popc_below(co, m), line);       // order-preserving compaction
        np += cn_popc(m);
      }
      cn_coop_sync(co);
      const CnF2 tempResult = result;
      if (cn_lp2_coop(co, proj, np, radius, f2(-li.dir.y, li.dir.x), true, result) < np) result = tempResult;
      distance = f2det(li.dir, f2sub(li.point, result));
      cn_coop_sync(co);
    }
  }
}

// The line stores of all lanes of a warp: lane L's lines start at smem0 + L (stride `stride`
// float4 per line) and its overflow lines (k >= cap) at ovf0 + L * ovf_stride.
struct CnWarpLines {
  float4* smem0;
  int stride, cap;
  float4* ovf0;
  int ovf_stride;
  CN_HD CnLineStore of(int lane) const {
    CnLineStore s; s.base = smem0 + lane; s.stride = stride; s.cap = cap; s.ovf = ovf0 + (size_t)lane * ovf_stride;
    return s;
  }
};

// linearProgram2 of all lanes of a warp (cooperative linearProgram1 servicing).  In: per-lane line
// count / max speed / preferred velocity (nl = 0 for idle lanes).  Out: per-lane velocity and LP2
// failure index (-1 = LP2 succeeded; otherwise linearProgram3 must follow from that line).
CN_HD void cn_orca_lp2_warp(const CnCoop& co, const CnWarpLines& W, int nl, float vmax, CnF2 pref, CnF2& result,
                            int& fail) {
  const CnLineStore mine = W.of(co.base + co.lane);
  // linearProgram2 initialisation (closest point, not direction)
  if (f2abssq(pref) > vmax * vmax) result = f2scale(vmax, f2normalize(pref));
  else result = pref;
  fail = -1;
  int i = 0;
  int running = nl > 0 ? 1 : 0;
  for (;;) {
    if (running) {           // own scan up to the next violated constraint (no collectives inside)
      while (i < nl) {
        const CnLine li = mine.get(i);
        if (f2det(li.dir, f2sub(li.point, result)) > 0.0f) break;
        ++i;
      }
      if (i == nl) running = 0;
    }
    cn_coop_sync(co);
    uint32_t req = cn_ballot(co, running != 0);
    if (!req) break;
    while (req) {
      const int L = cn_ffs(req);
      req &= req - 1;
      const int iL = cn_bcast_i(co, i, L);
      const float rL = cn_bcast_f(co, vmax, L);
      const CnF2 oL = f2(cn_bcast_f(co, pref.x, L), cn_bcast_f(co, pref.y, L));
      CnF2 r2 = f2(0.0f, 0.0f);
      const bool ok = cn_lp1_coop(co, W.of(co.base + L), iL, rL, oL, false, r2);
      if (co.lane == L) {
        if (ok) { result = r2; ++i; }
        else { fail = i; running = 0; }
      }
    }
  }
}

// linearProgram3 for the lanes whose LP2 failed, serviced within the warp one owner at a time
// (CPU harness / fallback; the step kernel balances these tasks across the whole CTA instead).
CN_HD void cn_orca_lp3_warp(const CnCoop& co, const CnWarpLines& W, int nl, float vmax, CnLineStore proj, CnF2& result,
                            int fail) {
  uint32_t req = cn_ballot(co, fail >= 0);
  while (req) {
    const int L = cn_ffs(req);
    req &= req - 1;
    const int nlL = cn_bcast_i(co, nl, L), fL = cn_bcast_i(co, fail, L);
    const float rL = cn_bcast_f(co, vmax, L);
    CnF2 resL = f2(cn_bcast_f(co, result.x, L), cn_bcast_f(co, result.y, L));
    cn_lp3_coop(co, W.of(co.base + L), nlL, fL, rL, resL, proj);
    if (co.lane == L) result = resL;
  }
}
