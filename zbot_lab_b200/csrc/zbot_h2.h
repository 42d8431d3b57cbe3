// zbot_h2.h -- both HALVES of one environment in the two FP32 lanes of one thread.
//
// zbot_halves.h eliminates the walking robot's chain from both feet towards body 3.  The two sweeps are the same
// arithmetic on different data (3 joints, 3 bodies, the same contact candidates; only the feet's mass constants and a
// few signs differ), so one thread can run them TOGETHER in packed FP32 (F2 of zbot_pair.h: FFMA2 / FMUL2 / FADD2 on
// sm_100a): lane x = side A (foot_0, bodies 0..2, joints 0..2), lane y = side B (foot_1, bodies 6..4, joints 5..3).
// Against two ENVIRONMENTS per thread (zbot_step2_kernel, profiles/r1_notes.md: 255 registers + spills) the thread's
// state does not double -- the two lanes hold the two halves of the SAME robot -- and against two warps per env
// (zbot_w2_kernel.cuh) nothing is exchanged or computed twice.  What stays scalar: the kinematics sweep (sequential along
// the chain from the floating root), body 3 itself (its rigid terms belong to side A, its ground contact to side B), the
// 6x6 solve for body 3's acceleration and the root's integration.
//
// Per-joint scratch words are float2: word k holds (joint k, joint 5 - k), i.e. both lanes in ELIMINATION order, so
// iteration k of either packed sweep reads one 64-bit word per quantity.
//
// Same model, same discretisation, same state as physics_substep / physics_substep_halves: results agree to round-off
// (tests/test_cpu_oracles.py: host build against the halves form; tests/test_gpu_parity.py: kernel against the oracle).
#pragma once
#include "zbot_halves.h"
#include "zbot_pair.h"

namespace zbot {

// what a thread carries across the substeps of a control step: the floating root (foot_0) as scalars, the joints packed
struct H2State {
  float p[3], Q[4], v[3], w[3];
  F2 q[3], qd[3];            // word k = (joint k, joint 5 - k)
};

ZB_HD void h2_from_sim(const SimState<float>& s, H2State& h) {
  ZB_UNROLL for (int i = 0; i < 3; ++i) { h.p[i] = s.p[i]; h.v[i] = s.v[i]; h.w[i] = s.w[i]; }
  ZB_UNROLL for (int i = 0; i < 4; ++i) h.Q[i] = s.Q[i];
  ZB_UNROLL for (int k = 0; k < 3; ++k) { h.q[k] = F2(s.q[k], s.q[5 - k]); h.qd[k] = F2(s.qd[k], s.qd[5 - k]); }
}
ZB_HD void h2_to_sim(const H2State& h, SimState<float>& s) {
  ZB_UNROLL for (int i = 0; i < 3; ++i) { s.p[i] = h.p[i]; s.v[i] = h.v[i]; s.w[i] = h.w[i]; }
  ZB_UNROLL for (int i = 0; i < 4; ++i) s.Q[i] = h.Q[i];
  ZB_UNROLL for (int k = 0; k < 3; ++k) { s.q[k] = h.q[k].x; s.q[5 - k] = h.q[k].y; s.qd[k] = h.qd[k].x; s.qd[5 - k] = h.qd[k].y; }
}

// per-lane constants of elimination step k (outer body: k on side A, 6 - k on side B; joint: k / 5 - k)
ZB_HD F2 h2_joint_z(int k) { return F2((k == 0) ? float(model::JOINT_Z_FIRST) : float(model::JOINT_Z_REST), float(model::JOINT_Z_REST)); }
ZB_HD F2 h2_joint_sg(int k) {   // axis sign alternates along the chain: + for joints 0, 2, 4
  return (k & 1) ? F2(-float(model::AXIS_S), float(model::AXIS_S)) : F2(float(model::AXIS_S), -float(model::AXIS_S));
}
template <typename Model>
ZB_HD void h2_body(int k, F2& mass, F2& cx, F2& cz, F2& ixx, F2& iyy, F2& izz, F2& ixz) {
  float a[7], b[7];
  Model::template body<float>(k, a[0], a[1], a[2], a[3], a[4], a[5], a[6]);
  Model::template body<float>(6 - k, b[0], b[1], b[2], b[3], b[4], b[5], b[6]);
  mass = F2(a[0], b[0]); cx = F2(a[1], b[1]); cz = F2(a[2], b[2]); ixx = F2(a[3], b[3]); iyy = F2(a[4], b[4]);
  izz = F2(a[5], b[5]); ixz = F2(a[6], b[6]);
}
template <typename Model>
ZB_HD void h2_point(int k, int c, F2& lx, F2& ly, F2& lz, F2& drop) {
  float a[4], b[4];
  Model::template point<float>(k, c, a[0], a[1], a[2], a[3]);
  Model::template point<float>(6 - k, c, b[0], b[1], b[2], b[3]);
  lx = F2(a[0], b[0]); ly = F2(a[1], b[1]); lz = F2(a[2], b[2]); drop = F2(a[3], b[3]);
}

#ifndef ZB_H2_UNROLL_BWD
#define ZB_H2_UNROLL_BWD 1
#endif
#ifndef ZB_H2_UNROLL_PTS
#define ZB_H2_UNROLL_PTS 1
#endif

// the packed sweep's running state: pose / twist of the current (outer) body of each side about O, the articulated
// inertia and bias force accumulated so far
struct H2Sweep {
  F2 Q[4], r[3], w[3], v[3];
  SpInertia<F2> IA;
  F2 pAt[3], pAb[3];
  F2 mid2;
};

// One elimination step of both sides: rigid terms and ground contacts of the outer bodies (k | 6 - k), elimination of the
// joints (k | 5 - k), kinematics of the next bodies towards body 3.  kFoot: k == 0 (compile-time constants of the feet);
// otherwise k = 1, 2 at run time over the identical merged bodies.
template <typename Model, bool kFoot, typename PS, typename Scr2>
ZB_HD void h2_eliminate(const Params<PS>& P, int k, H2Sweep& S, Scr2& scr, F2 mu, F2 pz, ContactAgg<F2>& agg,
                        float* mid_force_out) {
  using namespace model;
  const int kc = kFoot ? 0 : 1;        // the body whose constants this instantiation uses
  F2 R[9];
  quat_to_mat(S.Q, R);
  {
    F2 mass, cx, cz, ixx, iyy, izz, ixz;
    h2_body<Model>(kc, mass, cx, cz, ixx, iyy, izz, ixz);
    body_rigid_terms(P, mass, cx, cz, ixx, iyy, izz, ixz, R, S.r, S.w, S.v, S.IA, S.pAt, S.pAb);
  }
  contact_agg_zero(agg);
  const int npts = kFoot ? 4 : 1;
#if defined(__CUDACC__)
ZB_PRAGMA_UNROLL(ZB_H2_UNROLL_PTS)
#endif
  for (int c = 0; c < npts; ++c) {
    F2 lx, ly, lz, drop;
    h2_point<Model>(kc, c, lx, ly, lz, drop);
    F2 rho[3] = {S.r[0] + R[0] * lx + R[1] * ly + R[2] * lz, S.r[1] + R[3] * lx + R[4] * ly + R[5] * lz,
                 S.r[2] + R[6] * lx + R[7] * ly + R[8] * lz - drop};
    contact_point(P, mu, rho, pz + rho[2], S.w, S.v, S.IA, S.pAt, S.pAb, &agg, (F2*)nullptr);
  }
  if (!kFoot) {
    S.mid2 = zb_max(S.mid2, agg.F0[0] * agg.F0[0] + agg.F0[1] * agg.F0[1] + agg.F0[2] * agg.F0[2]);
    if (mid_force_out) {   // lane x: body k (1, 2); lane y: body 6 - k (5, 4)
      ZB_UNROLL for (int i = 0; i < 3; ++i) { mid_force_out[3 * (k - 1) + i] = agg.F0[i].x; mid_force_out[3 * (5 - k) + i] = agg.F0[i].y; }
    }
  }
  // ---- eliminate the joint between this body (outer) and the next one towards body 3 ----
  const F2 Sa[3] = {scr(k, SC_SA), scr(k, SC_SA + 1), scr(k, SC_SA + 2)};
  const F2 Sm[3] = {scr(k, SC_SM), scr(k, SC_SM + 1), scr(k, SC_SM + 2)};
  const F2 qd = scr(k, SC_QD);
  const F2 sa[3] = {Sa[0] * qd, Sa[1] * qd, Sa[2] * qd};
  const F2 sm[3] = {Sm[0] * qd, Sm[1] * qd, Sm[2] * qd};
  F2 ct[3], cb[3], tmp[3];
  cross3(S.w, sa, ct);
  cross3(S.w, sm, cb);
  cross3(S.v, sa, tmp);
  cb[0] += tmp[0]; cb[1] += tmp[1]; cb[2] += tmp[2];
  F2 Ut[3], Ub[3];
  spi_mul(S.IA, Sa, Sm, Ut, Ub);
  const F2 D = dot3(Sa, Ut) + dot3(Sm, Ub) + F2(float(P.arm));
  const F2 Dinv = zb_rcp(D);
  const F2 u = F2(scr(k, SC_U)) - (dot3(Sa, S.pAt) + dot3(Sm, S.pAb));
  F2 Ict[3], Icb[3];
  spi_mul(S.IA, ct, cb, Ict, Icb);
  const F2 g = (u - (dot3(Ut, ct) + dot3(Ub, cb))) * Dinv;
  ZB_UNROLL for (int i = 0; i < 3; ++i) {
    S.pAt[i] += Ict[i] + Ut[i] * g;
    S.pAb[i] += Icb[i] + Ub[i] * g;
    scr(k, SC_UT + i) = Ut[i];
    scr(k, SC_UB + i) = Ub[i];
  }
  scr(k, SC_DINV) = Dinv;
  scr(k, SC_U) = u;
  spi_rank1_sub(S.IA, Ut, Ub, Dinv);
  // ---- kinematics of the next body: V_inner = V_outer - S' qd;  A: Q (x) qj, r + jz z_outer;  B: Q (x) conj(qj), r - jz z_inner ----
  ZB_UNROLL for (int i = 0; i < 3; ++i) { S.w[i] -= sa[i]; S.v[i] -= sm[i]; }
  const F2 dir = F2(1.f, -1.f);     // A applies the joint rotation (walks the chain forwards), B undoes it
  const F2 jz = h2_joint_z(kc);
  const F2 sg = kFoot ? h2_joint_sg(0) : h2_joint_sg(k);
  const F2 sn = scr(k, SC_SN), cs = scr(k, SC_CS);
  const float zo[3] = {R[2].x, R[5].x, R[8].x};          // lane x: third column of the OUTER body's rotation
  quat_mul_joint(S.Q, cs, dir * sg * sn, dir * (float(AXIS_S) * sn));
  const F2 zn[3] = {2.f * (S.Q[1] * S.Q[3] + S.Q[0] * S.Q[2]), 2.f * (S.Q[2] * S.Q[3] - S.Q[0] * S.Q[1]),
                    1.f - 2.f * (S.Q[1] * S.Q[1] + S.Q[2] * S.Q[2])};   // lane y: third column of the INNER body's rotation
  const F2 js = dir * jz;
  ZB_UNROLL for (int i = 0; i < 3; ++i) S.r[i] += js * F2(zo[i], zn[i].y);
}

// One physics substep.  `target3[k]` = (target of joint k, of joint 5 - k).  `mid_force_out`: [5][3] predictor contact
// forces of bodies 1..5, or null.  Scr2: float2 words, operator()(k, slot) convertible to / assignable from F2.
template <typename Model, typename PS, typename Scr2>
ZB_HD void physics_substep_h2(const Params<PS>& P, H2State& h, const F2* target3, SubstepOut<float>& out, Scr2& scr,
                              float* mid_force_out) {
  using namespace model;
  static_assert(!Model::kFullInertia && !Model::kPerEnvFriction && Model::kGroundForceSensor, "walking robot on flat ground");
  const float dt = float(P.dt);
  const F2 mu = F2(float(P.c_mu));
  // ---- PD (implicit part lives in P.arm), packed ----
  F2 applied[3];
  ZB_UNROLL for (int k = 0; k < 3; ++k) {
    const F2 e = target3[k] - h.q[k];
    applied[k] = zb_clamp(float(P.kp) * e - float(P.kd) * h.qd[k], F2(-float(P.effort)), F2(float(P.effort)));
    scr(k, SC_U) = zb_clamp(float(P.kp) * (e - dt * h.qd[k]) - float(P.kd) * h.qd[k], F2(-float(P.effort)), F2(float(P.effort)));
    scr(k, SC_QD) = h.qd[k];
  }
  // ---- kinematics sweep, scalar: root -> foot_1 through joints 0..5.  Parks S' = sigma (a; r x a) (sigma = -1 on side A:
  //      the outer body is the joint's parent) and the half-angle (sin, cos); side A's values wait in registers for their
  //      side-B partner (joint 5 - k) so that every scratch store is one 64-bit word. ----
  float Q[4] = {h.Q[0], h.Q[1], h.Q[2], h.Q[3]};
  float r[3] = {0.f, 0.f, 0.f};
  float w[3] = {h.w[0], h.w[1], h.w[2]};
  float vO[3] = {h.v[0], h.v[1], h.v[2]};
  {
    float hold[3][8];
    ZB_UNROLL for (int j = 0; j < 6; ++j) {
      float R[9];
      quat_to_mat(Q, R);
      const float jz = (j == 0) ? float(JOINT_Z_FIRST) : float(JOINT_Z_REST);
      const float sg = (j & 1) ? -float(AXIS_S) : float(AXIS_S);
      r[0] += jz * R[2]; r[1] += jz * R[5]; r[2] += jz * R[8];
      const float a[3] = {sg * R[0] + float(AXIS_S) * R[2], sg * R[3] + float(AXIS_S) * R[5], sg * R[6] + float(AXIS_S) * R[8]};
      float m[3];
      cross3(r, a, m);
      const int kk = (j < 3) ? j : 5 - j;
      const float qj = (j < 3) ? h.q[kk].x : h.q[kk].y;
      const float qd = (j < 3) ? h.qd[kk].x : h.qd[kk].y;
      float sn, cs;
      zb_sincos(0.5f * qj, &sn, &cs);
      ZB_UNROLL for (int i = 0; i < 3; ++i) { w[i] += a[i] * qd; vO[i] += m[i] * qd; }
      if (j < 3) {
        ZB_UNROLL for (int i = 0; i < 3; ++i) { hold[j][i] = -a[i]; hold[j][3 + i] = -m[i]; }
        hold[j][6] = sn; hold[j][7] = cs;
      } else {
        ZB_UNROLL for (int i = 0; i < 3; ++i) {
          scr(kk, SC_SA + i) = F2(hold[kk][i], a[i]);
          scr(kk, SC_SM + i) = F2(hold[kk][3 + i], m[i]);
        }
        scr(kk, SC_SN) = F2(hold[kk][6], sn);
        scr(kk, SC_CS) = F2(hold[kk][7], cs);
      }
      quat_mul_joint(Q, cs, sg * sn, float(AXIS_S) * sn);
    }
  }
  // ---- elimination from both feet towards body 3, packed: lane x walks bodies 0, 1, 2, lane y bodies 6, 5, 4 ----
  H2Sweep S;
  ZB_UNROLL for (int i = 0; i < 4; ++i) S.Q[i] = F2(h.Q[i], Q[i]);
  ZB_UNROLL for (int i = 0; i < 3; ++i) { S.r[i] = F2(0.f, r[i]); S.w[i] = F2(h.w[i], w[i]); S.v[i] = F2(h.v[i], vO[i]); }
  ZB_UNROLL for (int i = 0; i < 6; ++i) { S.IA.I[i] = F2(0.f); S.IA.M[i] = F2(0.f); }
  ZB_UNROLL for (int i = 0; i < 9; ++i) S.IA.H[i] = F2(0.f);
  ZB_UNROLL for (int i = 0; i < 3; ++i) { S.pAt[i] = F2(0.f); S.pAb[i] = F2(0.f); }
  S.mid2 = F2(0.f);
  ContactAgg<F2> agg_foot;
  const F2 pz = F2(h.p[2]);
  // step 0 (the feet: their own mass constants, four rim points) is peeled so that every model constant of either part is
  // a compile-time literal; steps 1, 2 (identical merged bodies, one sphere each) share one rolled body
  h2_eliminate<Model, true>(P, 0, S, scr, mu, pz, agg_foot, mid_force_out);
#if defined(__CUDACC__)
ZB_PRAGMA_UNROLL(ZB_H2_UNROLL_BWD)
#endif
  for (int k = 1; k < 3; ++k) {
    ContactAgg<F2> agg;
    h2_eliminate<Model, false>(P, k, S, scr, mu, pz, agg, mid_force_out);
  }
  F2 (&Q2)[4] = S.Q;
  F2 (&r2)[3] = S.r;
  F2 (&w2)[3] = S.w;
  F2 (&v2)[3] = S.v;
  const SpInertia<F2>& IA = S.IA;
  const F2 (&pAt)[3] = S.pAt;
  const F2 (&pAb)[3] = S.pAb;
  const F2 mid2 = S.mid2;
  // ---- body 3, scalar: side A adds its rigid terms, side B its ground contact; the two shares are summed and solved ----
  RootShare<float> rA, rB;
  ZB_UNROLL for (int i = 0; i < 6; ++i) { rA.IA.I[i] = IA.I[i].x; rB.IA.I[i] = IA.I[i].y; rA.IA.M[i] = IA.M[i].x; rB.IA.M[i] = IA.M[i].y; }
  ZB_UNROLL for (int i = 0; i < 9; ++i) { rA.IA.H[i] = IA.H[i].x; rB.IA.H[i] = IA.H[i].y; }
  ZB_UNROLL for (int i = 0; i < 3; ++i) { rA.pt[i] = pAt[i].x; rB.pt[i] = pAt[i].y; rA.pb[i] = pAb[i].x; rB.pb[i] = pAb[i].y; }
  {
    const float QA[4] = {Q2[0].x, Q2[1].x, Q2[2].x, Q2[3].x}, rr[3] = {r2[0].x, r2[1].x, r2[2].x};
    const float ww[3] = {w2[0].x, w2[1].x, w2[2].x}, vv[3] = {v2[0].x, v2[1].x, v2[2].x};
    float R[9];
    quat_to_mat(QA, R);
    model_body_terms<Model>(P, 3, R, rr, ww, vv, rA.IA, rA.pt, rA.pb);
  }
  float mid3;
  {
    const float QB[4] = {Q2[0].y, Q2[1].y, Q2[2].y, Q2[3].y}, rr[3] = {r2[0].y, r2[1].y, r2[2].y};
    const float ww[3] = {w2[0].y, w2[1].y, w2[2].y}, vv[3] = {v2[0].y, v2[1].y, v2[2].y};
    float R[9];
    quat_to_mat(QB, R);
    ContactAgg<float> agg;
    contact_agg_zero(agg);
    float lx, ly, lz, drop;
    Model::point(3, 0, lx, ly, lz, drop);
    float rho[3] = {rr[0] + R[0] * lx + R[1] * ly + R[2] * lz, rr[1] + R[3] * lx + R[4] * ly + R[5] * lz,
                    rr[2] + R[6] * lx + R[7] * ly + R[8] * lz - drop};
    contact_point(P, float(P.c_mu), rho, h.p[2] + rho[2], ww, vv, rB.IA, rB.pt, rB.pb, &agg, (float*)nullptr);
    mid3 = agg.F0[0] * agg.F0[0] + agg.F0[1] * agg.F0[1] + agg.F0[2] * agg.F0[2];
    if (mid_force_out) { ZB_UNROLL for (int i = 0; i < 3; ++i) mid_force_out[6 + i] = agg.F0[i]; }
  }
  float At[3], Ab[3];
  half_root_solve(rA, rB, At, Ab);
  // ---- outward sweep from body 3 to both feet, packed: joint accelerations, new joint velocities ----
  F2 At2[3] = {F2(At[0]), F2(At[1]), F2(At[2])}, Ab2[3] = {F2(Ab[0]), F2(Ab[1]), F2(Ab[2])};
  half_forward(P, 0, scr, w2, v2, At2, Ab2);       // side 0 indexing: joint word k at iteration k, for both lanes
  {
    F2 ff[3];
    contact_agg_force(agg_foot, F2(dt), At2, Ab2, ff);
    ZB_UNROLL for (int i = 0; i < 3; ++i) { out.foot_force[0][i] = ff[i].x; out.foot_force[1][i] = ff[i].y; }
  }
  out.mid_force2_max = zb_max(zb_max(mid2.x, mid2.y), mid3);
  ZB_UNROLL for (int k = 0; k < 3; ++k) { out.applied_torque[k] = applied[k].x; out.applied_torque[5 - k] = applied[k].y; }
  // ---- semi-implicit Euler: joints packed, the floating root (foot_0 = lane x of the outward sweep) scalar ----
  ZB_UNROLL for (int k = 0; k < 3; ++k) { h.qd[k] = scr(k, SC_QD); h.q[k] = zb_wrap_joint(h.q[k] + dt * h.qd[k]); }
  {
    float wxv[3];
    cross3(h.w, h.v, wxv);
    ZB_UNROLL for (int i = 0; i < 3; ++i) {
      h.w[i] += dt * At2[i].x;
      h.v[i] += dt * (Ab2[i].x + wxv[i]);
    }
    ZB_UNROLL for (int i = 0; i < 3; ++i) h.p[i] += dt * h.v[i];
    const float hh = 0.5f * dt;
    const float qw = h.Q[0], qx = h.Q[1], qy = h.Q[2], qz = h.Q[3];
    const float nw = qw + hh * (-h.w[0] * qx - h.w[1] * qy - h.w[2] * qz);
    const float nx = qx + hh * (h.w[0] * qw + h.w[1] * qz - h.w[2] * qy);
    const float ny = qy + hh * (-h.w[0] * qz + h.w[1] * qw + h.w[2] * qx);
    const float nz = qz + hh * (h.w[0] * qy - h.w[1] * qx + h.w[2] * qw);
    const float inv = zb_rsqrt(nw * nw + nx * nx + ny * ny + nz * nz);
    h.Q[0] = nw * inv; h.Q[1] = nx * inv; h.Q[2] = ny * inv; h.Q[3] = nz * inv;
  }
}

// host scratch (tests / CPU port)
struct H2ArrayScratch {
  F2 a[HALF_SCR_WORDS];
  ZB_HD F2& operator()(int j, int slot) { return a[j * SCR_PER_JOINT + slot]; }
};

// drop-in for physics_substep<Model> on a SimState (host build, tests)
template <typename Model, typename PS>
ZB_HD void physics_substep_h2_sim(const Params<PS>& P, SimState<float>& s, const float* target, SubstepOut<float>& out,
                                  float* mid_force_out) {
  H2State h;
  h2_from_sim(s, h);
  const F2 t3[3] = {F2(target[0], target[5]), F2(target[1], target[4]), F2(target[2], target[3])};
  H2ArrayScratch scr;
  physics_substep_h2<Model>(P, h, t3, out, scr, mid_force_out);
  h2_to_sim(h, s);
}

// Phase B of the control step with the packed substep (cf. env_step_physics in zbot_core.h; walking-v2 / v4 semantics)
template <typename Model, typename Scr2>
ZB_HD void env_step_physics_h2(const Params<float>& P, EnvState<float>& e, const float* raw_actions, PhysOut<float>& po, Scr2& scr,
                               StepExport<float>* ex) {
  float new_actions[6], target[7];
  mdp_pre_physics<Model>(P, raw_actions, e.mdp, new_actions, target);
  const F2 t3[3] = {F2(target[0], target[5]), F2(target[1], target[4]), F2(target[2], target[3])};
  po.fz[4][0] = (Model::kFresh) ? 0.f : e.carry_feet_fz[0];
  po.fz[4][1] = (Model::kFresh) ? 0.f : e.carry_feet_fz[1];
  po.mid2 = (Model::kFresh) ? 0.f : e.carry_mid_max * e.carry_mid_max;
  if (ex) {
    ZB_UNROLL for (int b = 0; b < 5; ++b) { ex->mid_force_hist[4][b][0] = (b == 0) ? e.carry_mid_max : 0.f;
      ex->mid_force_hist[4][b][1] = 0.f; ex->mid_force_hist[4][b][2] = 0.f; }
    ZB_UNROLL for (int j = 0; j < 2; ++j) { ex->feet_force_hist[4][j][0] = 0.f; ex->feet_force_hist[4][j][1] = 0.f;
      ex->feet_force_hist[4][j][2] = e.carry_feet_fz[j]; }
  }
  SubstepOut<float> so;
  if (Model::kFresh) po.mid2_h3 = 0.f;
  H2State h;
  h2_from_sim(e.sim, h);
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int sub = 0; sub < P.decimation; ++sub) {
    float midf[15];
    if (Model::kFresh) {
      if (sub == P.decimation - 1) { ZB_UNROLL for (int k = 0; k < 3; ++k) { po.qd_prev[k] = h.qd[k].x; po.qd_prev[5 - k] = h.qd[k].y; } }
    }
    physics_substep_h2<Model>(P, h, t3, so, scr, ex ? midf : (float*)nullptr);
    const int slot = P.decimation - 1 - sub;  // newest first
    ZB_UNROLL for (int j = 0; j < 2; ++j) {
      const float* ff = so.foot_force[j];
      const float nrm = zb_sqrt(ff[0] * ff[0] + ff[1] * ff[1] + ff[2] * ff[2]);
      contact_timers_update(e.timers[j], nrm > 1.0f, P.dt);
      ZB_UNROLL for (int k = 0; k < 4; ++k) po.fz[k][j] = (slot == k) ? ff[2] : po.fz[k][j];   // select, not index
      if (ex && slot < 4) { ex->feet_force_hist[slot][j][0] = ff[0]; ex->feet_force_hist[slot][j][1] = ff[1];
        ex->feet_force_hist[slot][j][2] = ff[2]; }
    }
    if (slot < 4) po.mid2 = zb_max(po.mid2, so.mid_force2_max);
    if (Model::kFresh) { if (slot < 3) po.mid2_h3 = zb_max(po.mid2_h3, so.mid_force2_max); }
    if (ex && slot < 4) {
      ZB_UNROLL for (int b = 0; b < 5; ++b)
        ZB_UNROLL for (int i = 0; i < 3; ++i) ex->mid_force_hist[slot][b][i] = midf[3 * b + i];
    }
  }
  h2_to_sim(h, e.sim);
  if (!Model::kFresh) {
    e.carry_feet_fz[0] = so.foot_force[0][2];
    e.carry_feet_fz[1] = so.foot_force[1][2];
    e.carry_mid_max = zb_sqrt(so.mid_force2_max);
  }
  ZB_UNROLL for (int k = 0; k < 6; ++k) po.applied_torque[k] = so.applied_torque[k];
}

}  // namespace zbot
