// zbot_halves.h -- the physics substep of the walking robot split into two HALVES per environment.
//
// Same model, same discretisation, same state as `physics_substep` (zbot_core.h): the floating root is still foot_0 and is
// integrated exactly as before.  What changes is the ORDER OF ELIMINATION of the articulated-body algorithm: the 7-body
// chain  foot_0 -j0- b1 -j1- b2 -j2- [body 3 = b3+base] -j3- b4 -j4- b5 -j5- foot_1  is eliminated from BOTH ends towards
// body 3,
//     side A: bodies 0, 1, 2 through joints 0, 1, 2  (outer body = the joint's parent: motion subspace S' = -S),
//     side B: bodies 6, 5, 4 through joints 5, 4, 3  (outer body = the joint's child:  S' = +S),
// the two partial articulated inertias / bias forces of body 3 (21 + 6 words each) are added, the 6x6 system is solved for
// the spatial acceleration of body 3, and each side sweeps back out to its foot.  Side A's sweep ends at body 0 with the
// root acceleration the old formulation solved for directly.  Everything is expressed in world-aligned Pluecker
// coordinates about the root origin O (as in zbot_core.h), so re-rooting is pure algebra: no frame changes.
//
// Why: on the GPU the two sides run in TWO WARPS of the same CTA (warp-specialised halves, named barriers + shared memory
// in between; zbot_kernels.cu: zbot_step_w2_kernel).  A warp carries half the per-env state and half the dependent chain
// of the one-thread-per-env kernel, so twice as many warps fit a scheduler at the same register file, and no instruction
// is issued twice (a lane-pair split would execute the serial kinematics chain redundantly).
// On a CPU (the host build used by the tests) the same functions run one after the other: `physics_substep_halves`.
//
// Derivation of the reversed joint (side A).  Joint k: V_{k+1} = V_k + S qd,  a_{k+1} = a_k + S qdd + V_{k+1} x S qd,
// tau = S^T f_{k+1} (f = force the joint transmits to the child).  Seen from the child: a_k = a_{k+1} + S' qdd + c' with
// S' = -S and c' = V_k x (S' qd)  (V_{k+1} x S qd = V_k x S qd since S x S = 0); the force on the parent is -f_{k+1}, so
// S'^T f_k = tau: the generalised force keeps its sign.  The leaf-elimination formulas are therefore the usual ones with
// S' in place of S:  U = IA S', D = S'^T U + armature, u = tau - S'^T pA, IA_parent += IA - U U^T / D, ...
#pragma once
#include "zbot_core.h"

namespace zbot {

constexpr int HALF_JOINTS = 3;
constexpr int HALF_SCR_WORDS = HALF_JOINTS * SCR_PER_JOINT;   // 51 words per thread (slot map: ScrSlot in zbot_core.h)

// What one side carries across the substeps of a control step.  Joint arrays are in CHAIN order: t = 0, 1, 2 is joint
// t (side A) or joint 3 + t (side B).  The root (foot_0) state is only meaningful on side A.
template <typename T>
struct HalfState {
  T p[3], Q[4], v[3], w[3];
  T q[3], qd[3];
};

// pose / twist of a body about O (world-aligned), as the kinematics sweep carries it
template <typename T>
struct BodyKin {
  T Q[4], r[3], w[3], vO[3];
};

// one side's share of body 3's articulated inertia and bias force
template <typename T>
struct RootShare {
  SpInertia<T> IA;
  T pt[3], pb[3];
};
constexpr int ROOT_SHARE_WORDS = 27;
constexpr int FRAME3_WORDS = 14;   // BodyKin (13) + the root height p_z

template <typename T>
struct HalfSubstepOut {
  T foot_force[3];       // net contact force on this side's foot (applied, world frame)
  T mid_force2_max;      // max |predictor contact force|^2 over this side's merged bodies (A: 1, 2; B: 5, 4, 3)
  T applied_torque[3];   // ImplicitActuator bookkeeping before this substep (chain order)
};

template <typename T>
ZB_HD T half_joint_z(int side, int t) { return (side == 0 && t == 0) ? T(model::JOINT_Z_FIRST) : T(model::JOINT_Z_REST); }
// joint axis (sg, 0, AXIS_S): sg alternates along the chain, + for joints 0, 2, 4
template <typename T>
ZB_HD T half_joint_sg(int side, int t) { return ((t + side) & 1) ? T(-model::AXIS_S) : T(model::AXIS_S); }

// ---- PD (implicit part lives in P.arm) ----
template <typename PS, typename T, typename Scr>
ZB_HD void half_pd(const Params<PS>& P, const HalfState<T>& h, const T* target3, Scr& scr, T* applied_torque3) {
  const T dt = T(P.dt);
  ZB_UNROLL for (int t = 0; t < 3; ++t) {
    const T e = target3[t] - h.q[t];
    applied_torque3[t] = zb_clamp(T(P.kp) * e - T(P.kd) * h.qd[t], T(-P.effort), T(P.effort));
    scr(t, SC_U) = zb_clamp(T(P.kp) * (e - dt * h.qd[t]) - T(P.kd) * h.qd[t], T(-P.effort), T(P.effort));
    scr(t, SC_QD) = h.qd[t];
  }
}

// ---- kinematics along this side's three joints, inner direction of the CHAIN (A: from the root up to body 3;
//      B: from body 3 down to foot_1).  `k` enters as the first body of the walk and leaves as the last one.
//      Parks S' = sigma (a; r x a) and (sin, cos) of the half angle per joint. ----
template <typename T, typename Scr>
ZB_HD void half_fk(int side, const HalfState<T>& h, BodyKin<T>& k, Scr& scr) {
  using namespace model;
  const T sigma = side ? T(1) : T(-1);
  ZB_UNROLL for (int t = 0; t < 3; ++t) {
    T R[9];
    quat_to_mat(k.Q, R);
    const T jz = half_joint_z<T>(side, t);
    const T sg = half_joint_sg<T>(side, t);
    k.r[0] += jz * R[2]; k.r[1] += jz * R[5]; k.r[2] += jz * R[8];
    const T a[3] = {sg * R[0] + T(AXIS_S) * R[2], sg * R[3] + T(AXIS_S) * R[5], sg * R[6] + T(AXIS_S) * R[8]};
    T m[3];
    cross3(k.r, a, m);
    const T qd = h.qd[t];
    T sn, cs;
    zb_sincos(T(0.5) * h.q[t], &sn, &cs);
    ZB_UNROLL for (int i = 0; i < 3; ++i) {
      scr(t, SC_SA + i) = sigma * a[i];
      scr(t, SC_SM + i) = sigma * m[i];
      k.w[i] += a[i] * qd;
      k.vO[i] += m[i] * qd;
    }
    scr(t, SC_SN) = sn;
    scr(t, SC_CS) = cs;
    quat_mul_joint(k.Q, cs, sg * sn, T(AXIS_S) * sn);
  }
}

// ---- elimination sweep from this side's foot (k = 0) to body 3 (k = 3).  `k0` = the foot's pose / twist (A: the root
//      state, B: what half_fk arrived at).  Leaves this side's share of body 3's articulated inertia / bias force in
//      `rs`, body 3's twist in (w3, vO3), the foot's contact aggregate in `agg`, and per joint U, 1/D, u in `scr`. ----
#ifndef ZB_HALF_UNROLL_PTS
#define ZB_HALF_UNROLL_PTS 1
#endif
template <typename Model, typename PS, typename T, typename Scr>
ZB_HD void half_backward(const Params<PS>& P, int side, T mu, T pz, const BodyKin<T>& k0, Scr& scr, RootShare<T>& rs,
                         T* w3, T* vO3, ContactAgg<T>& agg_foot, T& mid2, T* mid_force_out /* or null */,
                         const int* mid_off = nullptr /* where k = 1, 2, 3 go in mid_force_out (< 0: skip); default 0, 3, 6 */) {
  using namespace model;
  SpInertia<T>& IA = rs.IA;
  ZB_UNROLL for (int i = 0; i < 6; ++i) { IA.I[i] = T(0); IA.M[i] = T(0); }
  ZB_UNROLL for (int i = 0; i < 9; ++i) IA.H[i] = T(0);
  T* pAt = rs.pt;
  T* pAb = rs.pb;
  ZB_UNROLL for (int i = 0; i < 3; ++i) { pAt[i] = T(0); pAb[i] = T(0); }
  mid2 = T(0);
  contact_agg_zero(agg_foot);
  T Q[4] = {k0.Q[0], k0.Q[1], k0.Q[2], k0.Q[3]};
  T r[3] = {k0.r[0], k0.r[1], k0.r[2]};
  T w[3] = {k0.w[0], k0.w[1], k0.w[2]};
  T vO[3] = {k0.vO[0], k0.vO[1], k0.vO[2]};
  const T dir = side ? T(-1) : T(1);     // A walks the chain forwards (apply the joint rotation), B backwards (undo it)
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int k = 0; k < 4; ++k) {
    T R[9];
    quat_to_mat(Q, R);
    const int gb = side ? 6 - k : k;     // global body index (table lookups of the Model)
    if (k < 3 || side == 0) model_body_terms<Model>(P, gb, R, r, w, vO, IA, pAt, pAb);
    if (k < 3 || side == 1) {
      ContactAgg<T> agg;
      contact_agg_zero(agg);
      const int npts = (k == 0) ? 4 : 1;
#if defined(__CUDACC__)
ZB_PRAGMA_UNROLL(ZB_HALF_UNROLL_PTS)
#endif
      for (int c = 0; c < npts; ++c) {
        T lx, ly, lz, drop;
        Model::point(gb, c, lx, ly, lz, drop);
        T rho[3] = {r[0] + R[0] * lx + R[1] * ly + R[2] * lz, r[1] + R[3] * lx + R[4] * ly + R[5] * lz,
                    r[2] + R[6] * lx + R[7] * ly + R[8] * lz - drop};
        contact_point(P, mu, rho, pz + rho[2], w, vO, IA, pAt, pAb, &agg, (T*)nullptr);
      }
      if (k == 0) {
        agg_foot = agg;
      } else {
        mid2 = zb_max(mid2, agg.F0[0] * agg.F0[0] + agg.F0[1] * agg.F0[1] + agg.F0[2] * agg.F0[2]);
        if (mid_force_out) {
          const int off = mid_off ? mid_off[k - 1] : 3 * (k - 1);
          if (off >= 0) { ZB_UNROLL for (int i = 0; i < 3; ++i) mid_force_out[off + i] = agg.F0[i]; }
        }
      }
    } else if (mid_force_out && !mid_off) {
      ZB_UNROLL for (int i = 0; i < 3; ++i) mid_force_out[3 * (k - 1) + i] = T(0);
    }
    if (k == 3) break;
    // ---- eliminate the joint between this body (outer) and the next one towards body 3 ----
    const int t = side ? 2 - k : k;      // chain index of that joint
    const T Sa[3] = {scr(t, SC_SA), scr(t, SC_SA + 1), scr(t, SC_SA + 2)};
    const T Sm[3] = {scr(t, SC_SM), scr(t, SC_SM + 1), scr(t, SC_SM + 2)};
    const T qd = scr(t, SC_QD);
    const T sa[3] = {Sa[0] * qd, Sa[1] * qd, Sa[2] * qd};
    const T sm[3] = {Sm[0] * qd, Sm[1] * qd, Sm[2] * qd};
    T ct[3], cb[3], tmp[3];
    cross3(w, sa, ct);
    cross3(w, sm, cb);
    cross3(vO, sa, tmp);
    cb[0] += tmp[0]; cb[1] += tmp[1]; cb[2] += tmp[2];
    T Ut[3], Ub[3];
    spi_mul(IA, Sa, Sm, Ut, Ub);
    const T D = dot3(Sa, Ut) + dot3(Sm, Ub) + T(P.arm);
    const T Dinv = zb_rcp(D);
    const T u = scr(t, SC_U) - (dot3(Sa, pAt) + dot3(Sm, pAb));
    T Ict[3], Icb[3];
    spi_mul(IA, ct, cb, Ict, Icb);
    const T g = (u - (dot3(Ut, ct) + dot3(Ub, cb))) * Dinv;
    ZB_UNROLL for (int i = 0; i < 3; ++i) {
      pAt[i] += Ict[i] + Ut[i] * g;
      pAb[i] += Icb[i] + Ub[i] * g;
      scr(t, SC_UT + i) = Ut[i];
      scr(t, SC_UB + i) = Ub[i];
    }
    scr(t, SC_DINV) = Dinv;
    scr(t, SC_U) = u;
    spi_rank1_sub(IA, Ut, Ub, Dinv);
    // ---- kinematics of the next body: V_inner = V_outer - S' qd; A: Q (x) qj, r + jz z_outer; B: Q (x) conj(qj), r - jz z_inner ----
    ZB_UNROLL for (int i = 0; i < 3; ++i) { w[i] -= sa[i]; vO[i] -= sm[i]; }
    const T jz = half_joint_z<T>(side, t);
    const T sg = half_joint_sg<T>(side, t);
    const T sn = scr(t, SC_SN), cs = scr(t, SC_CS);
    if (side == 0) { r[0] += jz * R[2]; r[1] += jz * R[5]; r[2] += jz * R[8]; }
    quat_mul_joint(Q, cs, dir * sg * sn, dir * T(AXIS_S) * sn);
    if (side == 1) {
      const T z2[3] = {T(2) * (Q[1] * Q[3] + Q[0] * Q[2]), T(2) * (Q[2] * Q[3] - Q[0] * Q[1]),
                       T(1) - T(2) * (Q[1] * Q[1] + Q[2] * Q[2])};     // third column of the next body's rotation
      r[0] -= jz * z2[0]; r[1] -= jz * z2[1]; r[2] -= jz * z2[2];
    }
  }
  ZB_UNROLL for (int i = 0; i < 3; ++i) { w3[i] = w[i]; vO3[i] = vO[i]; }
}

// ---- body 3: add the other side's share (commutative: both sides get bit-identical sums), solve IA a3 = -pA ----
template <typename T>
ZB_HD void half_root_solve(const RootShare<T>& mine, const RootShare<T>& other, T* At, T* Ab) {
  SpInertia<T> IA;
  ZB_UNROLL for (int i = 0; i < 6; ++i) { IA.I[i] = mine.IA.I[i] + other.IA.I[i]; IA.M[i] = mine.IA.M[i] + other.IA.M[i]; }
  ZB_UNROLL for (int i = 0; i < 9; ++i) IA.H[i] = mine.IA.H[i] + other.IA.H[i];
  T nt[3], nb[3];
  ZB_UNROLL for (int i = 0; i < 3; ++i) { nt[i] = -(mine.pt[i] + other.pt[i]); nb[i] = -(mine.pb[i] + other.pb[i]); }
  spi_solve(IA, nt, nb, At, Ab);
}

// ---- outward sweep from body 3 (acceleration At, Ab; twist w3, vO3) to this side's foot: joint accelerations, new joint
//      velocities into scr(t, SC_QD); (At, Ab) leave as the foot's spatial acceleration ----
template <typename PS, typename T, typename Scr>
ZB_HD void half_forward(const Params<PS>& P, int side, Scr& scr, const T* w3, const T* vO3, T* At, T* Ab) {
  const T dt = T(P.dt);
  T wk[3] = {w3[0], w3[1], w3[2]}, vk[3] = {vO3[0], vO3[1], vO3[2]};
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int k = 2; k >= 0; --k) {
    const int t = side ? 2 - k : k;
    const T Sa[3] = {scr(t, SC_SA), scr(t, SC_SA + 1), scr(t, SC_SA + 2)};
    const T Sm[3] = {scr(t, SC_SM), scr(t, SC_SM + 1), scr(t, SC_SM + 2)};
    const T Ut[3] = {scr(t, SC_UT), scr(t, SC_UT + 1), scr(t, SC_UT + 2)};
    const T Ub[3] = {scr(t, SC_UB), scr(t, SC_UB + 1), scr(t, SC_UB + 2)};
    T qd = scr(t, SC_QD);
    const T sa[3] = {Sa[0] * qd, Sa[1] * qd, Sa[2] * qd};
    const T sm[3] = {Sm[0] * qd, Sm[1] * qd, Sm[2] * qd};
    T ct[3], cb[3], tmp[3];
    cross3(wk, sa, ct);
    cross3(wk, sm, cb);
    cross3(vk, sa, tmp);
    cb[0] += tmp[0]; cb[1] += tmp[1]; cb[2] += tmp[2];
    ZB_UNROLL for (int i = 0; i < 3; ++i) { At[i] += ct[i]; Ab[i] += cb[i]; wk[i] += sa[i]; vk[i] += sm[i]; }
    const T qdd = (scr(t, SC_U) - (dot3(Ut, At) + dot3(Ub, Ab))) * scr(t, SC_DINV);
    ZB_UNROLL for (int i = 0; i < 3; ++i) { At[i] += Sa[i] * qdd; Ab[i] += Sm[i] * qdd; }
    qd += dt * qdd;
    scr(t, SC_QD) = qd;
  }
}

// ---- semi-implicit Euler: joints of this side; side A also the floating root (At, Ab = spatial acceleration of foot_0) ----
template <typename PS, typename T, typename Scr>
ZB_HD void half_integrate(const Params<PS>& P, int side, HalfState<T>& h, Scr& scr, const T* At, const T* Ab) {
  const T dt = T(P.dt);
  ZB_UNROLL for (int t = 0; t < 3; ++t) { h.qd[t] = scr(t, SC_QD); h.q[t] = zb_wrap_joint(h.q[t] + dt * h.qd[t]); }   // walking robot: all six joints wrap
  if (side == 0) {
    T wxv[3];
    cross3(h.w, h.v, wxv);
    ZB_UNROLL for (int i = 0; i < 3; ++i) {
      h.w[i] += dt * At[i];
      h.v[i] += dt * (Ab[i] + wxv[i]);
    }
    ZB_UNROLL for (int i = 0; i < 3; ++i) h.p[i] += dt * h.v[i];
    const T hh = T(0.5) * dt;
    const T qw = h.Q[0], qx = h.Q[1], qy = h.Q[2], qz = h.Q[3];
    const T nw = qw + hh * (-h.w[0] * qx - h.w[1] * qy - h.w[2] * qz);
    const T nx = qx + hh * (h.w[0] * qw + h.w[1] * qz - h.w[2] * qy);
    const T ny = qy + hh * (-h.w[0] * qz + h.w[1] * qw + h.w[2] * qx);
    const T nz = qz + hh * (h.w[0] * qy - h.w[1] * qx + h.w[2] * qw);
    const T inv = zb_rsqrt(nw * nw + nx * nx + ny * ny + nz * nz);
    h.Q[0] = nw * inv; h.Q[1] = nx * inv; h.Q[2] = ny * inv; h.Q[3] = nz * inv;
  }
}

// ------------------------------------------------------------------------------------
// the same substep, both sides one after the other (host build, tests): drop-in for physics_substep<Model>
// ------------------------------------------------------------------------------------
template <typename T>
struct HalfArrayScratch {
  T a[HALF_SCR_WORDS];
  ZB_HD T& operator()(int j, int slot) { return a[j * SCR_PER_JOINT + slot]; }
};

template <typename Model, typename PS, typename T>
ZB_HD void physics_substep_halves(const Params<PS>& P, SimState<T>& s, const T* target, SubstepOut<T>& out,
                                  T* mid_force_out /* [5][3] bodies 1..5 or null */) {
  const T mu = Model::kPerEnvFriction ? target[6] : T(P.c_mu);
  HalfState<T> hA, hB;
  ZB_UNROLL for (int i = 0; i < 3; ++i) { hA.p[i] = s.p[i]; hA.v[i] = s.v[i]; hA.w[i] = s.w[i]; hB.p[i] = hB.v[i] = hB.w[i] = T(0); }
  ZB_UNROLL for (int i = 0; i < 4; ++i) { hA.Q[i] = s.Q[i]; hB.Q[i] = T(0); }
  ZB_UNROLL for (int t = 0; t < 3; ++t) { hA.q[t] = s.q[t]; hA.qd[t] = s.qd[t]; hB.q[t] = s.q[3 + t]; hB.qd[t] = s.qd[3 + t]; }
  HalfArrayScratch<T> scA, scB;
  HalfSubstepOut<T> oA, oB;
  half_pd(P, hA, target, scA, oA.applied_torque);
  half_pd(P, hB, target + 3, scB, oB.applied_torque);
  // kinematics: A from the root to body 3, then B from body 3 to foot_1
  BodyKin<T> root, kA, kB;
  ZB_UNROLL for (int i = 0; i < 3; ++i) { root.r[i] = T(0); root.w[i] = s.w[i]; root.vO[i] = s.v[i]; }
  ZB_UNROLL for (int i = 0; i < 4; ++i) root.Q[i] = s.Q[i];
  kA = root;
  half_fk(0, hA, kA, scA);
  kB = kA;                       // = the frame of body 3 side A posts to side B
  half_fk(1, hB, kB, scB);
  // elimination from both feet towards body 3
  RootShare<T> rA, rB;
  T w3a[3], v3a[3], w3b[3], v3b[3], mid2a, mid2b, midA[9], midB[9];
  ContactAgg<T> aggA, aggB;
  half_backward<Model>(P, 0, mu, s.p[2], root, scA, rA, w3a, v3a, aggA, mid2a, mid_force_out ? midA : (T*)nullptr);
  half_backward<Model>(P, 1, mu, s.p[2], kB, scB, rB, w3b, v3b, aggB, mid2b, mid_force_out ? midB : (T*)nullptr);
  T AtA[3], AbA[3], AtB[3], AbB[3];
  half_root_solve(rA, rB, AtA, AbA);
  half_root_solve(rB, rA, AtB, AbB);
  half_forward(P, 0, scA, w3a, v3a, AtA, AbA);
  half_forward(P, 1, scB, w3b, v3b, AtB, AbB);
  if (Model::kGroundForceSensor) {
    contact_agg_force(aggA, T(P.dt), AtA, AbA, out.foot_force[0]);
    contact_agg_force(aggB, T(P.dt), AtB, AbB, out.foot_force[1]);
  }
  out.mid_force2_max = zb_max(mid2a, mid2b);
  if (mid_force_out) {           // bodies 1, 2 from side A (k = 1, 2); bodies 3, 4, 5 from side B (k = 3, 2, 1)
    ZB_UNROLL for (int i = 0; i < 3; ++i) {
      mid_force_out[0 + i] = midA[0 + i]; mid_force_out[3 + i] = midA[3 + i];
      mid_force_out[6 + i] = midB[6 + i]; mid_force_out[9 + i] = midB[3 + i]; mid_force_out[12 + i] = midB[0 + i];
    }
  }
  ZB_UNROLL for (int t = 0; t < 3; ++t) { out.applied_torque[t] = oA.applied_torque[t]; out.applied_torque[3 + t] = oB.applied_torque[t]; }
  half_integrate(P, 0, hA, scA, AtA, AbA);
  half_integrate(P, 1, hB, scB, AtB, AbB);
  ZB_UNROLL for (int i = 0; i < 3; ++i) { s.p[i] = hA.p[i]; s.v[i] = hA.v[i]; s.w[i] = hA.w[i]; }
  ZB_UNROLL for (int i = 0; i < 4; ++i) s.Q[i] = hA.Q[i];
  ZB_UNROLL for (int t = 0; t < 3; ++t) { s.q[t] = hA.q[t]; s.qd[t] = hA.qd[t]; s.q[3 + t] = hB.q[t]; s.qd[3 + t] = hB.qd[t]; }
}

}  // namespace zbot
