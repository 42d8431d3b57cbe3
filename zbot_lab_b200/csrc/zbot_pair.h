// zbot_pair.h -- F2: two float lanes that travel together (TWO ENVIRONMENTS per thread in the packed step kernel).
//
// sm_100a has packed FP32 arithmetic (FFMA2 / FMUL2 / FADD2: two independent IEEE fp32 operations per lane per
// instruction, with free scalar-broadcast and lane-swap operand forms).  Measured on B200 (tools/micro/ffma2_probe.cu):
// at 2 warps per SM sub-partition a stream of FFMA2 sustains 57.5 TFLOP/s where scalar FFMA sustains 39.2.  The fused
// step is FP32-issue bound at exactly that occupancy (register-limited), so the physics substep is instantiated with
// T = F2: every FFMA / FMUL / FADD of the articulated-body recursion serves two environments.
//
// The operators below use the round-to-nearest pair intrinsics; ptxas contracts a * b + c into one FFMA2 (checked in
// SASS).  On the host (CPU port, tests) F2 is two plain floats with the same per-lane semantics.
#pragma once
#ifndef ZB_PACKED_UNROLL
#define ZB_PACKED_UNROLL 1   // sweep unroll factor of the packed (two envs per thread) instantiation; tuning builds override
#endif
#include "zbot_core.h"

namespace zbot {

struct M2 { bool x, y; };   // per-lane predicate

struct F2 {
  float x, y;
  ZB_HD F2() {}
  ZB_HD F2(float a) : x(a), y(a) {}
  ZB_HD F2(double a) : x((float)a), y((float)a) {}
  ZB_HD F2(int a) : x((float)a), y((float)a) {}
  ZB_HD F2(float a, float b) : x(a), y(b) {}
};

#if defined(__CUDA_ARCH__)
#define ZB_F2_OP(name, expr_dev, expr_host) \
  { const float2 r_ = expr_dev; return F2(r_.x, r_.y); }
#else
#define ZB_F2_OP(name, expr_dev, expr_host) { return expr_host; }
#endif
#define ZB_F2V(a) make_float2((a).x, (a).y)

ZB_HD F2 operator+(F2 a, F2 b) ZB_F2_OP(add, __fadd2_rn(ZB_F2V(a), ZB_F2V(b)), F2(a.x + b.x, a.y + b.y))
ZB_HD F2 operator-(F2 a, F2 b) ZB_F2_OP(sub, __fadd2_rn(ZB_F2V(a), make_float2(-b.x, -b.y)), F2(a.x - b.x, a.y - b.y))
ZB_HD F2 operator*(F2 a, F2 b) ZB_F2_OP(mul, __fmul2_rn(ZB_F2V(a), ZB_F2V(b)), F2(a.x * b.x, a.y * b.y))
ZB_HD F2 operator*(float a, F2 b) ZB_F2_OP(muls, __fmul2_rn(make_float2(a, a), ZB_F2V(b)), F2(a * b.x, a * b.y))
ZB_HD F2 operator*(F2 a, float b) { return b * a; }
ZB_HD F2 operator+(float a, F2 b) { return F2(a) + b; }
ZB_HD F2 operator+(F2 a, float b) { return a + F2(b); }
ZB_HD F2 operator-(float a, F2 b) { return F2(a) - b; }
ZB_HD F2 operator-(F2 a, float b) { return a - F2(b); }
ZB_HD F2 operator-(F2 a) { return F2(-a.x, -a.y); }
ZB_HD F2& operator+=(F2& a, F2 b) { a = a + b; return a; }
ZB_HD F2& operator-=(F2& a, F2 b) { a = a - b; return a; }
ZB_HD F2& operator*=(F2& a, F2 b) { a = a * b; return a; }
ZB_HD F2& operator*=(F2& a, float b) { a = b * a; return a; }
ZB_HD F2& operator+=(F2& a, float b) { a = a + F2(b); return a; }

// per-lane (not packed in hardware): min / max / abs / special functions / compares / selects
ZB_HD F2 zb_min(F2 a, F2 b) { return F2(zb_min(a.x, b.x), zb_min(a.y, b.y)); }
ZB_HD F2 zb_max(F2 a, F2 b) { return F2(zb_max(a.x, b.x), zb_max(a.y, b.y)); }
ZB_HD F2 zb_abs(F2 a) { return F2(zb_abs(a.x), zb_abs(a.y)); }
ZB_HD F2 zb_sqrt(F2 a) { return F2(zb_sqrt(a.x), zb_sqrt(a.y)); }
ZB_HD F2 zb_rcp(F2 a) { return F2(zb_rcp(a.x), zb_rcp(a.y)); }
ZB_HD F2 zb_rsqrt(F2 a) { return F2(zb_rsqrt(a.x), zb_rsqrt(a.y)); }
ZB_HD void zb_sincos(F2 a, F2* s, F2* c) {
  zb_sincos(a.x, &s->x, &c->x);
  zb_sincos(a.y, &s->y, &c->y);
}
ZB_HD M2 zb_gt(F2 a, F2 b) { return M2{a.x > b.x, a.y > b.y}; }
ZB_HD bool zb_any(M2 m) { return m.x || m.y; }
ZB_HD F2 zb_sel(M2 m, F2 a, F2 b) { return F2(m.x ? a.x : b.x, m.y ? a.y : b.y); }

// ------------------------------------------------------------------------------------
// Phase B of the control step for TWO environments at once (cf. env_step_physics in zbot_core.h): the per-env
// bookkeeping (actuator targets, ContactSensor timers, force history) stays scalar per lane, the four physics
// substeps run once with T = F2.  Walking-v2 semantics (5-deep history with carry-over).
// ------------------------------------------------------------------------------------
ZB_HD F2 zb_pack(float a, float b) { return F2(a, b); }
ZB_HD float zb_lane(F2 v, int l) { return l ? v.y : v.x; }

template <typename Model, typename Scr2>
ZB_HD void env_step_physics2(const Params<float>& P, EnvState<float>& ea, EnvState<float>& eb, const float* raw_a,
                             const float* raw_b, PhysOut<float>& pa, PhysOut<float>& pb, Scr2& scr) {
  F2 target[6];
  {
    float na[6], nb[6], ta[6], tb[6];
    mdp_pre_physics<Model>(P, raw_a, ea.mdp, na, ta);
    mdp_pre_physics<Model>(P, raw_b, eb.mdp, nb, tb);
    ZB_UNROLL for (int k = 0; k < 6; ++k) target[k] = F2(ta[k], tb[k]);
  }
  pa.fz[4][0] = ea.carry_feet_fz[0]; pa.fz[4][1] = ea.carry_feet_fz[1]; pa.mid2 = ea.carry_mid_max * ea.carry_mid_max;
  pb.fz[4][0] = eb.carry_feet_fz[0]; pb.fz[4][1] = eb.carry_feet_fz[1]; pb.mid2 = eb.carry_mid_max * eb.carry_mid_max;
  SimState<F2> s;
  ZB_UNROLL for (int i = 0; i < 3; ++i) { s.p[i] = F2(ea.sim.p[i], eb.sim.p[i]); s.v[i] = F2(ea.sim.v[i], eb.sim.v[i]); s.w[i] = F2(ea.sim.w[i], eb.sim.w[i]); }
  ZB_UNROLL for (int i = 0; i < 4; ++i) s.Q[i] = F2(ea.sim.Q[i], eb.sim.Q[i]);
  ZB_UNROLL for (int i = 0; i < 6; ++i) { s.q[i] = F2(ea.sim.q[i], eb.sim.q[i]); s.qd[i] = F2(ea.sim.qd[i], eb.sim.qd[i]); }
  SubstepOut<F2> so;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int sub = 0; sub < P.decimation; ++sub) {
    physics_substep<Model, false, ZB_PACKED_UNROLL>(P, s, target, so, scr, (F2*)nullptr);   // plain sweep: the pipelined one needs 2 x 27 more registers
    const int slot = P.decimation - 1 - sub;   // newest first
    ZB_UNROLL for (int l = 0; l < 2; ++l) {
      EnvState<float>& e = l ? eb : ea;
      PhysOut<float>& po = l ? pb : pa;
      ZB_UNROLL for (int j = 0; j < 2; ++j) {
        const float fx = zb_lane(so.foot_force[j][0], l), fy = zb_lane(so.foot_force[j][1], l), fzz = zb_lane(so.foot_force[j][2], l);
        const float nrm = zb_sqrt(fx * fx + fy * fy + fzz * fzz);
        contact_timers_update(e.timers[j], nrm > 1.0f, P.dt);
        ZB_UNROLL for (int k = 0; k < 4; ++k) po.fz[k][j] = (slot == k) ? fzz : po.fz[k][j];
      }
      if (slot < 4) po.mid2 = zb_max(po.mid2, zb_lane(so.mid_force2_max, l));
    }
  }
  ZB_UNROLL for (int l = 0; l < 2; ++l) {
    EnvState<float>& e = l ? eb : ea;
    PhysOut<float>& po = l ? pb : pa;
    e.carry_feet_fz[0] = zb_lane(so.foot_force[0][2], l);
    e.carry_feet_fz[1] = zb_lane(so.foot_force[1][2], l);
    e.carry_mid_max = zb_sqrt(zb_lane(so.mid_force2_max, l));
    ZB_UNROLL for (int k = 0; k < 6; ++k) po.applied_torque[k] = zb_lane(so.applied_torque[k], l);
    ZB_UNROLL for (int i = 0; i < 3; ++i) { e.sim.p[i] = zb_lane(s.p[i], l); e.sim.v[i] = zb_lane(s.v[i], l); e.sim.w[i] = zb_lane(s.w[i], l); }
    ZB_UNROLL for (int i = 0; i < 4; ++i) e.sim.Q[i] = zb_lane(s.Q[i], l);
    ZB_UNROLL for (int i = 0; i < 6; ++i) { e.sim.q[i] = zb_lane(s.q[i], l); e.sim.qd[i] = zb_lane(s.qd[i], l); }
  }
}

#if defined(__CUDACC__)
struct SmemScratch2 {   // two lanes per slot: the thread's row of float2 words (odd stride in 8-byte units -> conflict-free)
  float2* base;
  struct Ref {
    float2* p;
    __device__ __forceinline__ operator F2() const { const float2 v = *p; return F2(v.x, v.y); }
    __device__ __forceinline__ Ref& operator=(F2 v) { *p = make_float2(v.x, v.y); return *this; }
  };
  __device__ __forceinline__ Ref operator()(int j, int slot) { return Ref{base + j * SCR_PER_JOINT + slot}; }
};
#endif

}  // namespace zbot
