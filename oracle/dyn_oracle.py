"""TEST INFRASTRUCTURE (oracle) -- float64 CPU restatement of one physics substep of
the 7-body floating-base ZBOT chain: implicit joint PD + linearly-implicit ground
contact, semi-implicit Euler.

PARITY UNPINNED vs the reference: the reference delegates this to PhysX 5 (closed,
absent here) and holds no trajectory fixture.  What IS pinned: the
kinematics (FK known answers printed at ``…env_v2.py:403-404`` and ``…env_v4.py:814-816``:
base height 0.2545, base quat (0.6003,-0.6003,-0.3735,-0.3739), foot_1 height 5.3035e-2)
and the physical constants (``assets/zbot_cfg.py:621-669``; ``zbot_6s_v04.usda:110-113,
192-195``).  Dynamics correctness is established by invariants (tests/test_dyn_oracle.py:
momentum/energy conservation, free fall, static stand).

DELIBERATELY a different formulation from the product kernel: the kernel runs an O(n)
articulated-body recursion in float32 with spatial (Pluecker) algebra about the root
origin; this oracle assembles the dense 12x12 joint-space mass matrix from classical
CoM Jacobians (Newton-Euler projected), adds the implicit PD / contact terms as dense
J^T K J blocks, and calls ``numpy.linalg.solve`` in float64.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import
this module.
"""
from __future__ import annotations

import numpy as np

from zbot_lab_b200.assets import zbot_6s as Z


def _cross(a, b):
    return np.cross(a, b)


def _quat_to_mat(q):
    w, x, y, z = (q[..., i] for i in range(4))
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - w * z)
    R[..., 0, 2] = 2 * (x * z + w * y)
    R[..., 1, 0] = 2 * (x * y + w * z)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - w * x)
    R[..., 2, 0] = 2 * (x * z - w * y)
    R[..., 2, 1] = 2 * (y * z + w * x)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _skew(v):
    S = np.zeros(v.shape[:-1] + (3, 3))
    S[..., 0, 1] = -v[..., 2]
    S[..., 0, 2] = v[..., 1]
    S[..., 1, 0] = v[..., 2]
    S[..., 1, 2] = -v[..., 0]
    S[..., 2, 0] = -v[..., 1]
    S[..., 2, 1] = v[..., 0]
    return S


class DynParams:
    def __init__(self, model=None, **kw):
        f32 = lambda x: float(np.float32(x))  # the C ABI carries parameters as float32 (ZbotCfg)
        model = model or Z.model_f32()
        self.dt = f32(Z.SIM_DT)
        self.kp, self.kd, self.effort = f32(model.kp), f32(model.kd), f32(model.effort_limit)
        self.gravity = model.gravity
        self.alpha, self.erp, self.vdep = f32(Z.CONTACT_ALPHA), f32(Z.CONTACT_ERP), f32(Z.CONTACT_VDEP)
        self.beta_max, self.mu, self.ramp = f32(Z.CONTACT_BETA_MAX), f32(Z.CONTACT_MU), f32(Z.CONTACT_RAMP)
        self.vt_eps = f32(Z.CONTACT_VT_EPS)
        self.margin = f32(Z.CONTACT_MARGIN)
        self.contacts = True
        self.pd = True
        self.__dict__.update(kw)


class DynOracle:
    """Batched (N envs) float64 state; positions are env-LOCAL."""

    def __init__(self, n, params: DynParams | None = None, model=None):
        self.n = n
        self.m = model or Z.model_f32()
        self.P = params or DynParams(self.m)
        self.root_pos = np.tile(self.m.default_root_pos, (n, 1))
        self.root_quat = np.tile(np.asarray(self.m.default_root_quat, float), (n, 1))
        self.root_lin_vel = np.zeros((n, 3))
        self.root_ang_vel = np.zeros((n, 3))
        self.q = np.tile(self.m.default_joint_pos, (n, 1))
        self.qd = np.zeros((n, 6))
        self.body_force = np.zeros((n, 7, 3))       # net contact force per reduced body (last substep)
        self.body_force_pred = np.zeros((n, 7, 3))  # predictor contact force (reported for bodies 1..5)
        self.applied_torque = np.zeros((n, 6))      # ImplicitActuator bookkeeping (before last substep)

    def set_state(self, st: dict, ids=None):
        sl = slice(None) if ids is None else ids
        self.root_pos[sl] = st["root_pos"]
        self.root_quat[sl] = st["root_quat"]
        self.root_lin_vel[sl] = st["root_lin_vel"]
        self.root_ang_vel[sl] = st["root_ang_vel"]
        self.q[sl] = st["joint_pos"]
        self.qd[sl] = st["joint_vel"]

    def reset(self, ids):
        self.root_pos[ids] = self.m.default_root_pos
        self.root_quat[ids] = np.asarray(self.m.default_root_quat, float)
        self.root_lin_vel[ids] = 0
        self.root_ang_vel[ids] = 0
        self.q[ids] = self.m.default_joint_pos
        self.qd[ids] = 0
        self.body_force[ids] = 0
        self.applied_torque[ids] = 0

    # ------------------------------------------------------------------ kinematics
    def kinematics(self):
        """Body frames + velocities.  child = parent o T(joint_pos) o Rot(axis, q)."""
        m, n = self.m, self.n
        bp = [self.root_pos]
        bq = [self.root_quat]
        for k in range(6):
            R = _quat_to_mat(bq[-1])
            bp.append(bp[-1] + np.einsum("nij,j->ni", R, m.joint_pos[k]))
            half = 0.5 * self.q[:, k:k + 1]
            qj = np.concatenate([np.cos(half), np.sin(half) * m.joint_axis[k]], -1)
            bq.append(Z.quat_mul(bq[-1], qj))
        bp = np.stack(bp, 1)          # (N,7,3)
        bq = np.stack(bq, 1)          # (N,7,4)
        R = _quat_to_mat(bq)          # (N,7,3,3)
        axis = np.stack([np.einsum("nij,j->ni", R[:, k + 1], m.joint_axis[k]) for k in range(6)], 1)  # (N,6,3)
        w = [self.root_ang_vel]
        v = [self.root_lin_vel]
        for k in range(6):
            v.append(v[-1] + _cross(w[-1], bp[:, k + 1] - bp[:, k]))
            w.append(w[-1] + axis[:, k] * self.qd[:, k:k + 1])
        return {"pos": bp, "quat": bq, "R": R, "axis": axis, "w": np.stack(w, 1), "v": np.stack(v, 1)}

    def link_state(self, kin=None):
        """Per-LINK pose / CoM velocity in articulation order (what ``robot.data`` exposes)."""
        kin = kin or self.kinematics()
        m = self.m
        b = m.link_body
        R = kin["R"][:, b]
        pos = kin["pos"][:, b] + np.einsum("nlij,lj->nli", R, m.link_offset)
        lquat = Z.quat_mul(kin["quat"][:, b], np.broadcast_to(m.link_rot, kin["quat"][:, b].shape))
        Rl = _quat_to_mat(lquat)
        com = pos + np.einsum("nlij,lj->nli", Rl, m.link_com)
        vcom = kin["v"][:, b] + _cross(kin["w"][:, b], com - kin["pos"][:, b])
        vlink = kin["v"][:, b] + _cross(kin["w"][:, b], pos - kin["pos"][:, b])
        out = {"body_link_pos": pos, "body_link_quat": lquat, "body_com_lin_vel": vcom,
               "body_link_lin_vel": vlink, "body_com_pos": com}
        if m.link_centre is not None:
            out["link_centre"] = kin["pos"][:, b] + np.einsum("nlij,lj->nli", R, m.link_centre)
        return out

    # ------------------------------------------------------------------ Jacobians
    def _point_jac(self, kin, body, r):
        """Jacobian (N,3,12) of the velocity of the material point at world position r on
        ``body`` w.r.t. nu = [v_root, w_root, qd]."""
        n = self.n
        J = np.zeros((n, 3, 12))
        J[:, :, 0:3] = np.eye(3)
        J[:, :, 3:6] = -_skew(r - kin["pos"][:, 0])
        for k in range(body):
            J[:, :, 6 + k] = _cross(kin["axis"][:, k], r - kin["pos"][:, k + 1])
        return J

    def _rot_jac(self, kin, body):
        J = np.zeros((self.n, 3, 12))
        J[:, :, 3:6] = np.eye(3)
        for k in range(body):
            J[:, :, 6 + k] = kin["axis"][:, k]
        return J

    def _vp_accels(self, kin):
        """Velocity-product (nu_dot = 0) classical accelerations: angular (N,7,3) and of body origins."""
        wd = [np.zeros((self.n, 3))]
        ap = [np.zeros((self.n, 3))]
        w, pos, axis = kin["w"], kin["pos"], kin["axis"]
        for k in range(6):
            d = pos[:, k + 1] - pos[:, k]
            ap.append(ap[-1] + _cross(wd[-1], d) + _cross(w[:, k], _cross(w[:, k], d)))
            wd.append(wd[-1] + _cross(w[:, k], axis[:, k]) * self.qd[:, k:k + 1])
        return np.stack(wd, 1), np.stack(ap, 1)

    def contact_points(self, kin):
        """List of (body, world position r) for every candidate contact point."""
        m = self.m
        pts = []
        for row in m.contact_list:
            b = int(row[0])
            c = kin["pos"][:, b] + np.einsum("nij,j->ni", kin["R"][:, b], row[1:4])
            pts.append((b, c - np.array([0.0, 0.0, row[4]])))
        return pts

    # ------------------------------------------------------------------ one substep
    def substep(self, q_target, ext_wrench=None):
        P, m, n = self.P, self.m, self.n
        dt = P.dt
        kin = self.kinematics()
        wd_vp, ap_vp = self._vp_accels(kin)
        nu = np.concatenate([self.root_lin_vel, self.root_ang_vel, self.qd], -1)

        M = np.zeros((n, 12, 12))
        rhs = np.zeros((n, 12))
        g = np.array([0.0, 0.0, -P.gravity])
        for i in range(7):
            R = kin["R"][:, i]
            c = kin["pos"][:, i] + np.einsum("nij,j->ni", R, m.body_com[i])
            Iw = np.einsum("nij,jk,nlk->nil", R, m.body_inertia[i], R)
            Jv = self._point_jac(kin, i, c)
            Jw = self._rot_jac(kin, i)
            M += m.body_mass[i] * np.einsum("nki,nkj->nij", Jv, Jv) + np.einsum("nki,nkl,nlj->nij", Jw, Iw, Jw)
            rho = c - kin["pos"][:, i]
            w = kin["w"][:, i]
            a_c = ap_vp[:, i] + _cross(wd_vp[:, i], rho) + _cross(w, _cross(w, rho))
            lin = m.body_mass[i] * (g - a_c)
            ang = -(np.einsum("nij,nj->ni", Iw, wd_vp[:, i]) + _cross(w, np.einsum("nij,nj->ni", Iw, w)))
            rhs += np.einsum("nki,nk->ni", Jv, lin) + np.einsum("nki,nk->ni", Jw, ang)

        # implicit PD (SURVEY §7 "Stiff implicit PD"; actuator cfg zbot_cfg.py:658-668)
        if P.pd:
            e = q_target - self.q
            self.applied_torque = np.clip(P.kp * e - P.kd * self.qd, -P.effort, P.effort)
            tau = np.clip(P.kp * (e - dt * self.qd) - P.kd * self.qd, -P.effort, P.effort)
            arm = dt * P.kd + dt * dt * P.kp
            rhs[:, 6:] += tau
            M[:, np.arange(6, 12), np.arange(6, 12)] += arm
        if ext_wrench is not None:
            rhs += ext_wrench

        # linearly-implicit ground contact
        clist = []
        pred = np.zeros((n, 7, 3))
        if P.contacts:
            k_n = P.alpha * P.erp / dt
            d_n = P.alpha * (1.0 - P.erp)
            for body, r in self.contact_points(kin):
                J = self._point_jac(kin, body, r)
                w = kin["w"][:, body]
                vc = np.einsum("nij,nj->ni", J, nu)
                vs = vc + dt * _cross(w, vc)
                rho = r - kin["pos"][:, body]
                a_vp = ap_vp[:, body] + _cross(wd_vp[:, body], rho) + _cross(w, _cross(w, rho))
                # classical accel of the material point = J nu_dot + a_vp; the predicted velocity
                # v + dt*a therefore carries dt*a_vp explicitly.  (w x v_c is the part of a_vp the
                # kernel's spatial formulation does not already include -- same total.)
                # ground height under the point: 0 on the plane; `self.ground(x, y)` (env-local) on a height field
                pen = -r[:, 2] if getattr(self, "ground", None) is None else self.ground(r[:, 0], r[:, 1]) - r[:, 2]
                s = np.clip(pen / P.ramp, 0.0, 1.0)
                fs = np.minimum(k_n * pen, P.alpha * P.vdep)
                gamma = k_n * dt + d_n * s
                fn0 = fs - gamma * vs[:, 2]
                # continuous activation: the implicit normal stiffness ramps in over fn0 in [-fband, 0]
                # (DESIGN.md §3); predictor force = max(fn0, 0)
                fband = 0.25 * k_n * P.ramp
                act = np.clip(fn0 / fband + 1.0, 0.0, 1.0) * (pen > -P.margin)
                fn0 = np.maximum(fn0, 0.0) * (pen > -P.margin)
                vt = np.sqrt(vs[:, 0] ** 2 + vs[:, 1] ** 2)
                beta = np.minimum(P.beta_max, P.mu * fn0 / np.maximum(vt, P.vt_eps))
                gamma = act * gamma
                K = np.stack([beta, beta, gamma], -1)                    # (N,3) diagonal
                # f = Fp - dt*K*(J nu_dot + a_vp - w x vc), Fp = (-beta v*_x, -beta v*_y, max(fn0, 0))
                #   = F0 - dt*K*J nu_dot   with F0 using the kernel's split (vs already has dt*w x vc)
                a_rest = a_vp - _cross(w, vc)
                Fp = np.stack([-beta * vs[:, 0], -beta * vs[:, 1], fn0], -1)
                F0 = Fp - dt * K * a_rest
                M += dt * np.einsum("nki,nk,nkj->nij", J, K, J)
                rhs += np.einsum("nki,nk->ni", J, F0)
                clist.append((body, J, K, F0))
                # predictor force (what the kernel reports for the mid bodies): F0 without the
                # velocity-product part
                pred[:, body] += Fp

        nud = np.linalg.solve(M, rhs[..., None])[..., 0]

        self.body_force_pred = pred
        self.body_force = np.zeros((n, 7, 3))
        for body, J, K, F0 in clist:
            self.body_force[:, body] += F0 - dt * K * np.einsum("nij,nj->ni", J, nud)

        nu = nu + dt * nud
        self.root_lin_vel, self.root_ang_vel, self.qd = nu[:, 0:3], nu[:, 3:6], nu[:, 6:]
        self.q = self.q + dt * self.qd
        # PhysX wraps revolute joints WITHOUT limits into [-2 pi, 2 pi] by shifting 4 pi (reference note
        # assets/test_articulation.py:18-20): all six joints of zbot_6s_new.usd, joints 1-5 of zbot_6s_v03.usd (joint6 has
        # +-720 deg limits), none of zbot_6s_v09.usd (+-360 deg limits)
        wrap = {"zbot_6s_new": np.ones(6, bool), "zbot_6s_v03": np.arange(6) != 5}.get(self.m.name, np.zeros(6, bool))
        hi, lo = (self.q > 2 * np.pi) & wrap, (self.q < -2 * np.pi) & wrap
        self.q = self.q - 4 * np.pi * hi + 4 * np.pi * lo
        self.root_pos = self.root_pos + dt * self.root_lin_vel
        wq = np.concatenate([np.zeros((n, 1)), self.root_ang_vel], -1)
        Q = self.root_quat + 0.5 * dt * Z.quat_mul(wq, self.root_quat)
        self.root_quat = Q / np.linalg.norm(Q, axis=-1, keepdims=True)
        return nud

    # ------------------------------------------------------------------ diagnostics
    def energy_momentum(self):
        """(kinetic+potential energy, linear momentum, angular momentum about the origin)."""
        kin = self.kinematics()
        m = self.m
        E = np.zeros(self.n)
        Pm = np.zeros((self.n, 3))
        L = np.zeros((self.n, 3))
        for i in range(7):
            R = kin["R"][:, i]
            c = kin["pos"][:, i] + np.einsum("nij,j->ni", R, m.body_com[i])
            vc = kin["v"][:, i] + _cross(kin["w"][:, i], c - kin["pos"][:, i])
            Iw = np.einsum("nij,jk,nlk->nil", R, m.body_inertia[i], R)
            w = kin["w"][:, i]
            Iww = np.einsum("nij,nj->ni", Iw, w)
            E += 0.5 * m.body_mass[i] * np.sum(vc * vc, -1) + 0.5 * np.sum(w * Iww, -1)
            E += m.body_mass[i] * self.P.gravity * c[:, 2]
            Pm += m.body_mass[i] * vc
            L += Iww + m.body_mass[i] * _cross(c, vc)
        return E, Pm, L
