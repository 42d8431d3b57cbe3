"""TEST INFRASTRUCTURE -- numpy float32 restatement of the MDP of ``zbot-6b-walking-m-v0`` (the manager-based task).

Only ``tests/`` and ``__graft_entry__.smoke()`` import this.  It restates, term by term,

* the reference's own functions: ``tasks/zbotlab_manager/mdp/rewards.py`` (foot_step_length :45-107, foot_downward :109-124,
  foot_forward :126-141, foot_clearance_reward :143-153, feet_gait :155-186, feet_air_time_positive_biped :211-230,
  air_time_variance_penalty :232-243, air_time_balance_penalty :245-252, feet_slide :254-269, base_vel_forward :271-281,
  feet_force_pattern :283-292, track_lin_vel_xy_yaw_frame_exp :294-305, track_ang_vel_z_world_exp :308-317,
  reset_my_data :37-43) and ``mdp/terminations.py`` (feet_close :55-60) -- PINNED: ``tests/golden/m_v0_*.npz`` holds the
  outputs of those functions, loaded unmodified (``oracle/ref_harness.RefMHarness``), on seeded synthetic data;
* Isaac Lab functions the cfg names but the reference does not vendor ([IL-upstream], version unpinned):
  ``joint_torques_l2``, ``joint_acc_l2``, ``action_rate_l2``, ``is_terminated``, ``time_out``,
  ``root_height_below_minimum``, ``UniformVelocityCommand`` (resample / compute), ``RelativeJointPositionAction``,
  ``reset_root_state_uniform``, ``reset_joints_by_scale``, and the ``ManagerBasedRLEnv.step`` order -- UNPINNED.

Inputs are the "view" the terms read at the end of physics (what ``zbot_m_step_export`` exports): see ``VIEW_COLS``.
"""
from __future__ import annotations

import numpy as np

F = np.float32

#: column map of the (N, 72) export row (MExport in csrc/zbot_core.h)
VIEW_COLS = {"root_pos": (0, 3), "root_quat": (3, 7), "root_lin_vel": (7, 10), "root_ang_vel": (10, 13),
             "feet_pos": (13, 19), "feet_quat": (19, 27), "feet_com_vel": (27, 33), "feet_fz_hist": (33, 39),
             "feet_fnorm_max": (39, 41), "last_air": (41, 43), "last_contact": (43, 45), "cur_air": (45, 47),
             "cur_contact": (47, 49), "q_chain": (49, 55), "tau": (55, 61), "joint_acc": (61, 67), "midb_max": (67, 72)}
_SHAPES = {"feet_pos": (2, 3), "feet_quat": (2, 4), "feet_com_vel": (2, 3), "feet_fz_hist": (3, 2)}


def split_view(rows: np.ndarray) -> dict:
    out = {}
    for k, (a, b) in VIEW_COLS.items():
        v = np.asarray(rows[:, a:b], F)
        out[k] = v.reshape((-1,) + _SHAPES[k]) if k in _SHAPES else v
    return out


def quat_apply(q, v):
    """isaaclab.utils.math.quat_apply (wxyz)."""
    xyz = q[..., 1:]
    t = (np.cross(xyz, v) * F(2)).astype(F)
    return (v + q[..., 0:1] * t + np.cross(xyz, t)).astype(F)


def yaw_cs(q):
    """cos / sin of the yaw of ``yaw_quat(q)`` [IL-upstream]: yaw = atan2(2 (w z + x y), 1 - 2 (y^2 + z^2))."""
    w, x, y, z = (q[..., i] for i in range(4))
    yaw = np.arctan2(F(2) * (w * z + x * y), F(1) - F(2) * (y * y + z * z)).astype(F)
    return np.cos(yaw).astype(F), np.sin(yaw).astype(F)


def forward_dir(root_quat, which=1.0):
    """cross(GRAVITY_VEC_W, quat_apply(root_quat_w, (0, which, 0))) -- rewards.py:62-64, 130-132, 274-276 (not normalised)."""
    y = np.zeros(root_quat.shape[:-1] + (3,), F)
    y[..., 1] = F(which)
    sh = quat_apply(root_quat, y)
    g = np.zeros_like(sh)
    g[..., 2] = F(-1)
    return np.cross(g, sh).astype(F)


class MTerms:
    """The RewTerm functions.  ``state`` = dict of the env attributes of rewards.py:29-35 (mutated like the reference)."""

    @staticmethod
    def track_lin_vel_xy_yaw_frame_exp(v, s, cmd, p):
        c, sn = yaw_cs(v["root_quat"])
        vx = c * v["root_lin_vel"][:, 0] + sn * v["root_lin_vel"][:, 1]
        vy = c * v["root_lin_vel"][:, 1] - sn * v["root_lin_vel"][:, 0]
        err = np.square(cmd[:, 0] - vx) + np.square(cmd[:, 1] - vy)
        return np.exp(-err / F(p["std"] ** 2)).astype(F)

    @staticmethod
    def track_ang_vel_z_world_exp(v, s, cmd, p):
        return np.exp(-np.square(cmd[:, 2] - v["root_ang_vel"][:, 2]) / F(p["std"] ** 2)).astype(F)

    @staticmethod
    def joint_torques_l2(v, s, cmd, p):          # [IL-upstream]
        return np.sum(np.square(v["tau"]), axis=1).astype(F)

    @staticmethod
    def joint_acc_l2(v, s, cmd, p):              # [IL-upstream]
        return np.sum(np.square(v["joint_acc"]), axis=1).astype(F)

    @staticmethod
    def action_rate_l2(v, s, cmd, p):            # [IL-upstream]: action - prev_action of the ActionManager (raw)
        return np.sum(np.square(s["action"] - s["prev_action"]), axis=1).astype(F)

    @staticmethod
    def foot_step_length(v, s, cmd, p):
        assert p.get("command_name") is None
        Fz = (((v["feet_fz_hist"][:, 0] + v["feet_fz_hist"][:, 1]) + v["feet_fz_hist"][:, 2]) / F(3)).astype(F)
        down = (Fz > F(10.0)) & (s["feet_contact_forces_last"] < F(10.0))
        fw = forward_dir(v["root_quat"])
        fw = (fw / (np.linalg.norm(fw, axis=1, keepdims=True).astype(F) + F(1e-6))).astype(F)
        vec = v["feet_pos"] - s["feet_down_pos_last"]
        ln = np.abs(np.sum(vec * fw[:, None, :], axis=-1)).astype(F)
        s["feet_step_length"] = np.where(down, ln, s["feet_step_length"]).astype(F)
        rew = np.min(s["feet_step_length"], axis=-1)
        s["feet_down_pos_last"] = np.where(down[..., None], v["feet_pos"], s["feet_down_pos_last"]).astype(F)
        s["feet_contact_forces_last"] = Fz.copy()
        return np.tanh(F(15.0) * rew).astype(F)

    @staticmethod
    def foot_downward(v, s, cmd, p):
        ax = np.array([[0, 1, 0], [0, -1, 0]], F)
        fz = quat_apply(v["feet_quat"], np.broadcast_to(ax, v["feet_pos"].shape))
        d = fz - np.array([0, 0, 1], F)
        return np.sum(np.linalg.norm(d, axis=-1), axis=-1).astype(F)

    @staticmethod
    def foot_forward(v, s, cmd, p):
        fw = forward_dir(v["root_quat"])
        fx = quat_apply(v["feet_quat"], np.broadcast_to(np.array([1, 0, 0], F), v["feet_pos"].shape))
        return np.sum(np.linalg.norm(fx - fw[:, None, :], axis=-1), axis=-1).astype(F)

    @staticmethod
    def feet_gait(v, s, cmd, p):
        period = F(p["period"])
        t = (s["episode_length_buf"].astype(F) * F(s["step_dt"])).astype(F)
        g = (np.mod(t, period) / period).astype(F)
        rew = np.zeros(len(g), F)
        for j, off in enumerate(p["offset"]):
            ph = np.mod(g + F(off), F(1.0)).astype(F)
            stance = ph < F(p.get("threshold", 0.5))
            rew += (~(stance ^ (v["cur_contact"][:, j] > 0))).astype(F)
        if p.get("command_name") is not None:
            rew = rew * (np.linalg.norm(cmd, axis=1) > F(0.05))
        return rew.astype(F)

    @staticmethod
    def feet_slide(v, s, cmd, p):
        contacts = v["feet_fnorm_max"] > F(1.0)
        sp = np.linalg.norm(v["feet_com_vel"][:, :, :2], axis=-1).astype(F)
        return np.sum(sp * contacts, axis=1).astype(F)

    @staticmethod
    def foot_clearance_reward(v, s, cmd, p):
        err = np.square(v["feet_pos"][:, :, 2] - F(p["target_height"]))
        vt = np.tanh(F(p["tanh_mult"]) * np.linalg.norm(v["feet_com_vel"][:, :, :2], axis=2).astype(F))
        return np.exp(-np.sum(err * vt, axis=1) / F(p["std"])).astype(F)

    @staticmethod
    def feet_air_time_positive_biped(v, s, cmd, p):
        inc = v["cur_contact"] > 0
        mode = np.where(inc, v["cur_contact"], v["cur_air"])
        single = inc.astype(np.int32).sum(1) == 1
        rew = np.min(np.where(single[:, None], mode, F(0)), axis=1)
        rew = np.minimum(rew, F(p["threshold"]))
        return (rew * (np.linalg.norm(cmd[:, :2], axis=1) > F(0.1))).astype(F)

    @staticmethod
    def air_time_variance_penalty(v, s, cmd, p):
        a = np.minimum(v["last_air"], F(0.5))
        c = np.minimum(v["last_contact"], F(0.5))
        return (np.var(a, axis=1, ddof=1) + np.var(c, axis=1, ddof=1)).astype(F)

    @staticmethod
    def air_time_balance_penalty(v, s, cmd, p):
        return np.abs(v["last_air"][:, 0] - v["last_air"][:, 1]).astype(F)

    @staticmethod
    def base_vel_forward(v, s, cmd, p):
        fw = forward_dir(v["root_quat"], p.get("which_forward", 1))
        return np.sum(v["root_lin_vel"] * fw, axis=-1).astype(F)

    @staticmethod
    def feet_force_pattern(v, s, cmd, p):
        Fz = (((v["feet_fz_hist"][:, 0] + v["feet_fz_hist"][:, 1]) + v["feet_fz_hist"][:, 2]) / F(3)).astype(F)
        diff = (Fz[:, 1] - Fz[:, 0]) * np.sign(s["feet_force_sum"])
        s["feet_force_sum"] = (s["feet_force_sum"] + F(0.001) * (Fz[:, 0] - Fz[:, 1])).astype(F)
        return (F(0.5) * diff - F(0.1) * np.abs(s["feet_force_sum"])).astype(F)

    @staticmethod
    def undesired_contacts(v, s, cmd, p):
        """isaaclab.envs.mdp.undesired_contacts [IL-upstream]: number of selected sensor bodies whose force norm exceeded the
        threshold anywhere in the history (zbotlab_env_cfg.py:367-371: base|a.*|b.* -> here the five merged bodies)."""
        return (v["midb_max"] > F(p.get("threshold", 1.0))).sum(1).astype(F)


class MMdpOracle:
    """Manager-ordered control step on a supplied view.  ``terms`` = [(name, func_name, weight, params)] in cfg order
    (``is_terminated`` included); ``P`` = dict(minimum_height, feet_close_min, cmd_ranges, rel_standing_envs,
    resampling_time_range, pose_range, max_episode_length, step_dt)."""

    def __init__(self, n, terms, P, model=None):
        from zbot_lab_b200.assets import zbot_6s_v2 as V
        self.n, self.terms, self.P = n, list(terms), dict(P)
        self.V = V
        self.m = model or V.model_f32()
        z = lambda *s: np.zeros(s, F)
        self.s = {"feet_force_sum": z(n), "feet_step_length": z(n, 2), "feet_contact_forces_last": z(n, 2),
                  "feet_down_pos_last": z(n, 2, 3), "action": z(n, 6), "prev_action": z(n, 6),
                  "episode_length_buf": np.zeros(n, np.int64), "step_dt": P["step_dt"]}
        self.cmd = z(n, 3)
        self.standing = np.zeros(n, bool)
        self.time_left = z(n)
        self.heading_target, self.is_heading = z(n), np.zeros(n, bool)      # heading_command=True [IL-upstream]
        self.push_left = z(n)                                               # EventTerm push_robot interval timer [IL-upstream]
        self.push_dv = z(n, 2)                                              # velocity the last step's push added (0 if none)
        # rough terrain (P["terrain"] = {"origins": (rows, cols, 3), "tile_size", "curriculum"}): TerrainImporter state [IL-upstream]
        self.levels, self.types = np.zeros(n, np.int64), np.zeros(n, np.int64)
        self.env_origins = z(n, 3)
        self.ep_sums = {name: z(n) for name, f, w, p in self.terms if float(w) != 0.0}

    # -- UniformVelocityCommand._resample [IL-upstream]
    def _resample(self, ids, u):
        P = self.P
        lo, hi = P["resampling_time_range"]
        self.time_left[ids] = u[:, 0] * F(hi - lo) + F(lo)
        for i in range(3):
            a, b = P["cmd_ranges"][i]
            self.cmd[ids, i] = u[:, 1 + i] * (F(b) - F(a)) + F(a)
        self.standing[ids] = u[:, 4] <= F(P["rel_standing_envs"])

    def _resample_heading(self, ids, u):
        h = self.P["heading"]
        lo, hi = h["range"]
        self.heading_target[ids] = u[:, 0] * (F(hi) - F(lo)) + F(lo)
        self.is_heading[ids] = u[:, 1] <= F(h["rel_heading_envs"])

    def step(self, view: dict, raw_actions: np.ndarray, rnd: np.ndarray):
        n, P, s = self.n, self.P, self.s
        a = np.asarray(raw_actions, F)
        s["prev_action"], s["action"] = s["action"], a.copy()          # ActionManager.process_action
        s["episode_length_buf"] = s["episode_length_buf"] + 1
        # TerminationManager
        time_out = s["episode_length_buf"] >= P["max_episode_length"]
        low = (view["root_pos"][:, 2] + self.env_origins[:, 2]).astype(F) < F(P["minimum_height"])     # root_pos_w: world height
        close = np.zeros(n, bool)
        if P.get("feet_close_min"):
            close = np.linalg.norm(view["feet_pos"][:, 0] - view["feet_pos"][:, 1], axis=-1).astype(F) < F(P["feet_close_min"])
        illegal = np.zeros(n, bool)
        if P.get("illegal_contact"):                               # mdp.illegal_contact: any selected body with max_t |F| > threshold
            thr, mask = P["illegal_contact"]
            sel = np.array([(mask >> b) & 1 for b in range(5)], bool)
            illegal = (view["midb_max"][:, sel] > F(thr)).any(1)
        terminated = low | close | illegal
        # RewardManager
        reward = np.zeros(n, F)
        values = {}
        dt = F(P["step_dt"])
        for name, func, w, p in self.terms:
            if float(w) == 0.0:
                continue
            val = terminated.astype(F) if func == "is_terminated" else getattr(MTerms, func)(view, s, self.cmd, p)
            values[name] = val
            r = (val * F(w) * dt).astype(F)
            reward = (reward + r).astype(F)
            self.ep_sums[name] = (self.ep_sums[name] + r).astype(F)
        reset = terminated | time_out
        ids = np.nonzero(reset)[0]
        log = None
        new_root = None
        if len(ids):
            log = {name: float(np.mean(v[ids]) / F(P["max_episode_length"] * P["step_dt"])) for name, v in self.ep_sums.items()}
            log["#base_height"], log["#feet_close"], log["#time_out"] = int(low[ids].sum()), int(close[ids].sum()), int(time_out[ids].sum())
            log["#illegal_contact"] = int(illegal[ids].sum())
            if P.get("terrain") and P["terrain"].get("curriculum"):
                # CurriculumManager.compute(env_ids) runs first in _reset_idx: terrain_levels_vel (mdp/curriculums.py:26-55) on
                # the pre-reset root position / command, then TerrainImporter.update_env_origins [IL-upstream]
                from zbot_lab_b200.terrain import terrain_levels_vel
                tp = P["terrain"]
                rows = tp["origins"].shape[0]
                up, down = terrain_levels_vel(view["root_pos"][ids, :2], self.cmd[ids, :2], tp["tile_size"],
                                              P["max_episode_length"] * P["step_dt"])
                lv = self.levels[ids] + up.astype(np.int64) - down.astype(np.int64)
                rand_lv = np.minimum((np.asarray(rnd, F)[ids, 21] * F(rows)).astype(np.int64), rows - 1)
                self.levels[ids] = np.where(lv >= rows, rand_lv, np.clip(lv, 0, None))
                self.env_origins[ids] = tp["origins"][self.levels[ids], self.types[ids]]
                log["#move_up"], log["#move_down"] = int(up.sum()), int(down.sum())
            # reset_root_state_uniform on the root link + reset_joints_by_scale (1,1) + reset_my_data
            u = np.asarray(rnd, F)[ids]
            pr = P["pose_range"]
            smp = [u[:, i] * (F(pr[i][1]) - F(pr[i][0])) + F(pr[i][0]) for i in range(3)]
            base_pos = np.tile(np.asarray(self.V.DEFAULT_ROOT_POS, F), (len(ids), 1))
            base_pos[:, 0] += smp[0]
            base_pos[:, 1] += smp[1]
            yaw = smp[2]
            new_root = {"ids": ids, "base_pos": base_pos, "yaw": yaw}
            fp, rq = self.post_reset_feet_and_root_quat(base_pos, yaw)
            s["feet_force_sum"][ids] = 0
            s["feet_step_length"][ids] = 0
            s["feet_contact_forces_last"][ids] = 0
            s["feet_down_pos_last"][ids] = fp
            new_root["root_quat"] = rq
            s["action"][ids] = 0
            s["prev_action"][ids] = 0
            for v in self.ep_sums.values():
                v[ids] = 0
            s["episode_length_buf"][ids] = 0
            self._resample(ids, u[:, 3:8])
            if P.get("heading"):
                self._resample_heading(ids, u[:, 13:15])
            if P.get("push"):                                          # EventManager.reset re-draws the interval timer
                lo, hi = P["push"]["interval_range_s"]
                self.push_left[ids] = u[:, 17] * (F(hi) - F(lo)) + F(lo)
        # CommandManager.compute
        self.time_left = (self.time_left - dt).astype(F)
        rs = np.nonzero(self.time_left <= 0)[0]
        if len(rs):
            self._resample(rs, np.asarray(rnd, F)[rs][:, 8:13])
            if P.get("heading"):
                self._resample_heading(rs, np.asarray(rnd, F)[rs][:, 15:17])
        if P.get("heading"):
            # _update_command: ang_vel_z = clip(stiffness * wrap_to_pi(target - heading_w), ang_vel_z range) for the heading envs;
            # heading_w = atan2 of the root x axis (post-reset root for the envs reset above)
            rq_h = view["root_quat"].copy()
            if len(ids):
                rq_h[ids] = new_root["root_quat"]
            ex_ = np.zeros((n, 3), F)
            ex_[:, 0] = 1
            fw = quat_apply(rq_h, ex_)
            hw = np.arctan2(fw[:, 1], fw[:, 0]).astype(F)
            d = (self.heading_target - hw).astype(F)
            err = np.arctan2(np.sin(d), np.cos(d)).astype(F)
            a_, b_ = P["cmd_ranges"][2]
            wz = np.clip(F(P["heading"]["stiffness"]) * err, F(a_), F(b_)).astype(F)
            self.cmd[self.is_heading, 2] = wz[self.is_heading]
        self.cmd[self.standing] = 0
        # EventManager.apply(mode="interval"): push_by_setting_velocity
        self.push_dv[:] = 0
        if P.get("push"):
            self.push_left = (self.push_left - dt).astype(F)
            ps = np.nonzero(self.push_left < F(1e-6))[0]
            if len(ps):
                up = np.asarray(rnd, F)[ps]
                lo, hi = P["push"]["interval_range_s"]
                self.push_left[ps] = up[:, 18] * (F(hi) - F(lo)) + F(lo)
                for i, k in enumerate(("x", "y")):
                    a_, b_ = P["push"]["velocity_range"].get(k, (0.0, 0.0))
                    self.push_dv[ps, i] = up[:, 19 + i] * (F(b_) - F(a_)) + F(a_)
        # observation (clean): root_quat, command, joint_pos_rel, joint_vel_rel (Isaac Lab joint order), last_action
        il = np.asarray(self.V.CHAIN_TO_IL)
        q_rel = np.zeros((n, 6), F)
        q_rel[:, il] = view["q_chain"] - np.asarray(self.m.default_joint_pos, F)
        qd = np.zeros((n, 6), F)
        if "qd_chain" in view:
            qd[:, il] = view["qd_chain"]
        rq = view["root_quat"].copy()
        if len(ids):
            q_rel[ids] = 0
            qd[ids] = 0
            rq[ids] = new_root["root_quat"]
        obs = np.concatenate([rq, self.cmd, q_rel, qd, s["action"]], axis=1).astype(F)
        return {"obs": obs, "reward": reward, "terminated": terminated, "time_outs": time_out, "reset_ids": ids,
                "values": values, "log": log, "resample_ids": rs, "new_root": new_root,
                "low": low, "close": close, "illegal": illegal}

    def post_reset_feet_and_root_quat(self, base_pos, yaw):
        """Feet LINK positions and root quaternion right after reset_base: the default pose moved by (x, y, yaw)."""
        from zbot_lab_b200.assets import zbot_6s as Z
        lp, lq = self.V.default_link_poses()
        i0, i1, ib = (self.V.link_index(k) for k in ("foot0", "foot1", "base"))
        rel = np.stack([lp[i0] - lp[ib], lp[i1] - lp[ib]])                      # feet relative to the base link, default pose
        qz = np.stack([np.cos(yaw / 2), 0 * yaw, 0 * yaw, np.sin(yaw / 2)], -1).astype(np.float64)
        fp = base_pos[:, None, :].astype(np.float64) + Z.quat_rotate(qz[:, None, :], np.broadcast_to(rel, (len(yaw), 2, 3)))
        rq = Z.quat_mul(qz, np.broadcast_to(lq[ib], qz.shape))
        return fp.astype(F), rq.astype(F)


def synth_m_views(seed: int, n: int, steps: int):
    """Seeded synthetic end-of-physics views + raw actions + uniforms (golden cases and CPU tests)."""
    rng = np.random.default_rng(seed)

    def rq(shape, tilt):
        ax = rng.normal(0, 1, shape + (3,))
        ax /= np.linalg.norm(ax, axis=-1, keepdims=True)
        ang = rng.normal(0, tilt, shape)
        yaw = rng.uniform(-np.pi, np.pi, shape)
        q1 = np.concatenate([np.cos(ang / 2)[..., None], np.sin(ang / 2)[..., None] * ax], -1)
        qz = np.stack([np.cos(yaw / 2), 0 * yaw, 0 * yaw, np.sin(yaw / 2)], -1)
        from zbot_lab_b200.assets import zbot_6s as Z
        return Z.quat_mul(qz, q1), yaw

    out = []
    for t in range(steps):
        root_quat, yaw = rq((n,), 0.15)
        from zbot_lab_b200.assets import zbot_6s as Z
        # feet link frames: link y up / down (foot0 / foot1), link x forward, plus a small tilt
        base_f = np.array([[0.70710678, 0.70710678, 0, 0], [0.70710678, -0.70710678, 0, 0]])
        tilt, _ = rq((n, 2), 0.2)
        qz = np.stack([np.cos(yaw / 2), 0 * yaw, 0 * yaw, np.sin(yaw / 2)], -1)
        feet_quat = Z.quat_mul(Z.quat_mul(np.broadcast_to(qz[:, None, :], (n, 2, 4)), tilt), np.broadcast_to(base_f, (n, 2, 4)))
        feet_quat = feet_quat / np.linalg.norm(feet_quat, axis=-1, keepdims=True)
        root_pos = np.stack([rng.normal(0, 0.3, n), rng.normal(0, 0.3, n), rng.normal(0.245, 0.03, n)], -1)
        feet_pos = root_pos[:, None, :] + np.stack([rng.normal(0, 0.05, (n, 2)), rng.normal(0, 0.03, (n, 2)) + np.array([-0.065, 0.065]),
                                                    -root_pos[:, None, 2].repeat(2, 1) + 0.053 + np.abs(rng.normal(0, 0.02, (n, 2)))], -1)
        hist = np.abs(rng.normal(0, 1, (n, 3, 2, 3))) * np.array([0.3, 0.3, 12.0]) * (rng.random((n, 1, 2, 1)) < 0.7)
        v = {"root_pos": root_pos, "root_quat": root_quat, "root_lin_vel": rng.normal(0, 0.3, (n, 3)),
             "root_ang_vel": rng.normal(0, 0.5, (n, 3)), "feet_pos": feet_pos, "feet_quat": feet_quat,
             "feet_com_vel": rng.normal(0, 0.3, (n, 2, 3)), "feet_force_hist": hist,
             "feet_fz_hist": hist[..., 2], "feet_fnorm_max": np.linalg.norm(hist, axis=-1).max(1),
             "last_air": rng.uniform(0, 0.8, (n, 2)), "last_contact": rng.uniform(0, 0.8, (n, 2)),
             "cur_air": rng.uniform(0, 0.6, (n, 2)), "cur_contact": rng.uniform(0, 0.6, (n, 2)),
             "q_chain": rng.normal(0, 0.5, (n, 6)), "qd_chain": rng.normal(0, 1.0, (n, 6)), "tau": rng.normal(0, 5, (n, 6)),
             "joint_acc": rng.normal(0, 50, (n, 6))}
        inair = rng.random((n, 2)) < 0.5
        v["cur_contact"] = np.where(inair, 0.0, v["cur_contact"])
        v["cur_air"] = np.where(inair, v["cur_air"], 0.0)
        v = {k: np.asarray(x, F) for k, x in v.items()}
        # 13 uniforms from the main stream (the golden fixtures were generated with exactly this draw order) + the 9 slots
        # added later (heading command, push event) from a side stream
        u13 = rng.normal(0, 1, (n, 6)).astype(F), rng.random((n, 13)).astype(F)
        u8 = np.random.default_rng(100003 * seed + t).random((n, 9)).astype(F)
        out.append((v, u13[0], np.concatenate([u13[1], u8], axis=1)))
    return out
