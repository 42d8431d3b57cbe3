"""``ManagerBasedRLEnv`` -- drop-in for ``isaaclab.envs:ManagerBasedRLEnv`` as the reference registers it for
``zbot-6b-walking-m-v0`` (``/root/reference/source/zbot/zbot/tasks/zbotlab_manager/config/zbot6b_manager/__init__.py:14-22``),
over the fused sm_100a step (``zbot_m_step_kernel``).

The managers of Isaac Lab are *compiled away*: at construction the cfg's ``rewards`` / ``terminations`` / ``events`` /
``commands`` / ``actions`` / ``observations`` sections are read term by term (function NAME, weight, params) and turned
into the kernel's term table (``native.make_m_cfg``); every step is then ONE launch that does what
``ManagerBasedRLEnv.step`` does ([IL-upstream] order): action term at every substep, physics x decimation with the contact
sensor, terminations, rewards, reset of the done envs (reset events + command resample), command update, observation
group with additive uniform noise.  A term the kernel does not implement raises ``NotImplementedError`` here -- loud,
never dropped.  ``extras["log"]`` carries ``Episode_Reward/<term>``, ``Episode_Termination/<term>`` and
``Curriculum/lin_vel_cmd_levels`` like the managers' reset logs.  Host-side: only the ``lin_vel_cmd_levels`` curriculum
(``mdp/curriculums.py:57-83``), evaluated once per ``max_episode_length`` steps as in the reference."""
from __future__ import annotations

import math

import torch

from ... import native
from ...stepper import NativeStepper
from ..zbot6b_direct.walking_v2 import ZbotDirectEnvV2, _Box, _Terrain
from .env_cfg import Zbot6BFlatEnvCfg

#: observation term function -> (first column, width) of the 25-wide policy group
_OBS_LAYOUT = (("root_quat_w", 4), ("generated_commands", 3), ("joint_pos_rel", 6), ("joint_vel_rel", 6), ("last_action", 6))


def _terms(section):
    """(name, term) pairs of a cfg section in declaration order, ``None`` entries skipped (as the managers do)."""
    if section is None:
        return []
    out = []
    for k, v in vars(section).items():
        if k.startswith("_") or v is None or not hasattr(v, "func"):
            continue
        out.append((k, v))
    return out


class ManagerBasedRLEnv(ZbotDirectEnvV2):
    metadata = {"render_modes": [None]}

    def __init__(self, cfg=None, render_mode: str | None = None, **kwargs):
        self.cfg = cfg if cfg is not None else Zbot6BFlatEnvCfg()
        if render_mode not in (None,):
            raise NotImplementedError("rendering is out of scope of the B200 step (SURVEY.md §2 row 2)")
        self.render_mode = render_mode
        c = self.cfg
        self.device = torch.device(c.sim.device)
        self.num_envs = int(c.scene.num_envs)
        if c.seed is not None:
            self.seed(c.seed)
        self.physics_dt = float(c.sim.dt)
        self.step_dt = self.physics_dt * c.decimation
        self.max_episode_length_s = float(c.episode_length_s)
        self.max_episode_length = math.ceil(self.max_episode_length_s / self.step_dt)
        ttype = getattr(c.scene.terrain, "terrain_type", "plane")
        if ttype not in ("plane", "generator"):
            raise NotImplementedError(f"terrain_type {ttype!r}: 'plane' (flat cfgs) and 'generator' (rough cfgs) are built")
        self._compile_cfg()
        self._stepper = NativeStepper(self.num_envs, self.device, self._native_cfg())
        if ttype == "generator":
            self._setup_generated_terrain()
        else:
            if self._curr_terrain:
                raise NotImplementedError("terrain_levels_vel needs terrain_type='generator' (mdp/curriculums.py:35-37)")
            self._terrain = _Terrain(self.num_envs, c.scene.env_spacing, self.device)
        self.single_observation_space = {"policy": _Box((native.M_NUM_OBS,))}
        self.single_action_space = _Box((6,))
        self.observation_space = {"policy": _Box((self.num_envs, native.M_NUM_OBS))}
        self.action_space = _Box((self.num_envs, 6))
        self.common_step_counter = 0
        self.extras: dict = {}
        self._log_cache: dict = {}
        self._host_views: dict = {}
        ring = max(2, int(getattr(c, "output_ring", 4)))
        n, dev, no = self.num_envs, self.device, native.M_NUM_OBS
        self._num_obs = no
        self._packed = [torch.zeros(n * (no + 1) * 4 + 2 * n, dtype=torch.uint8, device=dev) for _ in range(ring)]
        self._out = [(b[:n * no * 4].view(torch.float32).view(n, no), b[n * no * 4:n * (no + 1) * 4].view(torch.float32),
                      b[n * (no + 1) * 4:n * (no + 1) * 4 + n], b[n * (no + 1) * 4 + n:n * (no + 1) * 4 + 2 * n])
                     for b in self._packed]
        self._out_i = 0
        self.reset_terminated = torch.zeros(n, dtype=torch.bool, device=dev)
        self.reset_time_outs = torch.zeros(n, dtype=torch.bool, device=dev)
        self.reset_buf = torch.zeros(n, dtype=torch.bool, device=dev)
        self._check_all_reset = False
        self._startup_events()
        self._initial_reset()
        self._sim_step_counter = 0

    # ------------------------------------------------------------------ cfg -> term table
    def _compile_cfg(self):
        c = self.cfg
        # rewards (RewardManager: zero-weight terms are skipped)
        self._reward_terms, self._is_terminated_weight = [], 0.0
        for name, t in _terms(c.rewards):
            fn = t.func.__name__
            if float(t.weight) == 0.0:
                continue
            if fn == "is_terminated":
                self._is_terminated_weight, self._is_terminated_name = float(t.weight), name
                continue
            params = {k: v for k, v in dict(t.params).items() if k not in ("asset_cfg", "sensor_cfg")}
            self._reward_terms.append((name, fn, float(t.weight), params))
        self._term_names = [n for n, _, _, _ in self._reward_terms]
        # terminations
        self._min_height, self._feet_close, self._illegal = -1.0e30, 0.0, None
        self._done_names = []
        for name, t in _terms(c.terminations):
            fn = t.func.__name__
            if fn == "time_out":
                if not t.time_out:
                    raise NotImplementedError("mdp.time_out must be a time_out=True term")
            elif fn == "root_height_below_minimum":
                self._min_height = float(t.params["minimum_height"])
            elif fn == "feet_close":
                self._feet_close = float(t.params["minimum_distance"])
            elif fn == "illegal_contact":
                # isaaclab.envs.mdp.illegal_contact [IL-upstream] (zbotlab_env_cfg.py:385-388): the sensor bodies the pattern
                # selects -> the merged bodies of the reduced chain that sense for them (a merged body reports on its "a" link)
                self._illegal = (float(t.params.get("threshold", 1.0)), self._merged_body_mask(t.params["sensor_cfg"].body_names))
            else:
                raise NotImplementedError(f"termination term {fn!r} is not built into the fused step")
            self._done_names.append((name, fn))
        # events
        self._pose_range, self._friction_range, self._push = ((0.0, 0.0),) * 3, None, None
        for name, t in _terms(c.events):
            fn, p = t.func.__name__, dict(t.params)
            if fn in ("init_my_data", "reset_my_data"):
                continue
            if fn == "randomize_rigid_body_material":
                self._friction_range = (tuple(p["dynamic_friction_range"]), int(p.get("num_buckets", 64)))
            elif fn == "reset_root_state_uniform":
                pr = p.get("pose_range", {})
                if any(tuple(v) != (0.0, 0.0) for v in p.get("velocity_range", {}).values()) or \
                        any(k in pr and tuple(pr[k]) != (0.0, 0.0) for k in ("z", "roll", "pitch")):
                    raise NotImplementedError("reset_root_state_uniform: only x / y / yaw pose ranges are built")
                self._pose_range = tuple(tuple(pr.get(k, (0.0, 0.0))) for k in ("x", "y", "yaw"))
            elif fn == "push_by_setting_velocity":
                if getattr(t, "mode", "interval") != "interval":
                    raise NotImplementedError("push_by_setting_velocity is built as an interval-mode event")
                vr = dict(p.get("velocity_range", {}))
                if any(k not in ("x", "y") and tuple(v) != (0.0, 0.0) for k, v in vr.items()):
                    raise NotImplementedError("push_by_setting_velocity: only x / y velocity ranges are built")
                self._push = {"interval_range_s": tuple(t.interval_range_s), "velocity_range": {k: tuple(v) for k, v in vr.items()}}
            elif fn == "reset_joints_by_scale":
                if tuple(p["position_range"]) != (1.0, 1.0) or tuple(p["velocity_range"]) != (1.0, 1.0):
                    raise NotImplementedError("reset_joints_by_scale: only the (1, 1) scale of the reference cfg is built")
            else:
                raise NotImplementedError(f"event term {fn!r} is not built into the fused step")
        # command / action / observation terms
        cmd = c.commands.base_velocity
        self._heading = None
        if cmd.heading_command:          # zbotlab_env_cfg.py:86-97 (the commented-out alternative of the reference cfg)
            self._heading = {"range": tuple(cmd.ranges.heading), "stiffness": float(getattr(cmd, "heading_control_stiffness", 1.0)),
                             "rel_heading_envs": float(cmd.rel_heading_envs)}
        self._cmd = cmd
        act = c.actions.joint_pos
        if type(act).__name__ != "RelativeJointPositionActionCfg" or not act.use_zero_offset:
            raise NotImplementedError("only RelativeJointPositionActionCfg(use_zero_offset=True) is built")
        clip = None
        if act.clip:
            lo, hi = next(iter(act.clip.values()))
            if abs(lo + hi) > 1e-12 or len({tuple(v) for v in act.clip.values()}) != 1:
                raise NotImplementedError("action clip must be one symmetric range")
            clip = float(hi)
        self._act = (float(act.scale), clip)
        pol = c.observations.policy
        got = [(n, t.func.__name__) for n, t in _terms(pol)]
        if [f for _, f in got] != [f for f, _ in _OBS_LAYOUT]:
            raise NotImplementedError(f"the policy observation group must be {[f for f, _ in _OBS_LAYOUT]}, got {[f for _, f in got]}")
        self._obs_noise = {}
        col = 0
        for (name, t), (_, width) in zip(_terms(pol), _OBS_LAYOUT):
            if pol.enable_corruption and t.noise is not None:
                self._obs_noise[(col, col + width)] = (float(t.noise.n_min), float(t.noise.n_max))
            col += width
        self._curr_lin_vel = any(t.func.__name__ == "lin_vel_cmd_levels" for _, t in _terms(c.curriculum))
        self._curr_terrain = any(t.func.__name__ == "terrain_levels_vel" for _, t in _terms(c.curriculum))
        for _, t in _terms(c.curriculum):
            if t.func.__name__ not in ("lin_vel_cmd_levels", "terrain_levels_vel"):
                raise NotImplementedError(f"curriculum term {t.func.__name__!r} is not built")

    def _native_cfg(self) -> native.ZbotCfg:
        c, r, cm = self.cfg, self.cfg.scene.robot, self._cmd
        seed = c.seed if c.seed is not None else int(torch.initial_seed() & 0x7FFFFFFF)
        ct = c.contact
        cfg = native.make_m_cfg(
            self.num_envs, [(f, w, p) for _, f, w, p in self._reward_terms], is_terminated_weight=self._is_terminated_weight,
            minimum_height=self._min_height, feet_close_min=self._feet_close,
            cmd_ranges=(tuple(cm.ranges.lin_vel_x), tuple(cm.ranges.lin_vel_y), tuple(cm.ranges.ang_vel_z)),
            rel_standing_envs=float(cm.rel_standing_envs), resampling_time_range=tuple(cm.resampling_time_range),
            act_scale=self._act[0], act_clip=self._act[1], pose_range=self._pose_range,
            episode_length_s=self.max_episode_length_s, friction=float(c.scene.terrain.dynamic_friction), rng_seed=int(seed),
            sim_dt=self.physics_dt, decimation=int(c.decimation), kp=float(r.stiffness), kd=float(r.damping),
            effort_limit=float(r.effort_limit), gravity=-float(c.sim.gravity[2]), contact_alpha=ct.alpha, contact_erp=ct.erp,
            contact_vdep=ct.max_depenetration_velocity, contact_beta_max=ct.beta_max, contact_ramp=ct.ramp,
            contact_margin=ct.margin, illegal_contact=self._illegal, heading=self._heading, push=self._push)
        for i in range(24):
            cfg.obs_noise_lo[i] = cfg.obs_noise_hi[i] = 0.0
        cfg.obs_noise_enable = 1 if self._obs_noise else 0
        for (a, b), (lo, hi) in self._obs_noise.items():
            for i in range(a, min(b, 24)):
                cfg.obs_noise_lo[i], cfg.obs_noise_hi[i] = lo, hi
        return cfg

    def _setup_generated_terrain(self):
        """TerrainImporterCfg(terrain_type="generator") (zbotlab_env_cfg.py:44-48): generate the height field
        (``zbot_lab_b200/terrain.py``), place the envs on their (level, type) tiles as TerrainImporter does, and bind both
        to the kernel (``zbot_bind_terrain``); the ``terrain_levels`` curriculum then runs inside the step kernel."""
        import numpy as np
        from ...terrain import Terrain
        tc = self.cfg.scene.terrain
        gen = tc.terrain_generator
        gen.curriculum = bool(self._curr_terrain)          # ZbotLabRoughEnvCfg.__post_init__ (zbotlab_env_cfg.py:445-452)
        seed = self.cfg.seed if self.cfg.seed is not None else 0
        ter = Terrain(gen, seed=int(seed))
        rng = np.random.default_rng(int(seed) + 7919)
        levels, types = ter.initial_levels_types(self.num_envs, tc.max_init_terrain_level, rng)
        dev = self.device
        org4 = np.zeros((self.num_envs, 4), np.float32)
        org4[:, :3] = ter.origins[levels, types]
        self._terrain_gen = ter
        self._terrain_heights = torch.from_numpy(ter.heights).to(dev)
        self._terrain_tile_origins = torch.from_numpy(ter.origins).to(dev).contiguous()
        self._env_origins4 = torch.from_numpy(org4).to(dev)
        self._terrain = type("TerrainImporter", (), {})()
        self._terrain.env_origins = self._env_origins4[:, :3]         # a VIEW: the kernel rewrites it when an env changes level
        self._terrain.terrain_origins = self._terrain_tile_origins
        self._terrain.cfg = tc
        self._stepper.bind_terrain(self._terrain_heights, ter.x0, ter.y0, ter.cell, self._terrain_tile_origins, float(gen.size[0]),
                                   self._env_origins4, bool(self._curr_terrain))
        pd = self._stepper.state.get("p_delta")
        pd[:, 3] = torch.from_numpy(levels.astype(np.float32)).to(dev)
        pd[:, 4] = torch.from_numpy(types.astype(np.float32)).to(dev)
        self._stepper.state.set("p_delta", pd)

    @property
    def terrain_levels(self) -> torch.Tensor:
        """``scene.terrain.terrain_levels`` (generator terrains): the level of every env."""
        return self._stepper.state.column("p_delta", 3).to(torch.int64)

    @staticmethod
    def _merged_body_mask(body_names) -> int:
        """Bit b-1 for every merged body 1..5 of the reduced chain one of whose links matches the sensor pattern."""
        import re
        from ...assets import zbot_6s_v2 as V
        pats = [body_names] if isinstance(body_names, str) else list(body_names)
        m = V.model_f32()
        mask = 0
        for li, name in enumerate(V.LINK_NAMES):
            if any(re.fullmatch(p, name) for p in pats):
                b = int(m.link_body[li])
                if b in (0, 6):
                    raise NotImplementedError(f"illegal_contact on a foot link ({name!r}) is not built (feet contact is the gait)")
                mask |= 1 << (b - 1)
        if mask == 0:
            raise ValueError(f"illegal_contact: no link matches {body_names!r}")
        return mask

    def _startup_events(self):
        """EventManager mode "startup": per-env friction from 64 buckets (randomize_rigid_body_material), combined with
        the ground material by "multiply" (zbotlab_env_cfg.py:50-55)."""
        if self._friction_range is None:
            return
        (lo, hi), buckets = self._friction_range
        table = torch.rand(buckets, device=self.device) * (hi - lo) + lo
        mu = table[torch.randint(0, buckets, (self.num_envs,), device=self.device)]
        self._stepper.state.set("joint_speed_limit", (mu * float(self.cfg.scene.terrain.dynamic_friction)).unsqueeze(-1))

    def _initial_reset(self):
        self._stepper.reset_idx_m(None)

    # ------------------------------------------------------------------ attribute surface
    @property
    def friction(self) -> torch.Tensor:
        return self._stepper.state.get("joint_speed_limit")[:, 0]

    @property
    def command(self) -> torch.Tensor:
        """``command_manager.get_command("base_velocity")``: (N, 3) lin_vel x / y, ang_vel z."""
        g = self._stepper.state.get
        return torch.cat([g("carry_feet_fz"), g("carry_mid_max")], dim=-1)

    def __getattr__(self, name):
        raise AttributeError(name)

    # ------------------------------------------------------------------ curriculum (host), mdp/curriculums.py:57-83
    def _lin_vel_cmd_levels(self, slot: int | None = None) -> bool:
        """Returns True when the command ranges (hence the kernel parameters) changed.  `slot`: statistics slot of the
        step the curriculum is evaluated for (default: the most recent one)."""
        if not self._curr_lin_vel or self.common_step_counter % self.max_episode_length != 0:
            return False
        names = self._term_names
        if "track_lin_vel_xy_exp" not in names:
            return False
        i = names.index("track_lin_vel_xy_exp")
        slot = max(self._stepper._slot, 0) if slot is None else slot
        s = self._stepper.stats_ring[slot].cpu()                 # one host read per max_episode_length steps
        if float(s[native.STAT_NUM_RESET]) == 0:                 # CurriculumManager.compute runs inside _reset_idx: no reset, no update
            return False
        reward = float(s[i])                                     # mean episodic sum of the reset envs / episode seconds
        weight = self._reward_terms[i][2]
        if reward > weight * 0.8:
            r, lim = self._cmd.ranges, self._cmd.limit_ranges
            clamp = lambda v, l: (min(max(v[0] - 0.1, l[0]), l[1]), min(max(v[1] + 0.1, l[0]), l[1]))
            r.lin_vel_x = clamp(tuple(r.lin_vel_x), tuple(lim.lin_vel_x))
            r.lin_vel_y = clamp(tuple(r.lin_vel_y), tuple(lim.lin_vel_y))
            new = self._native_cfg()
            for f, _ in native.ZbotCfg._fields_:
                setattr(self._stepper.cfg, f, getattr(new, f))
            self._stepper.update_cfg()
            return True
        return False

    def advance_host_curricula(self, steps: int) -> bool:
        """`steps` control steps ran inside a replayed rollout graph (no Python per step; `rl/ppo_runner.py`): advance the
        global step counter and evaluate `lin_vel_cmd_levels` for the replayed step that crossed a multiple of
        max_episode_length, on that step's own statistics slot.  True = kernel parameters changed (re-capture)."""
        changed = False
        last = self._stepper._slot                              # slot of the LAST replayed step
        slots = self._stepper.stats_ring.shape[0]
        for t in range(int(steps)):
            self.common_step_counter += 1
            changed |= self._lin_vel_cmd_levels((last - (int(steps) - 1 - t)) % slots)
        self._sim_step_counter += int(steps) * int(self.cfg.decimation)
        return changed

    def curriculum_log(self) -> dict:
        return {"Curriculum/lin_vel_cmd_levels": float(self._cmd.ranges.lin_vel_x[1])}

    # ------------------------------------------------------------------ log (managers' reset logs [IL-upstream])
    def _log_from_slot(self) -> dict:
        slot = max(self._stepper._slot, 0)
        log = self._log_cache.get(slot)
        if log is None:
            s = self._stepper.stats_ring[slot]
            log = {"Episode_Reward/" + k: s[i] for i, k in enumerate(self._term_names)}
            # statistics words 22..25 (include/zbot_b200.h): is_terminated's episodic sum, per-DoneTerm reset counts
            if self._is_terminated_weight:
                log["Episode_Reward/" + self._is_terminated_name] = s[22]
            word = {"root_height_below_minimum": 23, "feet_close": 24, "illegal_contact": 25}
            for name, fn in self._done_names:
                log["Episode_Termination/" + name] = s[native.STAT_NUM_TIMEOUT_RESET if fn == "time_out" else word[fn]]
            self._log_cache[slot] = log
        log["Curriculum/lin_vel_cmd_levels"] = float(self._cmd.ranges.lin_vel_x[1])
        if self._curr_terrain:                                   # terrain_levels_vel returns the mean level (mdp/curriculums.py:55)
            log["Curriculum/terrain_levels"] = self._stepper.state.column("p_delta", 3).mean()
        return log

    # ------------------------------------------------------------------ gym API
    def reset(self, seed: int | None = None, options=None):
        if seed is not None:
            self.seed(seed)
        self._initial_reset()
        self.extras["log"] = self._log_from_slot()
        return {"policy": self._stepper.observe().clone()}, self.extras

    def step(self, action: torch.Tensor, rand: torch.Tensor | None = None):
        st = self._stepper
        self._out_i = (self._out_i + 1) % len(self._out)
        st.obs, st.rew, st.terminated, st.truncated = self._out[self._out_i]
        obs, rew, term, trunc = st.step(action.to(self.device), rand=rand)
        self.common_step_counter += 1
        self._sim_step_counter += self.cfg.decimation
        self.reset_terminated = term.view(torch.bool)
        self.reset_time_outs = trunc.view(torch.bool)
        self._lin_vel_cmd_levels()
        self.extras["log"] = self._log_from_slot()
        return {"policy": obs}, rew, self.reset_terminated, self.reset_time_outs, self.extras

    def step_host(self, *a, **k):
        raise NotImplementedError("step_host is implemented for zbot-6b-walking-v2")
