"""``Zbot6SEnvV4`` -- drop-in for the reference task class of ``zbot-6b-walking-v4``
(``/root/reference/source/zbot/zbot/tasks/zbot6b_direct/zbot_direct_6dof_bipedal_env_v4.py:560-1304``) over the fused
sm_100a step (``zbot_v4_step_kernel``): velocity / heading commands with reset- and interval-mode resampling,
randomised reset pose, 24-wide observation, 15 cfg-driven reward terms, and the reference's two host-side curricula
(``my_curriculum`` :138-198, ``range_curriculum`` :200-265).

Per step everything is ONE kernel launch; the event random numbers come from the kernel's counter-based generator
(seeded from ``cfg.seed``) unless ``step(actions, rand=...)`` supplies the (N,10) uniforms.  The curricula only
touch host integers, except ``range_curriculum`` which reads the statistics ring once every
``12 * max_episode_length`` steps (the reference syncs on every reset step)."""
from __future__ import annotations

from collections import deque

import torch

from ... import native
from .walking_v2 import ZbotDirectEnvV2
from .walking_v4_cfg import Zbot6SEnvV4Cfg


class Zbot6SEnvV4(ZbotDirectEnvV2):
    cfg: Zbot6SEnvV4Cfg

    _TASK = native.TASK_WALKING_V4
    _TERM_IDS = native.V4_TERM_IDS
    _HOST_TERMS_SUPPORTED = False
    _DIED_LOG_KEY = "Episode_Termination/died"          # …env_v4.py:918

    def __init__(self, cfg: Zbot6SEnvV4Cfg | None = None, render_mode: str | None = None, **kwargs):
        cfg = cfg if cfg is not None else Zbot6SEnvV4Cfg()
        # live copies of what the curricula mutate (the reference mutates its cfg dict / EventManager term params)
        self._live_scales = dict(cfg.reward_cfg["reward_scales"])      # bare weights: v4 applies step_dt per evaluation
        cr = cfg.events.command_resample
        self.event_params = {"velocity_range": tuple(cr.velocity_range), "yaw_range": tuple(cr.yaw_range),
                             "dual_sign": bool(cr.dual_sign), "offset": float(cr.offset), "prob_pos": float(cr.prob_pos)}
        self.curriculum_stage = 0
        self.curriculum_vel_reward_buffer = deque(maxlen=24)
        self.curriculum_yaw_reward_buffer = deque(maxlen=24)
        super().__init__(cfg, render_mode, **kwargs)
        self.reward_scales = self._live_scales

    # ------------------------------------------------------------------ native cfg
    def _native_cfg(self) -> native.ZbotCfg:
        c, a, ev = self.cfg.contact, self.cfg.actuator, self.cfg.events
        pr = ev.reset_base.pose_range
        lo = [pr.get(k, (0.0, 0.0))[0] for k in ("x", "y", "yaw")]
        hi = [pr.get(k, (0.0, 0.0))[1] for k in ("x", "y", "yaw")]
        p = self.event_params
        seed = self.cfg.seed if self.cfg.seed is not None else int(torch.initial_seed() & 0x7FFFFFFF)
        return native.set_obs_noise(native.make_cfg(
            self.num_envs, reward_scales=self._live_scales, step_dt=self.step_dt, task=self._TASK,
            sim_dt=self.physics_dt, decimation=int(self.cfg.decimation), max_episode_length=int(self.max_episode_length),
            termination_height=float(self.cfg.termination_height), kp=a.stiffness, kd=a.damping,
            effort_limit=a.effort_limit, gravity=-float(self.cfg.sim.gravity[2]),
            contact_alpha=c.alpha, contact_erp=c.erp, contact_vdep=c.max_depenetration_velocity,
            contact_beta_max=c.beta_max, contact_mu=c.friction, contact_ramp=c.ramp, contact_margin=c.margin,
            ev_vel_lo=p["velocity_range"][0], ev_vel_hi=p["velocity_range"][1], ev_yaw_lo=p["yaw_range"][0],
            ev_yaw_hi=p["yaw_range"][1], ev_offset=p["offset"], ev_prob_pos=p["prob_pos"], ev_dual_sign=int(p["dual_sign"]),
            ev_pose_lo=lo, ev_pose_hi=hi, ev_interval_lo=float(ev.command_resample.interval_range_s[0]),
            ev_interval_hi=float(ev.command_resample.interval_range_s[1]), rng_seed=int(seed)),
            getattr(self.cfg, "observation_noise", None))

    def _push_cfg(self):
        """Re-derive the kernel's weight table / event parameters after a curriculum changed them."""
        new = self._native_cfg()
        st = self._stepper
        for f, _ in native.ZbotCfg._fields_:
            setattr(st.cfg, f, getattr(new, f))
        st.update_cfg()
        self._term_names = list(self.reward_scales.keys())

    def _initial_reset(self):
        st = self._stepper
        st.reset_idx_v4(None)
        lo, hi = self.cfg.events.command_resample.interval_range_s
        # [IL-upstream] EventManager: per-env interval timers start at U(lower, upper)
        st.state.set("base_pos_y_err_sum", torch.rand(self.num_envs, 1, device=self.device) * (hi - lo) + lo)

    # ------------------------------------------------------------------ reference attribute surface
    @property
    def commands(self):
        return self._stepper.state.get("carry_feet_fz")

    @property
    def target_heading_yaw(self):
        return self._stepper.state.get("carry_mid_max")[:, 0]

    @property
    def current_yaw(self):
        return self._stepper.state.get("base_heading_x_sum")[:, 0]

    def __getattr__(self, name):
        raise AttributeError(name)

    # ------------------------------------------------------------------ curricula (host)
    def _my_curriculum(self):
        """``my_curriculum`` (…env_v4.py:138-198): thresholds on the global step counter."""
        L, s, changed = self.max_episode_length, self.reward_scales, False
        if self.common_step_counter >= L * 12 and self.curriculum_stage == 0:
            s["airtime_variance"], s["feet_forward"], s["feet_slide"] = -10.0, -1.0, -2.0
            self.curriculum_stage += 1
            changed = True
        elif self.common_step_counter >= L * 24 and self.curriculum_stage == 1:
            s["airtime_variance"], s["feet_downward"] = -40.0, -5.0
            self.event_params["prob_pos"] = 0.8
            self.curriculum_stage += 1
            changed = True
        elif self.common_step_counter >= L * 144 and self.curriculum_stage == 2:
            s["feet_harmony"], s["feet_downward"], s["step_length"], s["track_heading_yaw"] = 1.0, -10.0, 7.0, 2.0
            self.event_params["prob_pos"] = 0.6
            s["feet_close"] = -120.0
            self.curriculum_stage += 1
            changed = True
        return changed

    def _refresh_curriculum_buffers(self):
        """Rebuild the two 24-deep reward buffers (…env_v4.py:903-906) from the statistics ring: one entry per
        recent step in which some env reset (one host read of the ring)."""
        st = self._stepper
        ring = st.stats_ring.cpu()
        order = [(st._slot - i) % ring.shape[0] for i in range(ring.shape[0])]
        names = self._term_names
        iv, iy = names.index("track_lin_vel_x") if "track_lin_vel_x" in names else -1, \
            names.index("track_heading_yaw") if "track_heading_yaw" in names else -1
        vel, yaw = [], []
        for sl in order:
            if ring[sl, native.STAT_NUM_RESET] > 0 and len(vel) < 24:
                if iv >= 0:
                    vel.append(float(ring[sl, iv]))
                if iy >= 0:
                    yaw.append(float(ring[sl, iy]))
        self.curriculum_vel_reward_buffer = deque(reversed(vel), maxlen=24)
        self.curriculum_yaw_reward_buffer = deque(reversed(yaw), maxlen=24)

    def _range_curriculum(self):
        """``range_curriculum`` (…env_v4.py:200-265): every 12 episodes' worth of steps after 48, widen the command
        ranges by 0.05 while the tracking rewards stay above 85 % of their weight."""
        L = self.max_episode_length
        if not (self.common_step_counter >= L * 48 and self.common_step_counter % (L * 12) == 0):
            return False
        self._refresh_curriculum_buffers()
        if len(self.curriculum_vel_reward_buffer) < 20:
            return False
        lim, lim_yaw = self.cfg.events.vel_range.limit_ranges, self.cfg.events.vel_range.limit_yaw_ranges
        changed = False
        clamp = lambda x, lo, hi: min(max(x, lo), hi)
        r = sum(self.curriculum_vel_reward_buffer) / len(self.curriculum_vel_reward_buffer)
        if r > self.reward_scales["track_lin_vel_x"] * 0.85:
            lo, hi = self.event_params["velocity_range"]
            self.event_params["velocity_range"] = (clamp(lo - 0.05, *lim), clamp(hi + 0.05, *lim))
            changed = True
        if self.curriculum_yaw_reward_buffer:
            r = sum(self.curriculum_yaw_reward_buffer) / len(self.curriculum_yaw_reward_buffer)
            if r > self.reward_scales["track_heading_yaw"] * 0.85:
                lo, hi = self.event_params["yaw_range"]
                self.event_params["yaw_range"] = (clamp(lo - 0.05, *lim_yaw), clamp(hi + 0.05, *lim_yaw))
                changed = True
        return changed

    def advance_host_curricula(self, steps: int) -> bool:
        """The curricula of `steps` control steps that ran inside a replayed rollout graph (no Python per step): the
        step-count thresholds of `my_curriculum` / `range_curriculum` are checked for every replayed step; a change takes
        effect from the next rollout (<= steps - 1 control steps later than in the eager loop)."""
        changed = False
        for _ in range(int(steps)):
            self.common_step_counter += 1
            if self.cfg.events.my_curric:
                changed |= bool(self._my_curriculum() | self._range_curriculum())
        self._sim_step_counter += int(steps) * int(self.cfg.decimation)
        if changed:
            self._push_cfg()
            self._log_cache.clear()
        return changed

    # ------------------------------------------------------------------ log
    def curriculum_log(self) -> dict:
        return {"Curriculum/curriculum_stage": self.curriculum_stage,                  # …env_v4.py:925-933
                "Curriculum/vel_lower_bound": self.event_params["velocity_range"][0],
                "Curriculum/vel_upper_bound": self.event_params["velocity_range"][1],
                "Curriculum/yaw_bound": self.event_params["yaw_range"][0]}

    def _log_from_slot(self) -> dict:
        log = dict(super()._log_from_slot())
        log.update(self.curriculum_log())
        return log

    # ------------------------------------------------------------------ gym API
    def reset(self, seed: int | None = None, options=None):
        if seed is not None:
            self.seed(seed)
        self._initial_reset()
        self.episode_length_buf = torch.randint_like(self._stepper.episode_length_buf, high=int(self.max_episode_length))
        self.extras["log"] = self._log_from_slot()
        return {"policy": self._stepper.observe().clone()}, self.extras

    def step(self, actions: torch.Tensor, rand: torch.Tensor | None = None):
        st = self._stepper
        self._out_i = (self._out_i + 1) % len(self._out)
        st.obs, st.rew, st.terminated, st.truncated = self._out[self._out_i]
        self.common_step_counter += 1
        # reset-mode curricula run inside the step's reset handling in the reference; their effect (new weights /
        # ranges) applies from this step's resampling on -- push before the launch
        if self.cfg.events.my_curric and self._my_curriculum() | self._range_curriculum():
            self._push_cfg()
        obs, rew, term, trunc = st.step(actions.to(self.device), rand=rand)
        self._sim_step_counter += self.cfg.decimation
        self.reset_terminated = term.view(torch.bool)
        self.reset_time_outs = trunc.view(torch.bool)
        self.extras["log"] = self._log_from_slot()
        return {"policy": obs}, rew, self.reset_terminated, self.reset_time_outs, self.extras

    def step_host(self, *a, **k):
        raise NotImplementedError("step_host is implemented for zbot-6b-walking-v2")
