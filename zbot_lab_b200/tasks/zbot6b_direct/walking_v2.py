"""``ZbotDirectEnvV2`` -- drop-in for the reference task class of ``zbot-6b-walking-v2``
(``/root/reference/source/zbot/zbot/tasks/zbot6b_direct/zbot_direct_6dof_bipedal_env_v2.py:208-605``)
over the fused sm_100a step.  Same constructor signature, ``step`` 5-tuple, ``reset`` 2-tuple,
attribute names and ``extras["log"]`` keys; the per-step work of the reference's
``_pre_physics_step / _apply_action / _get_dones / _get_rewards / _reset_idx / _get_observations``
plus Isaac Lab's articulation / contact-sensor stepping happens in ONE kernel launch.
"""
from __future__ import annotations

import math

import torch

from ... import native
from ...assets import zbot_6s as Z
from ...stepper import NativeStepper
from ...utils import synthetic as syn
from .host_terms import ContactSensor, RewardTermViews, StepView
from .walking_v2_cfg import ZbotDirectEnvCfgV2


class _Box:
    """Tiny stand-in for ``gymnasium.spaces.Box`` (only shape / bounds are ever read)."""

    def __init__(self, shape, low=-math.inf, high=math.inf):
        self.shape = tuple(shape)
        self.low, self.high = low, high
        self.dtype = "float32"

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, float32)"


class _ArticulationData:
    """Lazy ``robot.data`` view (world frame = env-local + env origin), computed on demand by the
    articulation-view kernel; names follow the call sites at …env_v2.py:315-326, 357-358, 554, 560."""

    def __init__(self, env):
        self._env = env
        n, dev = env.num_envs, env.device
        self.default_joint_pos = torch.tensor(Z.DEFAULT_JOINT_POS, dtype=torch.float32, device=dev).repeat(n, 1)
        self.default_joint_vel = torch.zeros(n, 6, device=dev)
        drs = torch.zeros(n, 13, device=dev)
        drs[:, :3] = torch.tensor(Z.DEFAULT_ROOT_POS, device=dev)
        drs[:, 3] = 1.0
        self.default_root_state = drs
        self.GRAVITY_VEC_W = torch.tensor([0.0, 0.0, -1.0], device=dev).repeat(n, 1)

    def _view(self):
        v = self._env.__dict__.get("_active_view")
        if v is not None:       # host reward terms are being evaluated: END of physics of this step, before any reset
            return v.robot.body_link_pos_w, v.robot.body_link_quat_w, v.robot.body_com_lin_vel_w
        pos, quat, vel = self._env._stepper.articulation_view()
        return pos + self._env._terrain.env_origins.unsqueeze(1), quat, vel

    @property
    def applied_torque(self):
        """ImplicitActuator bookkeeping before the last substep of the most recent step (…env_v2.py:560); needs a step view."""
        return self._env._need_view().robot.applied_torque

    @property
    def body_link_pos_w(self):
        return self._view()[0]

    @property
    def body_link_quat_w(self):
        return self._view()[1]

    @property
    def body_com_lin_vel_w(self):
        return self._view()[2]

    @property
    def joint_pos(self):
        v = self._env.__dict__.get("_active_view")
        return v.robot.joint_pos if v is not None else self._env._stepper.state.get("joint_pos")

    @property
    def joint_vel(self):
        v = self._env.__dict__.get("_active_view")
        return v.robot.joint_vel if v is not None else self._env._stepper.state.get("joint_vel")


class _Robot:
    def __init__(self, env):
        self.data = _ArticulationData(env)
        self._ALL_INDICES = torch.arange(env.num_envs, dtype=torch.long, device=env.device)
        self.body_names = list(Z.LINK_NAMES)
        self.joint_names = list(Z.JOINT_NAMES)

    def find_bodies(self, pattern):
        return Z.find_bodies(pattern, Z.LINK_NAMES)


class _Terrain:
    def __init__(self, n, spacing, device):
        self.env_origins = torch.from_numpy(syn.env_origins_grid(n, spacing)).to(device)


class ZbotDirectEnvV2(RewardTermViews):
    """``DirectRLEnv``-shaped vectorised env; all tensors live on ``cfg.sim.device``.

    Subclass hooks: the reward table is ``cfg.reward_cfg["reward_scales"]``; a name the fused kernel knows
    and that the subclass does not override is evaluated IN the kernel, any other name must have a ``_reward_<name>``
    method and is evaluated on the host after the kernel (``host_terms.py``).  ``_get_dones / _get_rewards /
    _get_observations / _reset_idx`` are host-side views of the fused step's results."""

    metadata = {"render_modes": [None]}
    cfg: ZbotDirectEnvCfgV2

    def __init__(self, cfg: ZbotDirectEnvCfgV2 | None = None, render_mode: str | None = None, **kwargs):
        self.cfg = cfg if cfg is not None else ZbotDirectEnvCfgV2()
        if render_mode not in (None,):
            raise NotImplementedError("rendering is out of scope of the B200 step (SURVEY.md §2 row 2)")
        self.render_mode = render_mode
        self.device = torch.device(self.cfg.sim.device)
        self.num_envs = int(self.cfg.scene.num_envs)
        if self.cfg.seed is not None:
            self.seed(self.cfg.seed)
        # timing
        self.physics_dt = float(self.cfg.sim.dt)
        self.step_dt = self.physics_dt * self.cfg.decimation
        self.max_episode_length_s = float(self.cfg.episode_length_s)
        self.max_episode_length = math.ceil(self.max_episode_length_s / self.step_dt)
        # reward table: a COPY scaled once by step_dt (the reference scales its class dict in place, C-3)
        self.reward_scales = {k: v * self.step_dt for k, v in self.cfg.reward_cfg["reward_scales"].items()}
        self._split_reward_terms()
        self._stepper = NativeStepper(self.num_envs, self.device, self._native_cfg())
        self._terrain = _Terrain(self.num_envs, self.cfg.scene.env_spacing, self.device)
        self._setup_scene()
        # spaces
        self.single_observation_space = {"policy": _Box((self.cfg.observation_space,))}
        self.single_action_space = _Box((self.cfg.action_space,))
        self.observation_space = {"policy": _Box((self.num_envs, self.cfg.observation_space))}
        self.action_space = _Box((self.num_envs, self.cfg.action_space))
        # buffers the scripts / wrappers read
        self.common_step_counter = 0
        self.extras: dict = {}
        self._term_names = list(self._fused_scales.keys())
        self._log_cache: dict = {}
        self._view: StepView | None = None
        self._active_view = None
        self.capture_step_view = bool(getattr(self.cfg, "capture_step_view", False)) or bool(self._host_terms)
        self._host_sums = {k: torch.zeros(self.num_envs, device=self.device) for k, _ in self._host_terms}
        self._host_log = {"Episode_Reward/" + k: torch.zeros((), device=self.device) for k, _ in self._host_terms}
        self._host_views: dict = {}
        ring = max(2, int(self.cfg.output_ring))
        n, dev = self.num_envs, self.device
        # one packed buffer per ring slot: [obs N*23 f32 | rew N f32 | terminated N u8 | truncated N u8] so a
        # host consumer can fetch a whole step result with ONE device->host copy (``last_step_packed``)
        no = self._num_obs = int(self.cfg.observation_space)
        self._packed = [torch.zeros(n * (no + 1) * 4 + 2 * n, dtype=torch.uint8, device=dev) for _ in range(ring)]
        self._out = [(b[:n * no * 4].view(torch.float32).view(n, no), b[n * no * 4:n * (no + 1) * 4].view(torch.float32),
                      b[n * (no + 1) * 4:n * (no + 1) * 4 + n], b[n * (no + 1) * 4 + n:n * (no + 1) * 4 + 2 * n])
                     for b in self._packed]
        self._out_i = 0
        self.reset_terminated = torch.zeros(n, dtype=torch.bool, device=dev)
        self.reset_time_outs = torch.zeros(n, dtype=torch.bool, device=dev)
        self.reset_buf = torch.zeros(n, dtype=torch.bool, device=dev)
        # all-envs-reset spread of the episode counters (…env_v2.py:418-422).  `check_all_envs_reset`:
        #   True  -> host sync on the reset count + torch.randint_like on the torch generator (bit-exact against the
        #            reference's call); None (default) picks this for N <= 256;
        #   None  -> N > 256: the statistics kernel decides and writes the counters ON THE DEVICE (in-kernel generator;
        #            no sync, no work unless the event fires, also inside a captured rollout graph);
        #   False -> no spread at all (deterministic counters, tests).
        chk = self.cfg.check_all_envs_reset
        self._check_all_reset = (self.num_envs <= 256) if chk is None else bool(chk)
        self._stepper.set_all_reset_spread(chk is None and not self._check_all_reset)
        self._initial_reset()
        self._sim_step_counter = 0

    # ------------------------------------------------------------------ task hooks (overridden by the snake task)
    _TASK = native.TASK_WALKING_V2
    _DIED_LOG_KEY = "Episode_Termination/body_contact"     # …env_v2.py:453
    _TERM_IDS = native.TERM_IDS                            # names the fused kernel evaluates for this task
    _HOST_TERMS_SUPPORTED = True                           # step view + host reward terms (walking-v2 export layout)

    def _split_reward_terms(self):
        """The reference resolves every key of ``reward_scales`` with ``getattr(self, "_reward_" + name)``
        (…env_v2.py:246-252).  Here: fused (kernel term table) unless the subclass defines / overrides the method."""
        self._fused_scales, self._host_terms = {}, []
        for name, w in self.cfg.reward_cfg["reward_scales"].items():
            meth = getattr(type(self), "_reward_" + name, None)
            builtin = getattr(RewardTermViews, "_reward_" + name, None)
            if name in self._TERM_IDS and (meth is None or meth is builtin):
                self._fused_scales[name] = w
            elif meth is not None:
                if not self._HOST_TERMS_SUPPORTED:
                    raise NotImplementedError(f"{type(self).__name__}: host-side reward terms ({name!r}) are implemented for "
                                              "zbot-6b-walking-v2 only; this task's terms are the kernel's")
                import warnings
                warnings.warn(f"reward term {name!r} is evaluated on the HOST ({type(self).__name__}._reward_{name}): off the fused "
                              "path -- the step runs through the export hook plus a few torch launches per such term")
                self._host_terms.append((name, float(w) * self.step_dt))
            else:
                raise AttributeError(f"{type(self).__name__} has no reward method '_reward_{name}' for the key {name!r} of "
                                     f"cfg.reward_cfg['reward_scales'] (the fused kernel knows {sorted(self._TERM_IDS)})")

    def _native_cfg(self) -> native.ZbotCfg:
        c, a = self.cfg.contact, self.cfg.actuator
        extra = {"termination_height": float(self.cfg.termination_height)} if hasattr(self.cfg, "termination_height") else {}
        if getattr(self.cfg, "observation_noise", None):
            extra["rng_seed"] = int(self.cfg.seed if self.cfg.seed is not None else torch.initial_seed() & 0x7FFFFFFF)
        return native.set_obs_noise(native.make_cfg(
            self.num_envs, reward_scales=self._fused_scales, step_dt=self.step_dt, task=self._TASK,
            sim_dt=self.physics_dt, decimation=int(self.cfg.decimation),
            max_episode_length=int(self.max_episode_length),
            kp=a.stiffness, kd=a.damping, effort_limit=a.effort_limit,
            gravity=-float(self.cfg.sim.gravity[2]),
            contact_alpha=c.alpha, contact_erp=c.erp, contact_vdep=c.max_depenetration_velocity,
            contact_beta_max=c.beta_max, contact_mu=c.friction, contact_ramp=c.ramp, contact_margin=c.margin, **extra),
            getattr(self.cfg, "observation_noise", None))

    def _initial_reset(self):
        self._stepper.reset_idx(None)

    def _setup_scene(self):
        self._robot = _Robot(self)
        self._contact_sensor = ContactSensor(self)
        # index sets, resolved by name exactly as …env_v2.py:227-230
        self._feet_ids, _ = Z.find_bodies("foot.*", Z.SENSOR_BODY_NAMES)
        self._undesired_contact_body_ids, _ = Z.find_bodies("base|a.*|b.*", Z.SENSOR_BODY_NAMES)
        self.base_body_idx = Z.find_bodies("base", Z.LINK_NAMES)[0]
        self.feet_body_idx = Z.find_bodies("foot.*", Z.LINK_NAMES)[0]

    # ------------------------------------------------------------------ reference attribute surface
    @property
    def unwrapped(self):
        return self

    @property
    def episode_length_buf(self) -> torch.Tensor:
        return self._stepper.episode_length_buf

    @episode_length_buf.setter
    def episode_length_buf(self, value):
        self._stepper.episode_length_buf.copy_(torch.as_tensor(value, device=self.device).to(torch.int64))

    @property
    def joint_speed_limit(self):
        return self._stepper.state.get("joint_speed_limit")

    @joint_speed_limit.setter
    def joint_speed_limit(self, value):
        self._stepper.state.set("joint_speed_limit", value)

    @property
    def p_delta(self):
        return self._stepper.state.get("p_delta")

    @property
    def _actions(self):
        return self._stepper.state.get("actions")

    @property
    def _previous_actions(self):
        return self._stepper.state.get("actions")

    @property
    def _episode_sums(self) -> dict:
        eps = self._stepper.state.get("episode_sums")
        return {k: eps[:, i] for i, k in enumerate(self._term_names)}

    def __getattr__(self, name):
        # MDP state attributes of the reference (…env_v2.py:231-245), served from the kernel's state
        fields = {"feet_contact_forces_last": 2, "feet_step_length": 2, "base_heading_x_sum": 1,
                  "base_pos_y_err_sum": 1, "feet_force_sum": 1}
        if name in fields and "_stepper" in self.__dict__:
            v = self._stepper.state.get(name)
            return v if fields[name] > 1 else v[:, 0]
        if name == "feet_down_pos_last" and "_stepper" in self.__dict__:
            return self._stepper.state.get(name).view(self.num_envs, 2, 3) + self._terrain.env_origins.unsqueeze(1)
        # the tensors the reference's _get_observations caches (…env_v2.py:315-345): start-of-step values of the step view
        view = self.__dict__.get("_view")
        if view is not None and name in view.stale:
            return self._need_view().stale[name]
        raise AttributeError(name)

    # ------------------------------------------------------------------ step view / host terms (host_terms.py)
    def _need_view(self) -> StepView:
        v = self.__dict__.get("_view")
        if v is None or not v.valid:
            raise RuntimeError("no step view: the articulation / contact-sensor tensors of a step are only exported when asked for -- "
                               "set `env.capture_step_view = True` (or cfg.capture_step_view) before stepping; envs with host "
                               "reward terms do it automatically")
        return v

    def _eval_host_terms(self, rew, term, trunc):
        """``reward += f() * scale`` for the terms the kernel does not evaluate, their episode sums and log entries
        (…env_v2.py:371-382, 441-447).  Runs on the step view: end-of-physics data, before the reset the kernel applied."""
        self._active_view = self._view
        try:
            done = (term.view(torch.bool) | trunc.view(torch.bool))
            cnt = done.sum()
            for name, scale in self._host_terms:
                r = getattr(self, "_reward_" + name)() * scale
                rew += r
                sums = self._host_sums[name]
                sums += r
                mean = (sums * done).sum() / cnt.clamp(min=1) / self.max_episode_length_s
                key = "Episode_Reward/" + name
                self._host_log[key].copy_(torch.where(cnt > 0, mean, self._host_log[key]))     # kept while nothing resets (:450)
                sums.masked_fill_(done, 0.0)
        finally:
            self._active_view = None

    # ------------------------------------------------------------------ DirectRLEnv hook names (views of the fused step)
    def _pre_physics_step(self, actions):
        """…env_v2.py:276-287 happens inside the kernel (``mdp_pre_physics``); kept so callers of the hook do not break."""

    def _apply_action(self):
        """…env_v2.py:309-310: the joint targets are handed to the actuator model inside the kernel, every substep."""

    def _get_dones(self):
        """(terminated, time_outs) of the most recent step (…env_v2.py:384-411, evaluated in the kernel)."""
        return self.reset_terminated, self.reset_time_outs

    def _get_rewards(self):
        """Reward of the most recent step (…env_v2.py:371-382; fused terms in the kernel + host terms)."""
        return self._out[self._out_i][1]

    def _get_observations(self):
        """{"policy": (N, 23)} of the current state (…env_v2.py:312-369, evaluated by the observe kernel)."""
        return {"policy": self._stepper.observe().clone()}

    def _reset_idx(self, env_ids):
        """…env_v2.py:413-459 for an explicit id list (the per-step partial reset happens inside the step kernel)."""
        if env_ids is None:
            env_ids = self._robot._ALL_INDICES
        env_ids = torch.as_tensor(env_ids, device=self.device, dtype=torch.int64)
        self._stepper.reset_idx(env_ids, self._out[self._out_i][2], self._out[self._out_i][3])
        if env_ids.numel() == self.num_envs:
            self.episode_length_buf = torch.randint_like(self._stepper.episode_length_buf, high=int(self.max_episode_length))
        for k, _ in self._host_terms:
            self._host_sums[k][env_ids] = 0.0
        self.extras["log"] = self._log_from_slot()

    def seed(self, seed: int = -1) -> int:
        if seed == -1:
            seed = int(torch.randint(0, 10000, (1,)).item())
        torch.manual_seed(seed)
        if self.device.type == "cuda":
            torch.cuda.manual_seed_all(seed)
        return seed

    # ------------------------------------------------------------------ log (…env_v2.py:441-459)
    def _log_from_slot(self) -> dict:
        # 0-dim VIEWS into this step's statistics slot (no launch, no sync); the dict of a ring slot is built once
        # and reused (the views alias the slot, the kernel refreshes the values)
        slot = max(self._stepper._slot, 0)
        log = self._log_cache.get(slot)
        if log is None:
            s = self._stepper.stats_ring[slot]
            log = {"Episode_Reward/" + k: s[i] for i, k in enumerate(self._term_names)}
            log[self._DIED_LOG_KEY] = s[native.STAT_NUM_TERMINATED_RESET]
            log["Episode_Termination/time_out"] = s[native.STAT_NUM_TIMEOUT_RESET]
            log.update(self._host_log)
            self._log_cache[slot] = log
        return log

    # ------------------------------------------------------------------ gym API
    def reset(self, seed: int | None = None, options=None):
        """``DirectRLEnv.reset``: _reset_idx(all) -> observations."""
        if seed is not None:
            self.seed(seed)
        self._stepper.reset_idx(None, self._out[self._out_i][2], self._out[self._out_i][3])
        # all envs reset: spread the episode counters (…env_v2.py:418-422), on the torch generator
        self.episode_length_buf = torch.randint_like(self._stepper.episode_length_buf, high=int(self.max_episode_length))
        self.extras["log"] = self._log_from_slot()
        obs = self._stepper.observe().clone()
        return {"policy": obs}, self.extras

    def step(self, actions: torch.Tensor):
        """(obs_dict, rew, terminated, truncated, extras) -- one fused kernel launch."""
        st = self._stepper
        self._out_i = (self._out_i + 1) % len(self._out)
        st.obs, st.rew, st.terminated, st.truncated = self._out[self._out_i]
        if self.capture_step_view:
            # off the fused fast path: the same kernel through the export hook, so host reward terms / callers of
            # `_robot.data`, `_contact_sensor.data` see this step's end-of-physics tensors
            if self._view is None:
                self._view = StepView(self)
            v = self._view
            v.actions_prev = st.state.get("actions")
            obs, rew, term, trunc = st.step(actions.to(self.device), export=v.export)
            v.actions_now = torch.tanh(actions.to(self.device))
            v.refresh()
            if self._host_terms:
                self._eval_host_terms(rew, term, trunc)
        else:
            obs, rew, term, trunc = st.step(actions.to(self.device))
        self.common_step_counter += 1
        self._sim_step_counter += self.cfg.decimation
        self.reset_terminated = term.view(torch.bool)
        self.reset_time_outs = trunc.view(torch.bool)
        if self._check_all_reset:
            # host sync, only for tiny env counts: the all-envs-reset RNG spread (…env_v2.py:418-422)
            if int(st.stats[native.STAT_NUM_RESET].item()) == self.num_envs:
                self.episode_length_buf = torch.randint_like(st.episode_length_buf, high=int(self.max_episode_length))
        self.extras["log"] = self._log_from_slot()
        return {"policy": obs}, rew, self.reset_terminated, self.reset_time_outs, self.extras

    def advance_host_curricula(self, steps: int) -> bool:
        """Host-side bookkeeping of `steps` control steps that ran WITHOUT `step()`'s Python (replay of a captured
        rollout graph, `rl/ppo_runner.py`): advances the global step counter and evaluates the task's host curricula.
        Returns True when kernel parameters changed (the caller must re-capture its graph).  v2 has no curriculum."""
        self.common_step_counter += int(steps)
        self._sim_step_counter += int(steps) * int(self.cfg.decimation)
        return False

    def alloc_host_buffers(self):
        """Pinned host buffers for ``step_host``: ``(actions (N,6) f32, rows (N,25) f32)``."""
        return (torch.zeros(self.num_envs, 6).pin_memory(),
                torch.zeros(self.num_envs, native.HOST_ROW_WORDS).pin_memory())

    def step_host(self, host_actions: torch.Tensor, host_rows: torch.Tensor):
        """``step`` for a HOST-resident consumer: pinned ``host_actions`` (N,6) in, the whole step result in the
        pinned ``host_rows`` ((N,25) f32: obs | reward | flags).  One C-ABI call (``zbot_step_host``) = one launch
        of the fused kernel reading / writing the pinned buffers directly over PCIe.  Synchronous.  Returns
        zero-copy host views ``(obs (N,23), rew (N,), terminated (N,) bool, truncated (N,) bool)`` into
        ``host_rows``; ``extras["log"]`` as in ``step``."""
        self._stepper.step_host(host_actions, host_rows)
        self.common_step_counter += 1
        self._sim_step_counter += self.cfg.decimation
        self.extras["log"] = self._log_from_slot()
        views = self._host_views.get(host_rows.data_ptr())
        if views is None:
            flags = host_rows.view(torch.uint8).view(self.num_envs, native.HOST_ROW_WORDS * 4)
            views = (host_rows[:, :23], host_rows[:, 23], flags[:, 96].view(torch.bool), flags[:, 97].view(torch.bool))
            self._host_views = {host_rows.data_ptr(): views}
        return views

    @property
    def last_step_packed(self) -> torch.Tensor:
        """uint8 view of the most recent step's outputs, laid out [obs | rew | terminated | truncated]."""
        return self._packed[self._out_i]

    def unpack_host(self, host_bytes: torch.Tensor):
        """Views (obs, rew, terminated, truncated) into a host copy of ``last_step_packed``."""
        n, no = self.num_envs, self._num_obs
        o1 = n * no * 4
        return (host_bytes[:o1].view(torch.float32).view(n, no), host_bytes[o1:o1 + 4 * n].view(torch.float32),
                host_bytes[o1 + 4 * n:o1 + 5 * n].view(torch.bool), host_bytes[o1 + 5 * n:o1 + 6 * n].view(torch.bool))

    def close(self):
        if getattr(self, "_stepper", None) is not None:
            self._stepper.close()

    def render(self, recompute: bool = False):
        return None
