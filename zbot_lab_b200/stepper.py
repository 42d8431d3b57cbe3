"""Torch-side owner of the device buffers + handle of the fused step (thin over the C ABI).

PyTorch is plumbing here: it allocates device memory and provides the current CUDA stream;
every computation of the step happens inside ``libzbot_b200.so``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import native
from .native import ZbotCfg, ZbotExport, ZbotMdpInputs

STATS_SLOTS = 64

#: (name, width) of every field of the 80-word fused state, see include/zbot_b200.h
STATE_FIELDS = {
    "root_pos": 3, "root_quat": 4, "root_lin_vel": 3, "root_ang_vel": 3, "joint_pos": 6, "joint_vel": 6,
    "p_delta": 6, "actions": 6, "carry_feet_fz": 2, "carry_mid_max": 1, "current_air_time": 2,
    "current_contact_time": 2, "last_air_time": 2, "last_contact_time": 2,
    "feet_contact_forces_last": 2, "feet_down_pos_last": 6, "feet_step_length": 2,
    "base_heading_x_sum": 1, "base_pos_y_err_sum": 1, "feet_force_sum": 1, "joint_speed_limit": 1,
    "episode_sums": 16,
}
MDP_STATE_FIELDS = {
    "p_delta": 6, "actions": 6, "feet_contact_forces_last": 2, "feet_down_pos_last": 6,
    "feet_step_length": 2, "base_heading_x_sum": 1, "base_pos_y_err_sum": 1, "feet_force_sum": 1,
    "joint_speed_limit": 1, "stale_base_pos": 3, "stale_forward": 3, "stale_feet_x": 6, "stale_feet_z": 6,
    "stale_feet_pos": 6, "stale_v_fwd": 1, "episode_sums": 16,
}


def _ptr(t: torch.Tensor | None):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _AoSoA:
    """``[NQ][N][4]`` float buffer with named strided field views."""

    def __init__(self, n, words, fields, word_fn, device):
        self.n = n
        self.buf = torch.zeros(words // 4, n, 4, dtype=torch.float32, device=device)
        self._fields = fields
        self._word = {k: word_fn(k.encode()) for k in fields}
        for k, w in self._word.items():
            if w < 0:
                raise RuntimeError(f"library does not know state field {k!r}")

    def get(self, name) -> torch.Tensor:
        """(N, width) copy of a field."""
        w0, width = self._word[name], self._fields[name]
        cols = [self.buf[(w0 + i) // 4, :, (w0 + i) % 4] for i in range(width)]
        return torch.stack(cols, dim=-1)

    def set(self, name, value, ids=None):
        w0, width = self._word[name], self._fields[name]
        value = torch.as_tensor(value, dtype=torch.float32, device=self.buf.device)
        if value.dim() == 0:
            value = value.expand(width)
        if value.dim() == 1 and value.shape[0] == width:
            value = value.unsqueeze(0)
        elif value.dim() == 1:
            value = value.unsqueeze(-1)
        sl = slice(None) if ids is None else ids
        for i in range(width):
            self.buf[(w0 + i) // 4, sl, (w0 + i) % 4] = value[..., i]

    def column(self, name, i=0) -> torch.Tensor:
        """strided (N,) VIEW of one word (writes go straight to the kernel's state)."""
        w = self._word[name] + i
        return self.buf[w // 4, :, w % 4]


#: column map of the snake task's (N, 41) export row (SnakeExport in csrc/zbot_core.h)
SNAKE_EXPORT = {"base_pos0": (0, 3), "base_quat0": (3, 7), "base_vel0": (7, 10), "base_pos1": (10, 13),
                "base_quat1": (13, 17), "base_vel1": (17, 20), "com_x1": (20, 22), "self_force1": (22, 23),
                "joint_pos1": (23, 29), "joint_vel1": (29, 35), "applied_torque1": (35, 41)}


class NativeStepper:
    """N envs of ``zbot-6b-walking-v2`` (or, with ``cfg.task = TASK_SNAKE_V0``, ``zbot-6s-snake-v0``) on one
    CUDA device."""

    def __init__(self, num_envs: int, device, cfg: ZbotCfg | None = None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("the zbot step has no CPU path: a CUDA device is required")
        self.lib = native.lib()
        self.n = int(num_envs)
        self.cfg = cfg if cfg is not None else native.make_cfg(self.n)
        if self.cfg.num_envs != self.n:
            raise ValueError("cfg.num_envs mismatch")
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.dev_index = dev_index
        self._h = C.c_void_p()
        with torch.cuda.device(dev_index):
            native.check(self.lib.zbot_create(C.byref(self.cfg), dev_index, C.byref(self._h)), "zbot_create")
        self.state = _AoSoA(self.n, native.STATE_WORDS, STATE_FIELDS, self.lib.zbot_state_word, self.device)
        self.episode_length_buf = torch.zeros(self.n, dtype=torch.int64, device=self.device)
        self.stats_ring = torch.zeros(STATS_SLOTS, native.STATS_WORDS, dtype=torch.float32, device=self.device)
        native.check(self.lib.zbot_bind(self._h, _ptr(self.state.buf), _ptr(self.episode_length_buf),
                                        _ptr(self.stats_ring), STATS_SLOTS), "zbot_bind")
        self.v4 = (self.cfg.task == native.TASK_WALKING_V4)
        self.mtask = (self.cfg.task == native.TASK_WALKING_M)
        self.num_obs = native.V4_NUM_OBS if self.v4 else native.M_NUM_OBS if self.mtask else native.NUM_OBS
        self.obs = torch.zeros(self.n, self.num_obs, dtype=torch.float32, device=self.device)
        self.rew = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.terminated = torch.zeros(self.n, dtype=torch.uint8, device=self.device)
        self.truncated = torch.zeros(self.n, dtype=torch.uint8, device=self.device)
        self._slot = -1
        # walking: joint_speed_limit = 1 (…env_v2.py:243); snake: per-env, set by the task class (snake_v0.py:121)
        self.state.set("joint_speed_limit", 3.14159265 if self.cfg.task == native.TASK_SNAKE_V0 else
                       float(self.cfg.contact_mu) if self.mtask else 1.0)      # manager task: per-env friction coefficient
        self.mdp_state = None
        self._host_ok: set = set()

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if self._h:
            self.lib.zbot_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream_arg(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def kernel_name(self) -> str:
        """The step-kernel instantiation this handle launches (as in an ncu launch list)."""
        return self.lib.zbot_step_kernel_name(self._h).decode()

    @property
    def launch_count(self) -> int:
        return int(self.lib.zbot_launch_count(self._h))

    def _next_slot(self):
        prev = self._slot
        self._slot = (self._slot + 1) % STATS_SLOTS
        return self._slot, prev

    @property
    def stats(self) -> torch.Tensor:
        """(32,) view of the statistics of the most recent step / reset."""
        return self.stats_ring[max(self._slot, 0)]

    # ------------------------------------------------------------------ fused step
    def step(self, actions: torch.Tensor, export=None, rand: torch.Tensor | None = None):
        """One control step, in place.  ``actions`` (N,6) float32 contiguous on the device."""
        if actions.dtype != torch.float32 or not actions.is_contiguous() or actions.device != self.obs.device:
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
        if actions.shape != (self.n, 6):
            raise ValueError(f"actions must be ({self.n}, 6), got {tuple(actions.shape)}")
        slot, prev = self._next_slot()
        if self.mtask:
            # zbot-6b-walking-m-v0: `rand` = (N,21) uniforms or None (in-kernel generator); `export` = (N,72) float32 (MExport)
            if rand is not None and (rand.dtype != torch.float32 or tuple(rand.shape) != (self.n, native.M_NUM_RAND)
                                     or not rand.is_contiguous() or rand.device != self.obs.device):
                raise ValueError(f"rand must be a contiguous float32 ({self.n}, {native.M_NUM_RAND}) tensor on the device")
            if export is None:
                rc = self.lib.zbot_m_step(self._h, _ptr(actions), _ptr(rand), _ptr(self.obs), _ptr(self.rew),
                                          _ptr(self.terminated), _ptr(self.truncated), slot, prev, _stream(self.device))
            else:
                rc = self.lib.zbot_m_step_export(self._h, _ptr(actions), _ptr(rand), _ptr(self.obs), _ptr(self.rew),
                                                 _ptr(self.terminated), _ptr(self.truncated), slot, prev, _ptr(export),
                                                 _stream(self.device))
            native.check(rc, "zbot_m_step")
            return self.obs, self.rew, self.terminated, self.truncated
        if self.v4:
            # zbot-6b-walking-v4: `rand` = (N,10) uniforms for this step or None (in-kernel generator);
            # `export` = one (N, 69) float32 tensor (V4Export)
            if rand is not None and (rand.dtype != torch.float32 or tuple(rand.shape) != (self.n, native.V4_NUM_RAND)
                                     or not rand.is_contiguous() or rand.device != self.obs.device):
                raise ValueError(f"rand must be a contiguous float32 ({self.n}, {native.V4_NUM_RAND}) tensor on the device")
            if export is None:
                rc = self.lib.zbot_v4_step(self._h, _ptr(actions), _ptr(rand), _ptr(self.obs), _ptr(self.rew),
                                           _ptr(self.terminated), _ptr(self.truncated), slot, prev, _stream(self.device))
            else:
                rc = self.lib.zbot_v4_step_export(self._h, _ptr(actions), _ptr(rand), _ptr(self.obs), _ptr(self.rew),
                                                  _ptr(self.terminated), _ptr(self.truncated), slot, prev, _ptr(export),
                                                  _stream(self.device))
            native.check(rc, "zbot_v4_step")
            return self.obs, self.rew, self.terminated, self.truncated
        if export is not None and self.cfg.task == native.TASK_SNAKE_V0:
            # snake task: `export` is one (N, 41) float32 tensor (include/zbot_b200.h: zbot_snake_step_export)
            rc = self.lib.zbot_snake_step_export(self._h, _ptr(actions), _ptr(self.obs), _ptr(self.rew),
                                                 _ptr(self.terminated), _ptr(self.truncated), slot, prev,
                                                 _ptr(export), _stream(self.device))
        elif export is None:
            rc = self.lib.zbot_step(self._h, _ptr(actions), _ptr(self.obs), _ptr(self.rew), _ptr(self.terminated),
                                    _ptr(self.truncated), slot, prev, _stream(self.device))
        else:
            ex = ZbotExport(*[export[name].data_ptr() for name, _ in ZbotExport._fields_])
            rc = self.lib.zbot_step_export(self._h, _ptr(actions), _ptr(self.obs), _ptr(self.rew),
                                           _ptr(self.terminated), _ptr(self.truncated), slot, prev, C.byref(ex),
                                           _stream(self.device))
        native.check(rc, "zbot_step")
        return self.obs, self.rew, self.terminated, self.truncated

    # ------------------------------------------------------------------ rollout halves
    def policy_act(self, pol: "native.ZbotPolicy", obs, obs_out, act, logp, value, mu, sigma, seed: int = 0):
        """Actor + Gaussian sample + log-prob + critic + rollout-buffer stores as one launch (``zbot_policy_act``).
        All tensors float32, contiguous, on this device; ``obs_out`` may be None."""
        for name, x, cols in (("obs", obs, pol.num_obs), ("act", act, pol.num_actions), ("mu", mu, pol.num_actions),
                              ("sigma", sigma, pol.num_actions), ("logp", logp, None), ("value", value, None),
                              ("obs_out", obs_out, pol.num_obs)):
            if x is None and name == "obs_out":
                continue
            want = (self.n,) if cols is None else (self.n, cols)
            if x.dtype != torch.float32 or tuple(x.shape) != want or not x.is_contiguous() or x.device != self.obs.device:
                raise ValueError(f"policy_act: {name} must be a contiguous float32 {want} tensor on {self.device}")
        native.check(self.lib.zbot_policy_act(self._h, C.byref(pol), _ptr(obs), _ptr(obs_out), _ptr(act), _ptr(logp),
                                              _ptr(value), _ptr(mu), _ptr(sigma), int(seed) & (2 ** 64 - 1),
                                              _stream(self.device)), "zbot_policy_act")

    def rollout_store(self, rew, terminated, truncated, value, gamma: float, rew_out, done_out):
        """rew_out = rew + gamma * value * truncated, done_out = float(terminated | truncated) (``zbot_rollout_store``)."""
        for name, x, dt in (("rew", rew, torch.float32), ("value", value, torch.float32), ("rew_out", rew_out, torch.float32),
                            ("done_out", done_out, torch.float32), ("terminated", terminated, None), ("truncated", truncated, None)):
            ok = x.dtype == dt if dt is not None else x.dtype in (torch.uint8, torch.bool)
            if not ok or tuple(x.shape) != (self.n,) or not x.is_contiguous() or x.device != self.obs.device:
                raise ValueError(f"rollout_store: {name} has the wrong dtype / shape / device")
        native.check(self.lib.zbot_rollout_store(self._h, _ptr(rew), _ptr(terminated), _ptr(truncated), _ptr(value),
                                                 float(gamma), _ptr(rew_out), _ptr(done_out), _stream(self.device)),
                     "zbot_rollout_store")

    def step_host(self, host_actions: torch.Tensor, host_rows: torch.Tensor):
        """One control step for HOST buffers (``zbot_step_host``): ``host_actions`` pinned (N,6) f32 in, one
        25-word row per env ``[obs 23 | reward | flags]`` into the pinned ``host_rows`` ((N,25) f32).  One kernel
        launch; actions and rows cross PCIe zero-copy from inside the kernel.  Synchronous: the host owns the
        result when this returns."""
        n = self.n
        pa, pr = host_actions.data_ptr(), host_rows.data_ptr()
        if (pa, pr) not in self._host_ok:     # validate a buffer pair once (is_pinned() is a driver query)
            if not (host_actions.is_pinned() and host_rows.is_pinned()):
                raise ValueError("step_host needs pinned host tensors (torch.Tensor.pin_memory())")
            if host_actions.dtype != torch.float32 or tuple(host_actions.shape) != (n, 6) or not host_actions.is_contiguous():
                raise ValueError(f"host_actions must be contiguous float32 ({n}, 6)")
            if host_rows.dtype != torch.float32 or tuple(host_rows.shape) != (n, native.HOST_ROW_WORDS) \
                    or not host_rows.is_contiguous():
                raise ValueError(f"host_rows must be contiguous float32 ({n}, {native.HOST_ROW_WORDS})")
            if len(self._host_ok) > 64:
                self._host_ok.clear()
            self._host_ok.add((pa, pr))
        slot, prev = self._next_slot()
        rc = self.lib.zbot_step_host(self._h, C.c_void_p(pa), C.c_void_p(pr), slot, prev, self._stream_arg())
        native.check(rc, "zbot_step_host")

    def alloc_export(self) -> dict:
        n, d = self.n, self.device
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=d)
        return {
            "body_link_pos_w0": z(n, 12, 3), "body_link_quat_w0": z(n, 12, 4), "body_com_lin_vel_w0": z(n, 12, 3),
            "body_link_pos_w1": z(n, 12, 3), "body_link_quat_w1": z(n, 12, 4), "body_com_lin_vel_w1": z(n, 12, 3),
            "joint_pos1": z(n, 6), "joint_vel1": z(n, 6), "applied_torque1": z(n, 6),
            "net_forces_w_history1": z(n, 5, 12, 3), "last_air_time1": z(n, 12), "current_contact_time1": z(n, 12),
        }

    def set_all_reset_spread(self, enable: bool):
        """Device-side `episode_length_buf[:] = randint(0, max_episode_length)` when every env reset in one step
        (…env_v2.py:418-422), decided and written by the statistics kernel: no host sync (include/zbot_b200.h)."""
        native.check(self.lib.zbot_set_all_reset_spread(self._h, 1 if enable else 0), "zbot_set_all_reset_spread")

    def bind_terrain(self, heights: torch.Tensor, x0: float, y0: float, cell: float, tile_origins: torch.Tensor, tile_size: float,
                     env_origins4: torch.Tensor, curriculum: bool):
        """Manager task on a generated height field (``zbot_bind_terrain``); the tensors must stay alive with the stepper."""
        for t, shape in ((heights, None), (tile_origins, None), (env_origins4, (self.n, 4))):
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device or (shape and tuple(t.shape) != shape):
                raise ValueError("bind_terrain: contiguous float32 device tensors (env_origins (N, 4))")
        self._terrain_keep = (heights, tile_origins, env_origins4)
        native.check(self.lib.zbot_bind_terrain(self._h, _ptr(heights), heights.shape[0], heights.shape[1], float(x0), float(y0),
                                                float(cell), _ptr(tile_origins), tile_origins.shape[0], tile_origins.shape[1],
                                                float(tile_size), _ptr(env_origins4), 1 if curriculum else 0), "zbot_bind_terrain")

    def update_cfg(self):
        """Push the (mutated) ``self.cfg`` reward weights / event parameters to the live handle (host curricula)."""
        native.check(self.lib.zbot_update_cfg(self._h, C.byref(self.cfg)), "zbot_update_cfg")

    def reset_idx_v4(self, env_ids: torch.Tensor | None = None, rand: torch.Tensor | None = None):
        """``_reset_idx`` of the v4 task for an explicit id list (construction / ``env.reset()``; NOT the per-step
        partial reset, which the step kernel does itself): reset_base pose randomisation, command resampling and
        the local state resets of …env_v4.py:59-136, 888-976, written into the kernel's state words."""
        from .assets import zbot_6s as Z
        dev, c = self.device, self.cfg
        ids = torch.arange(self.n, device=dev) if env_ids is None else env_ids.to(device=dev, dtype=torch.int64)
        k = ids.numel()
        if k == 0:
            return
        u = torch.rand(k, 6, device=dev) if rand is None else rand.to(dev)
        lo = torch.tensor(list(c.ev_pose_lo), device=dev)
        hi = torch.tensor(list(c.ev_pose_hi), device=dev)
        smp = u[:, :3] * (hi - lo) + lo
        pos = torch.tensor(Z.DEFAULT_ROOT_POS, dtype=torch.float32, device=dev).repeat(k, 1)
        pos[:, :2] += smp[:, :2]
        yaw = smp[:, 2]
        quat = torch.stack([torch.cos(yaw * 0.5), torch.zeros_like(yaw), torch.zeros_like(yaw), torch.sin(yaw * 0.5)], -1)
        z = lambda w: torch.zeros(k, w, device=dev)
        st = self.state
        st.set("root_pos", pos, ids)
        st.set("root_quat", quat, ids)
        for name, w in (("root_lin_vel", 3), ("root_ang_vel", 3), ("joint_vel", 6), ("p_delta", 6), ("actions", 6),
                        ("current_air_time", 2), ("current_contact_time", 2), ("last_air_time", 2), ("last_contact_time", 2),
                        ("feet_step_length", 2), ("feet_force_sum", 1), ("episode_sums", 16)):
            st.set(name, z(w), ids)
        st.set("joint_pos", torch.tensor(Z.DEFAULT_JOINT_POS, dtype=torch.float32, device=dev).repeat(k, 1), ids)
        st.set("feet_contact_forces_last", torch.full((k, 2), 15.0, device=dev), ids)
        # resample_commands (mode "reset")
        if c.ev_dual_sign:
            sign = (u[:, 3] < c.ev_prob_pos).float() * 2.0 - 1.0
            high = c.ev_vel_hi + c.ev_offset * (sign - 1.0)
            cmd0 = (u[:, 4] * (high - c.ev_vel_lo) + c.ev_vel_lo) * sign
        else:
            cmd0 = u[:, 4] * (c.ev_vel_hi - c.ev_vel_lo) + c.ev_vel_lo
        cmd1 = u[:, 5] * (c.ev_yaw_hi - c.ev_yaw_lo) + c.ev_yaw_lo
        wrapped = torch.remainder(yaw + cmd1 + torch.pi, 2 * torch.pi)
        target = torch.where((wrapped == 0) & (yaw + cmd1 > 0), torch.full_like(wrapped, torch.pi), wrapped - torch.pi)
        st.set("carry_feet_fz", torch.stack([cmd0, cmd1], -1), ids)       # commands
        st.set("carry_mid_max", target.unsqueeze(-1), ids)                # target_heading_yaw
        st.set("base_heading_x_sum", yaw.unsqueeze(-1), ids)              # current_yaw
        self.episode_length_buf[ids] = 0
        posl, _, _ = self.articulation_view()                             # post-reset feet link positions (env-local)
        st.set("feet_down_pos_last", posl[ids][:, [0, 11]].reshape(k, 6), ids)

    def reset_idx_m(self, env_ids: torch.Tensor | None = None, rand: torch.Tensor | None = None):
        """``ManagerBasedRLEnv._reset_idx`` of the manager task for an explicit id list (construction / ``env.reset()``;
        the per-step reset of done envs happens inside the step kernel): reset_base on the root link `base`
        (zbotlab_env_cfg.py:207-222), default joints, reset_my_data (mdp/rewards.py:37-43), command resample, written
        into the kernel's state words.  ``rand`` = (K, 8) uniforms: pose x / y / yaw, command time / vx / vy / wz / standing."""
        from .assets import zbot_6s as Z
        from .assets import zbot_6s_v2 as V
        dev, c = self.device, self.cfg
        ids = torch.arange(self.n, device=dev) if env_ids is None else env_ids.to(device=dev, dtype=torch.int64)
        k = ids.numel()
        if k == 0:
            return
        u = torch.rand(k, 8, device=dev) if rand is None else rand.to(dev)
        m = V.model_f32()
        lo = torch.tensor(list(c.ev_pose_lo), device=dev)
        hi = torch.tensor(list(c.ev_pose_hi), device=dev)
        smp = u[:, :3] * (hi - lo) + lo
        yaw = smp[:, 2]
        f32 = lambda a: torch.tensor(a, dtype=torch.float32, device=dev)
        base0, root0, q0 = f32(V.DEFAULT_ROOT_POS), f32(m.default_root_pos), f32(m.default_root_quat)
        rel = root0 - base0                                               # chain root (foot0 sole) relative to `base`
        cy, sy = torch.cos(yaw), torch.sin(yaw)
        pos = torch.stack([base0[0] + smp[:, 0] + cy * rel[0] - sy * rel[1],
                           base0[1] + smp[:, 1] + sy * rel[0] + cy * rel[1], root0[2].expand(k)], -1)
        ch, sh = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)              # (ch,0,0,sh) o q0
        quat = torch.stack([ch * q0[0] - sh * q0[3], ch * q0[1] - sh * q0[2], ch * q0[2] + sh * q0[1], ch * q0[3] + sh * q0[0]], -1)
        z = lambda w: torch.zeros(k, w, device=dev)
        st = self.state
        st.set("root_pos", pos, ids)
        st.set("root_quat", quat, ids)
        for name, w in (("root_lin_vel", 3), ("root_ang_vel", 3), ("joint_vel", 6), ("actions", 6),
                        ("current_air_time", 2), ("current_contact_time", 2), ("last_air_time", 2), ("last_contact_time", 2),
                        ("feet_step_length", 2), ("feet_contact_forces_last", 2), ("feet_force_sum", 1), ("episode_sums", 16)):
            st.set(name, z(w), ids)
        pd0 = st.get("p_delta")[ids]                # words 3 / 4 = terrain level / type of the env: they survive a reset
        pd0[:, :3] = 0.0
        pd0[:, 5] = 0.0
        st.set("p_delta", pd0, ids)
        st.set("joint_pos", f32(m.default_joint_pos).repeat(k, 1), ids)
        rng = lambda i: float(c.cmd_hi[i]) - float(c.cmd_lo[i])
        cmd = torch.stack([u[:, 4 + i] * rng(i) + float(c.cmd_lo[i]) for i in range(3)], -1)
        st.set("carry_feet_fz", cmd[:, :2], ids)
        st.set("carry_mid_max", cmd[:, 2:3], ids)
        st.set("base_heading_x_sum", (u[:, 7] <= c.cmd_rel_standing).float().unsqueeze(-1), ids)
        st.set("base_pos_y_err_sum", (u[:, 3] * (c.cmd_resample_hi - c.cmd_resample_lo) + c.cmd_resample_lo).unsqueeze(-1), ids)
        # heading command / push_robot interval timer live in the (otherwise unused) p_delta words 0..2; their uniforms are the
        # optional columns 8..10 of `rand`
        if c.cmd_heading or c.push_interval_hi > 0:
            ux = u[:, 8:11] if u.shape[1] >= 11 else torch.rand(k, 3, device=dev)
            pd = st.get("p_delta")[ids]
            if c.cmd_heading:
                pd[:, 0] = ux[:, 0] * (c.cmd_heading_hi - c.cmd_heading_lo) + c.cmd_heading_lo
                pd[:, 1] = (ux[:, 1] <= c.cmd_rel_heading).float()
            if c.push_interval_hi > 0:
                pd[:, 2] = ux[:, 2] * (c.push_interval_hi - c.push_interval_lo) + c.push_interval_lo
            st.set("p_delta", pd, ids)
        self.episode_length_buf[ids] = 0
        posl, _, _ = self.articulation_view()         # chain view: link 1 = joint3 location = foot0 LINK origin, 11 = foot1
        st.set("feet_down_pos_last", posl[ids][:, [1, 11]].reshape(k, 6), ids)

    def reset_idx(self, env_ids: torch.Tensor | None = None, terminated=None, truncated=None):
        if self.v4:
            return self.reset_idx_v4(env_ids)
        if self.mtask:
            return self.reset_idx_m(env_ids)
        slot, _ = self._next_slot()
        if env_ids is None:
            rc = self.lib.zbot_reset_idx(self._h, None, -1, _ptr(terminated), _ptr(truncated), slot,
                                         _stream(self.device))
        else:
            env_ids = env_ids.to(device=self.device, dtype=torch.int64).contiguous()
            if env_ids.numel() == 0:
                return
            rc = self.lib.zbot_reset_idx(self._h, _ptr(env_ids), env_ids.numel(), _ptr(terminated), _ptr(truncated),
                                         slot, _stream(self.device))
        native.check(rc, "zbot_reset_idx")

    def observe(self) -> torch.Tensor:
        if self.mtask:
            # ObservationManager.compute of the manager task from the state words (reset / query path only)
            from .assets import zbot_6s_v2 as V
            m = V.model_f32()
            _, quat, _ = self.articulation_view()
            g = self.state.get
            lq = torch.tensor(m.link_rot[V.link_index("base")], dtype=torch.float32, device=self.device)
            a, b = quat[:, 6], lq                                      # body 3 chain quaternion o (chain -> base link)
            rq = torch.stack([a[:, 0] * b[0] - a[:, 1] * b[1] - a[:, 2] * b[2] - a[:, 3] * b[3],
                              a[:, 0] * b[1] + a[:, 1] * b[0] + a[:, 2] * b[3] - a[:, 3] * b[2],
                              a[:, 0] * b[2] - a[:, 1] * b[3] + a[:, 2] * b[0] + a[:, 3] * b[1],
                              a[:, 0] * b[3] + a[:, 1] * b[2] - a[:, 2] * b[1] + a[:, 3] * b[0]], -1)
            il = list(V.CHAIN_TO_IL)
            q0 = torch.tensor(m.default_joint_pos, dtype=torch.float32, device=self.device)
            qrel, qd = torch.zeros(self.n, 6, device=self.device), torch.zeros(self.n, 6, device=self.device)
            qrel[:, il] = g("joint_pos") - q0
            qd[:, il] = g("joint_vel")
            self.obs.copy_(torch.cat([rq, g("carry_feet_fz"), g("carry_mid_max"), qrel, qd, g("actions")], dim=-1))
            return self.obs
        if self.v4:
            # `_get_observations` of the v4 task (…env_v4.py:828-851) from the state words -- reset / query path only
            # (the step kernel writes its own observation)
            from .assets import zbot_6s as Z
            _, quat, _ = self.articulation_view()
            g = self.state.get
            diff = g("carry_mid_max")[:, 0] - g("base_heading_x_sum")[:, 0]
            heading_err = torch.atan2(torch.sin(diff), torch.cos(diff))
            q0 = torch.tensor(Z.DEFAULT_JOINT_POS, dtype=torch.float32, device=self.device)
            self.obs.copy_(torch.cat([quat[:, 6], g("joint_pos") - q0, g("joint_vel"), g("actions"),
                                      g("carry_feet_fz")[:, 0:1], heading_err.unsqueeze(-1)], dim=-1))
            return self.obs
        native.check(self.lib.zbot_observe(self._h, _ptr(self.obs), _stream(self.device)), "zbot_observe")
        return self.obs

    def articulation_view(self):
        n, d = self.n, self.device
        pos = torch.empty(n, 12, 3, device=d)
        quat = torch.empty(n, 12, 4, device=d)
        vel = torch.empty(n, 12, 3, device=d)
        native.check(self.lib.zbot_articulation_view(self._h, _ptr(pos), _ptr(quat), _ptr(vel), _stream(self.device)),
                     "zbot_articulation_view")
        return pos, quat, vel

    def set_sim_state(self, st: dict, ids=None):
        for k in ("root_pos", "root_quat", "root_lin_vel", "root_ang_vel", "joint_pos", "joint_vel"):
            self.state.set(k, st[k], ids)

    # ------------------------------------------------------------------ MDP-only path
    def mdp_init(self):
        self.mdp_state = _AoSoA(self.n, native.MDP_STATE_WORDS, MDP_STATE_FIELDS, self.lib.zbot_mdp_state_word,
                                self.device)
        self.mdp_episode_length_buf = torch.zeros(self.n, dtype=torch.int64, device=self.device)
        self.mdp_stats_ring = torch.zeros(STATS_SLOTS, native.STATS_WORDS, dtype=torch.float32, device=self.device)
        self._mdp_slot = -1
        native.check(self.lib.zbot_mdp_bind(self._h, _ptr(self.mdp_state.buf), _ptr(self.mdp_episode_length_buf),
                                            _ptr(self.mdp_stats_ring), STATS_SLOTS), "zbot_mdp_bind")
        self.mdp_state.set("joint_speed_limit", 1.0)

    @staticmethod
    def _mdp_inputs(S: dict, origins: torch.Tensor, step: bool) -> ZbotMdpInputs:
        def p(k):
            t = S.get(k)
            if t is None:
                return None
            assert t.dtype == torch.float32 and t.is_contiguous()
            return t.data_ptr()
        return ZbotMdpInputs(p("body_link_pos_w"), p("body_link_quat_w"), p("body_com_lin_vel_w"), p("joint_pos"),
                             p("joint_vel"), p("applied_torque") if step else None,
                             p("net_forces_w_history") if step else None, p("last_air_time") if step else None,
                             origins.data_ptr())

    def mdp_observe(self, S: dict, origins: torch.Tensor) -> torch.Tensor:
        mi = self._mdp_inputs(S, origins, False)
        native.check(self.lib.zbot_mdp_observe(self._h, C.byref(mi), _ptr(self.obs), _stream(self.device)),
                     "zbot_mdp_observe")
        return self.obs

    def mdp_step(self, S: dict, origins: torch.Tensor, actions: torch.Tensor):
        mi = self._mdp_inputs(S, origins, True)
        prev = self._mdp_slot
        self._mdp_slot = (self._mdp_slot + 1) % STATS_SLOTS
        native.check(self.lib.zbot_mdp_step(self._h, C.byref(mi), _ptr(actions), _ptr(self.obs), _ptr(self.rew),
                                            _ptr(self.terminated), _ptr(self.truncated), self._mdp_slot, prev,
                                            _stream(self.device)), "zbot_mdp_step")
        return self.obs, self.rew, self.terminated, self.truncated
