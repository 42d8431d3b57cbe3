"""TEST DOUBLE for ``zbot_lab_b200.stepper.NativeStepper`` backed by the CPU port (oracle/): lets the
CPU test suite drive the host-side surface (env class, wrapper, PPO runner, the reference's train.py)
without a GPU.  Lives under tests/ only; the product never imports it."""
import numpy as np
import torch

from oracle import cpu_port
from zbot_lab_b200 import native
from zbot_lab_b200.stepper import STATE_FIELDS


class _State:
    def __init__(self, port):
        self.port = port

    def get(self, name):
        return torch.from_numpy(self.port.field(name, STATE_FIELDS[name]).copy())

    def set(self, name, value, ids=None):
        v = torch.as_tensor(value, dtype=torch.float32).numpy()
        f = self.port.field(name, STATE_FIELDS[name])
        if ids is None:
            f[:] = v
        else:
            f[np.asarray(ids)] = v

    def column(self, name, i=0):
        return torch.from_numpy(self.port.field(name, STATE_FIELDS[name])[:, i])


class FakeStepper:
    def __init__(self, num_envs, device, cfg=None):
        self.n = int(num_envs)
        self.device = torch.device("cpu")
        self.cfg = cfg if cfg is not None else native.make_cfg(self.n)
        self.port = cpu_port.PortEnv(self.n, np.float32, self.cfg)
        self.state = _State(self.port)
        self.episode_length_buf = torch.from_numpy(self.port.ep_len)       # shares memory with the port
        self.stats_ring = torch.zeros(64, 32)
        self._slot = -1
        self.obs = torch.zeros(self.n, 23)
        self.rew = torch.zeros(self.n)
        self.terminated = torch.zeros(self.n, dtype=torch.uint8)
        self.truncated = torch.zeros(self.n, dtype=torch.uint8)
        self.launch_count = 0

    @property
    def stats(self):
        return self.stats_ring[max(self._slot, 0)]

    def set_all_reset_spread(self, enable):
        self.spread_all_reset = bool(enable)

    def _write_stats(self, reset_mask, rs, term, trunc, rew):
        prev = self._slot
        self._slot = (self._slot + 1) % 64
        s = self.stats_ring[self._slot]
        s.zero_()
        k = int(reset_mask.sum())
        if k > 0:
            s[:16] = torch.from_numpy(rs[reset_mask].sum(0) / k / 20.0)
            s[17] = float((term & reset_mask).sum())
            s[18] = float((trunc & reset_mask).sum())
        elif prev >= 0:
            s[:19] = self.stats_ring[prev][:19]
        s[16] = k
        s[19] = float(rew.sum())
        s[20] = float(term.sum())
        s[21] = float(trunc.sum())

    def step(self, actions, export=None):
        obs, rew, term, trunc, rs, _ = self.port.step(actions.detach().cpu().numpy())
        self.obs.copy_(torch.from_numpy(obs))
        self.rew.copy_(torch.from_numpy(rew))
        self.terminated.copy_(torch.from_numpy(term.astype(np.uint8)))
        self.truncated.copy_(torch.from_numpy(trunc.astype(np.uint8)))
        self._write_stats(term | trunc, rs, term, trunc, rew)
        self.launch_count += 2
        return self.obs, self.rew, self.terminated, self.truncated

    def reset_idx(self, env_ids=None, terminated=None, truncated=None):
        ids = np.arange(self.n) if env_ids is None else np.asarray(env_ids)
        rs = np.zeros((self.n, 16), np.float32)
        rs[ids] = self.port.field("episode_sums", 16)[ids]
        keep_speed = self.port.field("joint_speed_limit", 1)[ids].copy()
        keep_fl = self.port.field("feet_contact_forces_last", 2)[ids].copy()
        keep_fsl = self.port.field("feet_step_length", 2)[ids].copy()
        fresh = cpu_port.PortEnv(len(ids), np.float32)
        self.port.state[ids] = fresh.state
        self.port.field("joint_speed_limit", 1)[ids] = keep_speed
        self.port.field("feet_contact_forces_last", 2)[ids] = keep_fl
        self.port.field("feet_step_length", 2)[ids] = keep_fsl
        self.port.ep_len[ids] = 0
        mask = np.zeros(self.n, bool)
        mask[ids] = True
        t = terminated.numpy().astype(bool) if terminated is not None else np.zeros(self.n, bool)
        u = truncated.numpy().astype(bool) if truncated is not None else np.zeros(self.n, bool)
        self._write_stats(mask, rs, t, u, np.zeros(self.n, np.float32))

    def observe(self):
        sim = np.concatenate([self.port.field(k, w) for k, w in (("root_pos", 3), ("root_quat", 4), ("root_lin_vel", 3),
                                                                  ("root_ang_vel", 3), ("joint_pos", 6), ("joint_vel", 6))], -1)
        _, quat, _ = cpu_port.link_view(np.ascontiguousarray(sim, np.float32))
        from zbot_lab_b200.assets import zbot_6s as Z
        obs = np.concatenate([quat[:, 6], self.port.field("joint_pos", 6) - np.asarray(Z.DEFAULT_JOINT_POS, np.float32),
                              self.port.field("joint_vel", 6), self.port.field("actions", 6),
                              self.port.field("joint_speed_limit", 1)], -1)
        self.obs.copy_(torch.from_numpy(obs.astype(np.float32)))
        return self.obs

    def articulation_view(self):
        sim = np.concatenate([self.port.field(k, w) for k, w in (("root_pos", 3), ("root_quat", 4), ("root_lin_vel", 3),
                                                                  ("root_ang_vel", 3), ("joint_pos", 6), ("joint_vel", 6))], -1)
        p, q, v = cpu_port.link_view(np.ascontiguousarray(sim, np.float32))
        return torch.from_numpy(p), torch.from_numpy(q), torch.from_numpy(v)

    def close(self):
        pass


class FakeMStepper(FakeStepper):
    """CPU double of the manager task's stepper (``zbot-6b-walking-m-v0``): the kernel's arithmetic compiled for the host
    behind the same methods; ``reset_idx_m`` / ``observe`` are the shipped host code of ``NativeStepper`` itself."""

    def __init__(self, num_envs, device, cfg=None):
        assert cfg is not None and cfg.task == native.TASK_WALKING_M
        super().__init__(num_envs, device, cfg)
        self.mtask, self.v4 = True, False
        self.num_obs = native.M_NUM_OBS
        self.obs = torch.zeros(self.n, native.M_NUM_OBS)
        self._rng = np.random.default_rng(int(cfg.rng_seed) + 1)

    from zbot_lab_b200.stepper import NativeStepper as _NS
    reset_idx_m = _NS.reset_idx_m
    observe = _NS.observe
    del _NS

    def reset_idx(self, env_ids=None, terminated=None, truncated=None):
        return self.reset_idx_m(env_ids)

    def bind_terrain(self, heights, x0, y0, cell, tile_origins, tile_size, env_origins4, curriculum):
        h, to, eo = heights.numpy(), tile_origins.numpy(), env_origins4.numpy()      # CPU tensors: numpy shares their memory
        pt = cpu_port.PortTerrain(h.ctypes.data, h.shape[0], h.shape[1], float(x0), float(y0), float(cell), to.ctypes.data,
                                  to.shape[0], to.shape[1], float(tile_size), eo.ctypes.data, int(bool(curriculum)))
        pt._keep = (heights, tile_origins, env_origins4, h, to, eo)
        self.port.terrain = pt

    def update_cfg(self):
        pass          # the port reads self.cfg on every step

    def step(self, actions, export=None, rand=None):
        rnd = self._rng.random((self.n, native.M_NUM_RAND)).astype(np.float32) if rand is None else rand.numpy()
        obs, rew, term, trunc, rs, _ = self.port.step(actions.detach().cpu().numpy(), rnd=rnd)
        if self.cfg.obs_noise_enable:
            lo, hi = np.zeros(25, np.float32), np.zeros(25, np.float32)
            lo[:24], hi[:24] = list(self.cfg.obs_noise_lo), list(self.cfg.obs_noise_hi)
            obs = obs + (self._rng.random(obs.shape).astype(np.float32) * (hi - lo) + lo)
        self.obs.copy_(torch.from_numpy(obs))
        self.rew.copy_(torch.from_numpy(rew))
        self.terminated.copy_(torch.from_numpy(term.astype(np.uint8)))
        self.truncated.copy_(torch.from_numpy(trunc.astype(np.uint8)))
        mask = term | trunc
        prev = self._slot
        self._write_stats(mask, rs[:, :16], term, trunc, rew)
        s = self.stats_ring[self._slot]
        if mask.any():
            if self.cfg.num_terms <= 13:                 # raw-count tail slots (include/zbot_b200.h: zbot_m_step)
                s[14], s[15] = float(rs[mask][:, 14].sum()), float(rs[mask][:, 15].sum())
            # statistics words 22..25: is_terminated's episodic sum (normalised like a term) and the per-DoneTerm counts
            s[22] = float(rs[mask][:, 16].sum()) / int(mask.sum()) / 20.0
            s[23], s[24], s[25] = (float(rs[mask][:, 17 + i].sum()) for i in range(3))
        elif prev >= 0:
            s[22:26] = self.stats_ring[prev][22:26]
        self.launch_count += 2
        return self.obs, self.rew, self.terminated, self.truncated
