"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the Isaac Lab semantics the
``zbot-6b-walking-v2`` step depends on but which are NOT in /root/reference
(un-vendored ``isaaclab`` 2.x; see SURVEY.md Appendix B, "[IL-upstream]").

PARITY UNPINNED for everything in this file: the reference holds no test or golden
vector for these behaviours; they restate upstream Isaac Lab as described in
SURVEY.md Appendix B.  The call sites that rely on them are cited per function.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may
import this module.
"""
from __future__ import annotations

import math

import numpy as np


def env_origins_grid(num_envs: int, spacing: float = 4.0) -> np.ndarray:
    """``TerrainImporter`` plane-terrain env origins (consumed at
    ``…env_v2.py:406, 430, 494``).  float32 (N,3)."""
    n = int(num_envs)
    num_rows = int(np.ceil(n / int(np.sqrt(n))))
    num_cols = int(np.ceil(n / num_rows))
    ii, jj = np.meshgrid(np.arange(num_rows), np.arange(num_cols), indexing="ij")
    org = np.zeros((num_rows * num_cols, 3), dtype=np.float32)
    org[:, 0] = (-(ii.flatten().astype(np.float32) - np.float32((num_rows - 1) / 2)) * np.float32(spacing))
    org[:, 1] = ((jj.flatten().astype(np.float32) - np.float32((num_cols - 1) / 2)) * np.float32(spacing))
    return org[:n].copy()


def quat_apply(q: np.ndarray, v: np.ndarray) -> np.ndarray:
    """``isaaclab.utils.math.quat_apply`` (``…env_v2.py:322,344-345``):
    wxyz, ``t = 2 (q_xyz x v); v' = v + q_w t + q_xyz x t`` -- this association order."""
    xyz = q[..., 1:]
    t = np.cross(xyz, v) * q.dtype.type(2)
    return v + q[..., 0:1] * t + np.cross(xyz, t)


def implicit_actuator_applied_torque(q_target, q, qd, kp, kd, effort_limit):
    """``ImplicitActuator.compute`` bookkeeping (read at ``…env_v2.py:560``):
    ``clip(kp (q*-q) + kd (0-qd), +-effort_limit)``."""
    return np.clip(kp * (q_target - q) + kd * (0.0 - qd), -effort_limit, effort_limit)


class ContactSensorState:
    """``ContactSensor`` buffers for B bodies (cfg ``…env_v2.py:30-36``:
    history_length=5, update_period=0, track_air_time=True, force_threshold=1.0)."""

    def __init__(self, n, num_bodies, history=5, dtype=np.float32, threshold=1.0):
        self.dtype = dtype
        self.threshold = dtype(threshold)
        self.net_forces_w = np.zeros((n, num_bodies, 3), dtype)
        self.net_forces_w_history = np.zeros((n, history, num_bodies, 3), dtype)
        self.current_air_time = np.zeros((n, num_bodies), dtype)
        self.current_contact_time = np.zeros((n, num_bodies), dtype)
        self.last_air_time = np.zeros((n, num_bodies), dtype)
        self.last_contact_time = np.zeros((n, num_bodies), dtype)

    def update(self, net_forces_w: np.ndarray, dt: float):
        dt = self.dtype(dt)
        self.net_forces_w = net_forces_w.astype(self.dtype)
        self.net_forces_w_history = np.roll(self.net_forces_w_history, 1, axis=1)
        self.net_forces_w_history[:, 0] = self.net_forces_w
        f = self.net_forces_w
        norm = np.sqrt(f[..., 0] * f[..., 0] + f[..., 1] * f[..., 1] + f[..., 2] * f[..., 2])
        is_contact = norm > self.threshold
        first_contact = (self.current_air_time > 0) & is_contact
        first_detached = (self.current_contact_time > 0) & ~is_contact
        self.last_air_time = np.where(first_contact, self.current_air_time + dt, self.last_air_time)
        self.current_air_time = np.where(~is_contact, self.current_air_time + dt, self.dtype(0))
        self.last_contact_time = np.where(first_detached, self.current_contact_time + dt,
                                          self.last_contact_time)
        self.current_contact_time = np.where(is_contact, self.current_contact_time + dt, self.dtype(0))

    def reset(self, ids):
        for a in (self.net_forces_w, self.net_forces_w_history, self.current_air_time,
                  self.current_contact_time, self.last_air_time, self.last_contact_time):
            a[ids] = 0


# DirectRLEnv constants (cfg ``…env_v2.py:39-48``)
def max_episode_length(episode_length_s=20.0, sim_dt=1 / 200.0, decimation=4) -> int:
    return math.ceil(episode_length_s / (sim_dt * decimation))
