"""Deterministic synthetic inputs for the ``zbot-6b-walking-v2`` step.

Used by the parity tests, ``tests/golden/make_golden.py`` and ``bench.py`` so that
every consumer sees the same state for a given seed (SURVEY.md §8(d) "Synthetic inputs").
Pure numpy; nothing here touches the GPU.
"""
from __future__ import annotations

import numpy as np

from ..assets import zbot_6s as Z

F = np.float32


def index_sets() -> dict:
    """The four index lists of ``…env_v2.py:227-230`` resolved BY NAME:
    sensor order for the contact sets, articulation order for the pose sets."""
    feet_ids, _ = Z.find_bodies("foot.*", Z.SENSOR_BODY_NAMES)
    undesired, _ = Z.find_bodies("base|a.*|b.*", Z.SENSOR_BODY_NAMES)
    base_idx, _ = Z.find_bodies("base", Z.LINK_NAMES)
    feet_idx, _ = Z.find_bodies("foot.*", Z.LINK_NAMES)
    return {"feet_ids": feet_ids, "undesired_ids": undesired, "base_body_idx": base_idx,
            "feet_body_idx": feet_idx}


def reset_tables() -> dict:
    """Default-pose link poses relative to the env origin (constant rows written on reset)."""
    p, q = Z.default_link_poses()
    return {"body_link_pos_local": p.astype(F), "body_link_quat": q.astype(F)}


def env_origins_grid(num_envs: int, spacing: float = 4.0) -> np.ndarray:
    """Plane-terrain env origins (Isaac Lab grid).  float32 (N,3)."""
    n = int(num_envs)
    num_rows = int(np.ceil(n / int(np.sqrt(n))))
    num_cols = int(np.ceil(n / num_rows))
    ii, jj = np.meshgrid(np.arange(num_rows), np.arange(num_cols), indexing="ij")
    org = np.zeros((num_rows * num_cols, 3), dtype=F)
    org[:, 0] = -(ii.flatten().astype(F) - F((num_rows - 1) / 2)) * F(spacing)
    org[:, 1] = (jj.flatten().astype(F) - F((num_cols - 1) / 2)) * F(spacing)
    return org[:n].copy()


def synth_articulation_state(rng: np.random.Generator, n: int, origins: np.ndarray,
                             die_frac: float = 0.08) -> dict:
    """Synthetic end-of-physics articulation + contact-sensor state in the reference's
    tensor layouts (``robot.data`` / ``contact_sensor.data``)."""
    ids = index_sets()
    pos = rng.normal(0.0, 0.12, (n, 12, 3)).astype(F)
    pos[..., 2] += F(0.30)
    base = ids["base_body_idx"][0]
    pos[:, base, 2] = F(0.24) + np.abs(rng.normal(0, 0.03, n)).astype(F)
    low = rng.random(n) < die_frac * 0.4
    pos[low, base, 2] = F(0.2) - rng.random(int(low.sum())).astype(F) * F(0.05)
    far = rng.random(n) < die_frac * 0.2
    pos[far, base, 1] += F(0.7)
    pos = pos + origins[:, None, :]
    quat = rng.normal(size=(n, 12, 4)).astype(F)
    quat /= np.linalg.norm(quat, axis=-1, keepdims=True)
    vel = rng.normal(0.0, 0.3, (n, 12, 3)).astype(F)
    jp = (np.asarray(Z.DEFAULT_JOINT_POS, F) + rng.uniform(-0.2, 0.2, (n, 6)).astype(F)).astype(F)
    jv = rng.normal(0.0, 0.5, (n, 6)).astype(F)
    tau = rng.normal(0.0, 5.0, (n, 6)).astype(F)
    hist = np.abs(rng.normal(size=(n, 5, 12, 3))).astype(F) * np.array([0.3, 0.3, 8.0], F)
    scale = np.full(12, 0.02, F)
    scale[ids["feet_ids"]] = 1.0
    hist = hist * scale[None, None, :, None]
    # some feet in the air (forces below the 1 N / 10 N thresholds)
    air = rng.random((n, 2)) < 0.3
    for j, b in enumerate(ids["feet_ids"]):
        hist[air[:, j], :, b, :] *= F(0.05)
    hit = rng.random(n) < die_frac * 0.4
    which = rng.integers(0, len(ids["undesired_ids"]), n)
    for e in np.nonzero(hit)[0]:
        hist[e, rng.integers(0, 5), ids["undesired_ids"][which[e]], :] = np.array([0.5, -0.4, 2.0], F)
    return {
        "body_link_pos_w": pos, "body_link_quat_w": quat.astype(F), "body_com_lin_vel_w": vel,
        "joint_pos": jp, "joint_vel": jv, "applied_torque": tau,
        "net_forces_w_history": hist.astype(F),
        "last_air_time": rng.random((n, 12)).astype(F),
        "current_contact_time": rng.random((n, 12)).astype(F),
    }


def synth_mdp_case(seed: int, n: int, steps: int, die_frac: float = 0.08) -> dict:
    """A whole MDP-only parity case: initial state S0, per-step (actions, S1), initial
    episode counters (some close to the 999-step truncation)."""
    rng = np.random.default_rng(seed)
    origins = env_origins_grid(n)
    ep = rng.integers(0, 1000, n).astype(np.int64)
    near = rng.random(n) < 0.1
    ep[near] = 999 - rng.integers(1, max(2, steps), int(near.sum()))
    case = {"origins": origins, "episode_length_buf0": ep,
            "S0": synth_articulation_state(rng, n, origins, die_frac=0.0), "steps": []}
    for _ in range(steps):
        a = rng.normal(0.0, 1.0, (n, 6)).astype(F)
        case["steps"].append((a, synth_articulation_state(rng, n, origins, die_frac)))
    return case


def synth_sim_state(rng: np.random.Generator, n: int) -> dict:
    """Perturbed start state for the FULL fused step: default pose with
    ``q += U(-0.2,0.2)``, ``qd = N(0,0.5)``, root ``z += U(0,0.02)``, root lin vel ``N(0,0.1)``.
    Root position is env-LOCAL (relative to the env origin)."""
    q = np.asarray(Z.DEFAULT_JOINT_POS, F) + rng.uniform(-0.2, 0.2, (n, 6)).astype(F)
    qd = rng.normal(0, 0.5, (n, 6)).astype(F)
    root_pos = np.tile(np.asarray(Z.DEFAULT_ROOT_POS, F), (n, 1))
    root_pos[:, 2] += rng.uniform(0.0, 0.02, n).astype(F)
    root_quat = np.tile(np.asarray(Z.DEFAULT_ROOT_QUAT, F), (n, 1))
    root_lin = rng.normal(0, 0.1, (n, 3)).astype(F)
    root_ang = np.zeros((n, 3), F)
    return {"root_pos": root_pos, "root_quat": root_quat, "root_lin_vel": root_lin,
            "root_ang_vel": root_ang, "joint_pos": q.astype(F), "joint_vel": qd}


# ------------------------------------------------------------------------------------------------ snake task
SNAKE_SENSOR_WIDTHS = (5, 4, 3, 2)   # filter bodies of the four self-contact sensors (snake_v0.py:23-48)


def snake_reset_tables() -> dict:
    """Default-pose link poses / CoM positions of the snake robot, relative to the env origin."""
    from ..assets import zbot_d_6s as S

    m = S.model_f32()
    p, q = Z.fk_links(m.default_root_pos, m.default_root_quat, m.default_joint_pos, m)
    com = p + Z.quat_rotate(q, m.link_com)
    return {"body_link_pos_local": p.astype(F), "body_link_quat": q.astype(F), "body_com_pos_local": com.astype(F)}


def synth_snake_state(rng: np.random.Generator, n: int, origins: np.ndarray, die_frac: float = 0.06) -> dict:
    """Synthetic end-of-physics state in the tensor layouts the reference snake task reads."""
    t = snake_reset_tables()
    pos = t["body_link_pos_local"][None] + rng.normal(0, 0.03, (n, 12, 3)).astype(F)
    far = rng.random(n) < die_frac * 0.5
    pos[far, 6, 0] += F(0.3)
    pos = (pos + origins[:, None, :]).astype(F)
    com = (t["body_com_pos_local"][None] + rng.normal(0, 0.03, (n, 12, 3)).astype(F) + origins[:, None, :]).astype(F)
    quat = rng.normal(size=(n, 12, 4)).astype(F)
    quat /= np.linalg.norm(quat, axis=-1, keepdims=True)
    out = {
        "body_link_pos_w": pos, "body_link_quat_w": quat.astype(F), "body_com_pos_w": com,
        "body_link_vel_w": rng.normal(0, 0.3, (n, 12, 6)).astype(F),
        "joint_pos": rng.uniform(-1.0, 1.0, (n, 6)).astype(F), "joint_vel": rng.normal(0, 0.5, (n, 6)).astype(F),
        "applied_torque": rng.normal(0, 3.0, (n, 6)).astype(F),
    }
    hit = rng.random(n) < die_frac * 0.5
    for i, w in enumerate(SNAKE_SENSOR_WIDTHS, start=1):
        fm = (np.abs(rng.normal(size=(n, 1, w, 3))) * 0.1).astype(F)
        sel = hit & (rng.integers(0, 4, n) == i - 1)
        fm[sel, 0, rng.integers(0, w), :] = np.array([0.9, -0.8, 0.7], F)
        out[f"force_matrix_w_{i}"] = fm
    return out


def synth_snake_case(seed: int, n: int, steps: int) -> dict:
    rng = np.random.default_rng(seed)
    origins = env_origins_grid(n)
    ep = rng.integers(0, 800, n).astype(np.int64)
    near = rng.random(n) < 0.1
    ep[near] = 799 - rng.integers(1, max(2, steps), int(near.sum()))
    speed = ((rng.random(n) * 1.8 + 0.2) * np.pi).astype(F)          # snake_v0.py:121
    case = {"origins": origins, "episode_length_buf0": ep, "joint_speed_limit": speed,
            "S0": synth_snake_state(rng, n, origins, 0.0), "steps": []}
    for _ in range(steps):
        case["steps"].append((rng.normal(0, 1, (n, 6)).astype(F), synth_snake_state(rng, n, origins)))
    return case


# ------------------------------------------------------------------------------------------------ walking v4
def synth_v4_state(rng: np.random.Generator, n: int, origins: np.ndarray, die_frac: float = 0.06) -> dict:
    """Synthetic end-of-physics state in the tensor layouts ``Zbot6SEnvV4`` reads (3-deep force history, all four
    contact timers, link AND CoM velocities, joint accelerations)."""
    S = synth_articulation_state(rng, n, origins, die_frac=0.0)
    ids = index_sets()
    base = ids["base_body_idx"][0]
    low = rng.random(n) < die_frac * 0.5
    S["body_link_pos_w"][low, base, 2] = F(0.19) - rng.random(int(low.sum())).astype(F) * F(0.05)
    # feet close together for some envs (feet_close term)
    close = rng.random(n) < 0.3
    f0, f1 = ids["feet_body_idx"]
    S["body_link_pos_w"][close, f1, :2] = S["body_link_pos_w"][close, f0, :2] + rng.normal(0, 0.05, (int(close.sum()), 2)).astype(F)
    hist = S.pop("net_forces_w_history")[:, :3].copy()
    hist[:, :, ids["undesired_ids"], :] *= F(0.5)
    hit = rng.random(n) < die_frac * 0.5
    which = rng.integers(0, len(ids["undesired_ids"]), n)
    for e in np.nonzero(hit)[0]:
        hist[e, rng.integers(0, 3), ids["undesired_ids"][which[e]], :] = np.array([0.3, -0.3, 0.4], F)
    S["net_forces_w_history"] = hist
    S["body_link_lin_vel_w"] = rng.normal(0.0, 0.3, (n, 12, 3)).astype(F)
    S["joint_acc"] = rng.normal(0.0, 30.0, (n, 6)).astype(F)
    S["last_contact_time"] = rng.random((n, 12)).astype(F)
    S["current_air_time"] = (rng.random((n, 12)) * (rng.random((n, 12)) < 0.5)).astype(F)
    S["current_contact_time"] = np.where(S["current_air_time"] > 0, F(0), rng.random((n, 12)).astype(F)).astype(F)
    return S


def synth_v4_case(seed: int, n: int, steps: int) -> dict:
    rng = np.random.default_rng(seed)
    origins = env_origins_grid(n)
    ep = rng.integers(0, 1000, n).astype(np.int64)
    near = rng.random(n) < 0.1
    ep[near] = 999 - rng.integers(1, max(2, steps), int(near.sum()))
    time_left = rng.uniform(3.0, 6.0, n).astype(F)
    soon = rng.random(n) < 0.25
    time_left[soon] = (rng.integers(1, max(2, steps + 1), int(soon.sum())) * 0.02 - 0.01).astype(F)   # fire within the case
    case = {"origins": origins, "episode_length_buf0": ep, "interval_time_left0": time_left,
            "commands0": np.stack([rng.uniform(-0.3, 0.3, n), rng.uniform(-0.1, 0.1, n)], -1).astype(F),
            "target_heading_yaw0": rng.uniform(-3.0, 3.0, n).astype(F),
            "S0": synth_v4_state(rng, n, origins, 0.0), "steps": []}
    for _ in range(steps):
        case["steps"].append((rng.normal(0, 1, (n, 6)).astype(F), synth_v4_state(rng, n, origins),
                              rng.random((n, 10)).astype(F)))
    return case
