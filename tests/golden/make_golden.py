"""Generates ``tests/golden/mdp_v2_*.npz`` by running the reference's OWN code
(``/root/reference/.../zbot_direct_6dof_bipedal_env_v2.py``, unmodified, loaded through
``oracle/ref_loader.py``) on deterministic synthetic articulation state.

Build-container only (needs /root/reference).  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_harness import RefMdpHarness  # noqa: E402
from oracle import ref_loader  # noqa: E402
from zbot_lab_b200.assets import zbot_6s as Z  # noqa: E402
from zbot_lab_b200.utils import synthetic as syn  # noqa: E402

CASES = [  # (name, seed, num_envs, steps)
    ("mdp_v2_n64", 0, 64, 6),      # BASELINE.json configs[0]: 64 envs, CPU torch
    ("mdp_v2_n7", 1, 7, 4),        # ragged (not a multiple of any tile size)
    ("mdp_v2_n300", 2, 300, 3),
]


def run_case(seed, n, steps):
    torch.set_num_threads(1)
    case = syn.synth_mdp_case(seed, n, steps)
    dj = torch.tensor(Z.DEFAULT_JOINT_POS, dtype=torch.float32).repeat(n, 1)
    drs = torch.zeros(n, 13)
    drs[:, :3] = torch.tensor(Z.DEFAULT_ROOT_POS)
    drs[:, 3] = 1.0
    h = RefMdpHarness(n, torch.from_numpy(case["origins"]), syn.reset_tables(), syn.index_sets(), dj, drs)
    h.env.episode_length_buf[:] = torch.from_numpy(case["episode_length_buf0"])
    out = {"seed": seed, "n": n, "steps": steps}
    h.attach(case["S0"])
    out["obs0"] = h.observe().numpy()
    for t, (a, S1) in enumerate(case["steps"]):
        obs, rew, term, trunc, ids, log = h.step(torch.from_numpy(a), S1)
        out[f"obs{t + 1}"] = obs.numpy()
        out[f"rew{t + 1}"] = rew.numpy()
        out[f"terminated{t + 1}"] = term.numpy()
        out[f"truncated{t + 1}"] = trunc.numpy()
        out[f"reset_ids{t + 1}"] = ids.numpy()
        if log is not None:
            for k, v in log.items():
                out[f"log{t + 1}/{k}"] = np.float32(v)
        for k, v in h.mdp_state().items():
            out[f"state{t + 1}/{k}"] = v.numpy()
    return out


SNAKE_CASES = [("snake_v0_n64", 10, 64, 5), ("snake_v0_n9", 11, 9, 4)]


def run_snake_case(seed, n, steps):
    """BASELINE.json configs[3] (snake task) -- the MDP of the reference's own ZbotDirectEnvV0."""
    from oracle.ref_harness import RefSnakeHarness
    from zbot_lab_b200.assets import zbot_d_6s as S
    torch.set_num_threads(1)
    case = syn.synth_snake_case(seed, n, steps)
    m = S.model_f32()
    drs = torch.zeros(n, 13)
    drs[:, :3] = torch.tensor(m.default_root_pos, dtype=torch.float32)
    drs[:, 3:7] = torch.tensor(m.default_root_quat, dtype=torch.float32)
    h = RefSnakeHarness(n, torch.from_numpy(case["origins"]), syn.snake_reset_tables(), drs,
                        torch.from_numpy(case["joint_speed_limit"]))
    h.env.episode_length_buf[:] = torch.from_numpy(case["episode_length_buf0"])
    out = {"seed": seed, "n": n, "steps": steps}
    h.attach(case["S0"])
    out["obs0"] = h.observe().numpy()
    for t, (a, S1) in enumerate(case["steps"]):
        obs, rew, term, trunc, ids, log = h.step(torch.from_numpy(a), S1)
        out[f"obs{t + 1}"], out[f"rew{t + 1}"] = obs.numpy(), rew.numpy()
        out[f"terminated{t + 1}"], out[f"truncated{t + 1}"] = term.numpy(), trunc.numpy()
        out[f"reset_ids{t + 1}"] = ids.numpy()
        if log is not None:
            for k, v in log.items():
                out[f"log{t + 1}/{k}"] = np.float32(v)
        for k, v in h.mdp_state().items():
            out[f"state{t + 1}/{k}"] = v.numpy()
    return out


V4_CASES = [("v4_n64", 20, 64, 6), ("v4_n11", 21, 11, 5)]


def run_v4_case(seed, n, steps):
    """zbot-6b-walking-v4: the MDP + events of the reference's own Zbot6SEnvV4."""
    from oracle.ref_harness import RefV4Harness
    torch.set_num_threads(1)
    case = syn.synth_v4_case(seed, n, steps)
    dj = torch.tensor(Z.DEFAULT_JOINT_POS, dtype=torch.float32).repeat(n, 1)
    drs = torch.zeros(n, 13)
    drs[:, :3] = torch.tensor(Z.DEFAULT_ROOT_POS)
    drs[:, 3] = 1.0
    h = RefV4Harness(n, torch.from_numpy(case["origins"]), syn.index_sets(), dj, drs, case["interval_time_left0"])
    h.env.episode_length_buf[:] = torch.from_numpy(case["episode_length_buf0"])
    h.env.commands[:] = torch.from_numpy(case["commands0"])
    h.env.target_heading_yaw[:] = torch.from_numpy(case["target_heading_yaw0"])
    out = {"seed": seed, "n": n, "steps": steps}
    h.attach(case["S0"])
    out["obs0"] = h.observe().numpy()
    for t, (a, S1, rnd) in enumerate(case["steps"]):
        obs, rew, term, trunc, ids, iv_ids, log = h.step(torch.from_numpy(a), S1, rnd)
        out[f"obs{t + 1}"], out[f"rew{t + 1}"] = obs.numpy(), rew.numpy()
        out[f"terminated{t + 1}"], out[f"truncated{t + 1}"] = term.numpy(), trunc.numpy()
        out[f"reset_ids{t + 1}"] = ids.numpy()
        out[f"interval_ids{t + 1}"] = iv_ids.numpy()
        if log is not None:
            for k, v in log.items():
                out[f"log{t + 1}/{k}"] = np.float32(v)
        for k, v in h.mdp_state().items():
            out[f"state{t + 1}/{k}"] = v.numpy()
    return out


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for name, seed, n, steps in V4_CASES:
        out = run_v4_case(seed, n, steps)
        np.savez_compressed(os.path.join(here, name + ".npz"), **out)
        print(name, "resets:", sum(len(out[f"reset_ids{t + 1}"]) for t in range(steps)),
              "terminated:", sum(int(out[f"terminated{t + 1}"].sum()) for t in range(steps)),
              "interval resamples:", sum(len(out[f"interval_ids{t + 1}"]) for t in range(steps)))
    sc = ref_loader.reference_v4_reward_scales()
    np.savez(os.path.join(here, "reward_scales_v4.npz"), names=np.array(list(sc.keys())),
             values=np.array(list(sc.values()), dtype=np.float64))
    for name, seed, n, steps in SNAKE_CASES:
        out = run_snake_case(seed, n, steps)
        np.savez_compressed(os.path.join(here, name + ".npz"), **out)
        print(name, "resets:", sum(len(out[f"reset_ids{t + 1}"]) for t in range(steps)),
              "terminated:", sum(int(out[f"terminated{t + 1}"].sum()) for t in range(steps)))
    sc = ref_loader.reference_snake_reward_scales()
    np.savez(os.path.join(here, "reward_scales_snake.npz"), names=np.array(list(sc.keys())),
             values=np.array(list(sc.values()), dtype=np.float64))
    for name, seed, n, steps in CASES:
        out = run_case(seed, n, steps)
        np.savez_compressed(os.path.join(here, name + ".npz"), **out)
        nres = sum(len(out[f"reset_ids{t + 1}"]) for t in range(steps))
        print(name, "resets:", nres, "terminated:", sum(int(out[f'terminated{t+1}'].sum()) for t in range(steps)))
    scales = ref_loader.reference_reward_scales()
    np.savez(os.path.join(here, "reward_scales_v2.npz"), names=np.array(list(scales.keys())),
             values=np.array(list(scales.values()), dtype=np.float64))


# ---- zbot-6b-walking-m-v0 (manager-based task): the reference's own RewTerm / DoneTerm functions ------------------
M_CASES = [("m_v0_n64", 30, 64, 5), ("m_v0_n7", 31, 7, 4)]
#: every RewTerm function of zbotlab_manager/mdp/rewards.py that the cfg can name, with the cfg's params
#: (zbotlab_env_cfg.py:240-352)
M_FUNCS = [
    ("track_lin_vel_xy_yaw_frame_exp", {"command_name": "base_velocity", "std": 0.5}),
    ("track_ang_vel_z_world_exp", {"command_name": "base_velocity", "std": 0.5}),
    ("foot_step_length", {"command_name": None}),
    ("foot_downward", {}),
    ("foot_forward", {}),
    ("feet_gait", {"period": 2.0, "offset": [0.0, 0.5], "threshold": 0.55, "command_name": "base_velocity"}),
    ("feet_slide", {}),
    ("foot_clearance_reward", {"std": 0.05, "tanh_mult": 2.0, "target_height": 0.01}),
    ("feet_air_time_positive_biped", {"command_name": "base_velocity", "threshold": 0.3}),
    ("air_time_variance_penalty", {}),
    ("air_time_balance_penalty", {}),
    ("base_vel_forward", {"which_forward": 1}),
    ("feet_force_pattern", {}),
]


def run_m_case(seed, n, steps):
    from oracle.ref_harness import RefMHarness
    from oracle.m_mdp_oracle import synth_m_views
    torch.set_num_threads(1)
    h = RefMHarness(n)
    rng = np.random.default_rng(seed + 1000)
    out = {"seed": seed, "n": n, "steps": steps}
    h.env.feet_contact_forces_last[:] = torch.from_numpy(rng.uniform(0, 20, (n, 2)).astype(np.float32))
    h.env.feet_down_pos_last[:] = torch.from_numpy(rng.normal(0, 0.3, (n, 2, 3)).astype(np.float32))
    h.env.feet_force_sum[:] = torch.from_numpy(rng.normal(0, 0.05, n).astype(np.float32))
    out["state0/feet_contact_forces_last"] = h.env.feet_contact_forces_last.numpy().copy()
    out["state0/feet_down_pos_last"] = h.env.feet_down_pos_last.numpy().copy()
    out["state0/feet_force_sum"] = h.env.feet_force_sum.numpy().copy()
    for t, (view, _a, _r) in enumerate(synth_m_views(seed, n, steps)):
        cmd = np.stack([rng.uniform(-0.3, 0.3, n), rng.uniform(-0.2, 0.2, n), rng.uniform(-0.2, 0.2, n)], -1).astype(np.float32)
        cmd[rng.random(n) < 0.2] = 0.0
        ep = rng.integers(0, 1000, n)
        h.attach(view)
        h.command[:] = torch.from_numpy(cmd)
        h.env.episode_length_buf[:] = torch.from_numpy(ep)
        out[f"cmd{t}"], out[f"ep{t}"] = cmd, ep
        for func, params in M_FUNCS:
            out[f"val{t}/{func}"] = h.call(func, params).numpy().astype(np.float32).copy()
        out[f"feet_close{t}"] = h.feet_close(0.12).numpy().copy()
        for k in ("feet_contact_forces_last", "feet_down_pos_last", "feet_step_length", "feet_force_sum"):
            out[f"state{t + 1}/{k}"] = getattr(h.env, k).numpy().copy()
        if t == steps - 1:                                                  # reset_my_data on a few envs (rewards.py:37-43)
            ids = torch.from_numpy(np.sort(rng.choice(n, max(1, n // 4), replace=False)))
            h.reset_my_data(ids)
            out["reset_ids"] = ids.numpy()
            for k in ("feet_contact_forces_last", "feet_down_pos_last", "feet_step_length", "feet_force_sum"):
                out[f"state_reset/{k}"] = getattr(h.env, k).numpy().copy()
    return out


def make_m():
    here = os.path.dirname(os.path.abspath(__file__))
    for name, seed, n, steps in M_CASES:
        np.savez_compressed(os.path.join(here, name + ".npz"), **run_m_case(seed, n, steps))
        print("wrote", name)


if __name__ == "__main__":
    if "--m" not in sys.argv:
        main()
    make_m()
