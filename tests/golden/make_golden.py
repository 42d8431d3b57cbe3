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


if __name__ == "__main__":
    main()
