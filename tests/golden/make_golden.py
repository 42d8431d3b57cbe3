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


def main():
    here = os.path.dirname(os.path.abspath(__file__))
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
