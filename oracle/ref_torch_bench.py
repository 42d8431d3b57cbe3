"""TEST / BENCH INFRASTRUCTURE -- times the reference's OWN torch MDP code on the host cores.

north_star / BASELINE.md §2 row 1: "the reference's torch ... CPU path timed on the box's own host cores".  The physics
half of that path (Isaac Lab + PhysX) is closed and not installable, so only the MDP half can be the reference's own
code: `_pre_physics_step + _get_dones + _get_rewards + _reset_idx + _get_observations` of
`source/zbot/zbot/tasks/zbot6b_direct/zbot_direct_6dof_bipedal_env_v2.py`, loaded UNMODIFIED through
`oracle/ref_loader.py` (stub modules) and driven in `DirectRLEnv.step` order by `oracle/ref_harness.RefMdpHarness` on
synthetic articulation / contact state (the SURVEY §8(d) probe).

Needs a reference tree (`ZBOT_REFERENCE_ROOT`, default /root/reference).  It does not exist on the GPU box unless someone
puts one there; `bench.py` calls this only when it does and reports the result as `cpu_baseline_torch_mdp`.
   python -m oracle.ref_torch_bench [envs] [steps]
"""
from __future__ import annotations

import json
import os
import sys
import time


def available() -> bool:
    from . import ref_loader
    return ref_loader.reference_available()


def time_reference_torch_mdp(n_envs: int = 65536, steps: int = 20, warmup: int = 3, threads: int | None = None) -> dict:
    import numpy as np
    import torch

    from zbot_lab_b200.assets import zbot_6s as Z
    from zbot_lab_b200.utils import synthetic as syn

    from .ref_harness import RefMdpHarness

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    rng = np.random.default_rng(0)
    org = syn.env_origins_grid(n_envs)
    S = [{k: torch.from_numpy(v) for k, v in syn.synth_articulation_state(rng, n_envs, org, 0.002).items()} for _ in range(2)]
    dj = torch.tensor(Z.DEFAULT_JOINT_POS, dtype=torch.float32).repeat(n_envs, 1)
    drs = torch.zeros(n_envs, 13)
    drs[:, :3] = torch.tensor(Z.DEFAULT_ROOT_POS)
    drs[:, 3] = 1.0
    h = RefMdpHarness(n_envs, torch.from_numpy(org), syn.reset_tables(), syn.index_sets(), dj, drs)
    h.env.episode_length_buf[:] = torch.from_numpy(rng.integers(0, 1000, n_envs))
    h.attach(S[0])
    h.observe()
    acts = torch.randn(4, n_envs, 6)
    for i in range(warmup):
        h.step(acts[i % 4], S[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        h.step(acts[i % 4], S[i % 2])
    dt = time.perf_counter() - t0
    return {"value": n_envs * steps / dt, "unit": "env-steps/s", "cores": threads, "kind": "reference",
            "ms_per_step": 1e3 * dt / steps,
            "sample": f"{n_envs} envs x {steps} steps: the reference's own zbot_direct_6dof_bipedal_env_v2.py MDP methods "
                      f"(_pre_physics_step, _get_dones, _get_rewards, _reset_idx, _get_observations) on CPU torch, "
                      f"synthetic articulation state; NO physics (Isaac Lab / PhysX are closed)"}


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    print(json.dumps(time_reference_torch_mdp(n, k)))
