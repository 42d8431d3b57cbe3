"""Long random-action soak of every task's fused step: finiteness, unit root quaternion, bounded joint speed.
   python tools/soak.py [steps] [envs]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from time_task import make  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
for task in ("walk", "snake", "v4", "m"):
    st = make(task, n)
    g = torch.Generator(device="cuda:0").manual_seed(7)
    resets = 0.0
    worst_qd = 0.0
    for t in range(steps):
        scale = 0.3 + 2.7 * ((t // 500) % 2)                      # alternate gentle / violent actions
        obs, rew, term, trunc = st.step(torch.randn(n, 6, device="cuda:0", generator=g) * scale)
        if t % 1000 == 999:
            resets += float(st.stats_ring[:, 16].sum())
            q = st.state.get("root_quat")
            assert torch.isfinite(st.state.buf).all() and torch.isfinite(obs).all() and torch.isfinite(rew).all(), (task, t)
            assert float((q.norm(dim=1) - 1).abs().max()) < 1e-4, (task, t)
            worst_qd = max(worst_qd, float(st.state.get("joint_vel").abs().max()))
    print(f"{task:5s}: {steps} steps x {n} envs ok; resets in the sampled 64-step windows {resets:.0f}; max |joint vel| seen {worst_qd:.1f} rad/s",
          flush=True)
    st.close()
