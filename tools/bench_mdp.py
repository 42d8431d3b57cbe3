"""Times the MDP-only kernel (BASELINE.json configs[0] shape, scaled up) and reports its HBM
roofline fraction.  Algorithmic bytes per env-step (restated in DESIGN.md §4):
read 1220 B + write 350 B = 1570 B.   python tools/bench_mdp.py [envs] [steps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zbot_lab_b200.stepper import NativeStepper  # noqa: E402
from zbot_lab_b200.utils import synthetic as syn  # noqa: E402

MDP_ALGO_BYTES = 1570


def bench_mdp(n=65536, steps=50, warm=5, flush=True, dev="cuda:0"):
    st = NativeStepper(n, dev)
    st.mdp_init()
    rng = np.random.default_rng(0)
    org = torch.from_numpy(syn.env_origins_grid(n)).to(dev)
    S = [{k: torch.from_numpy(v).to(dev) for k, v in syn.synth_articulation_state(rng, n, org.cpu().numpy(), 0.002).items()}
         for _ in range(2)]
    acts = torch.randn(4, n, 6, device=dev)
    st.mdp_observe(S[0], org)
    fl = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush else None
    for i in range(warm):
        st.mdp_step(S[i % 2], org, acts[i % 4])
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        if fl is not None:
            fl.fill_(i & 0xFF)
        ev[i][0].record()
        st.mdp_step(S[i % 2], org, acts[i % 4])
        ev[i][1].record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    st.close()
    return ms


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    ms = bench_mdp(n, steps)
    peak = 6545.3
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    gbs = MDP_ALGO_BYTES * n / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "zbot_mdp_kernel<true>", "envs": n, "ms_per_step": ms, "env_steps_per_s": n / (ms * 1e-3),
                      "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak, "algorithmic_bytes_per_env_step": MDP_ALGO_BYTES}))
