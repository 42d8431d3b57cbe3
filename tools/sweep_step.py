"""Tuning sweep of the fused step kernel: register budget (ZBOT_STEP_MIN_BLOCKS) x block size.
Run on a GPU box:  python tools/sweep_step.py [envs ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zbot_lab_b200.stepper import NativeStepper  # noqa: E402
from zbot_lab_b200.utils import synthetic as syn  # noqa: E402


VARIANTS = (os.environ.get("SWEEP_VARIANTS") or "128x2,128x3,128x4,64x5,32x10,32x11,64x6,32x12,32x13,32x14,64x7,32x16").split(",")


def time_cfg(n, variant, steps=200, warm=30):
    os.environ["ZBOT_STEP_VARIANT"] = variant
    st = NativeStepper(n, "cuda:0")
    st.reset_idx(None)
    rng = np.random.default_rng(0)
    st.set_sim_state({k: torch.from_numpy(v).cuda() for k, v in syn.synth_sim_state(rng, n).items()})
    g = torch.Generator(device="cuda:0").manual_seed(1)
    st.episode_length_buf[:] = torch.randint(0, 1000, (n,), device="cuda:0", generator=g)
    acts = torch.randn(16, n, 6, device="cuda:0", generator=g)
    for i in range(warm):
        st.step(acts[i % 16])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        st.step(acts[i % 16])
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / steps
    st.close()
    return us


if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1:]] or [4096, 65536]
    for n in sizes:
        for v in VARIANTS:
            us = min(time_cfg(n, v) for _ in range(2))
            print(f"envs {n:6d} variant {v:>6s}: {us:8.2f} us/step  {n / us:8.2f} M env-steps/s", flush=True)
