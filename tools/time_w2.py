"""Back-to-back step time of the walking-v2 kernels: two warps per 32 envs (ZBOT_W2_CTAS register budgets) vs the
one-thread-per-env kernel (ZBOT_W2=0).   python tools/time_w2.py [envs ...]      (CUDA events, after warm-up)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zbot_lab_b200 import native  # noqa: E402
from zbot_lab_b200.stepper import NativeStepper  # noqa: E402
from zbot_lab_b200.utils import synthetic as syn  # noqa: E402


def time_one(n, env, steps=200, zero_actions=False):
    for k in ("ZBOT_W2", "ZBOT_W2_CTAS", "ZBOT_STEP_VARIANT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    st = NativeStepper(n, "cuda:0", native.make_cfg(n))
    st.reset_idx(None)
    g = torch.Generator(device="cuda:0").manual_seed(1)
    if not zero_actions:
        rng = np.random.default_rng(0)
        st.set_sim_state({k: torch.from_numpy(v).cuda() for k, v in syn.synth_sim_state(rng, n).items()})
        st.episode_length_buf[:] = torch.randint(0, 790, (n,), device="cuda:0", generator=g)
    acts = torch.zeros(4, n, 6, device="cuda:0") if zero_actions else torch.randn(4, n, 6, device="cuda:0", generator=g)
    for i in range(30):
        st.step(acts[i % 4])
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            st.step(acts[i % 4])
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / steps * 1e3)
    name = st.kernel_name
    st.close()
    return best, name


if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1:]] or [4096, 16384, 32768, 49152, 65536, 131072, 262144]
    variants = [("w2x8", {"ZBOT_W2_CTAS": "8"}), ("w2x6", {"ZBOT_W2_CTAS": "6"}), ("w2x10", {"ZBOT_W2_CTAS": "10"}),
                ("one-thread", {"ZBOT_W2": "0"})]
    out = {}
    for n in sizes:
        out[n] = {}
        for tag, env in variants:
            us, name = time_one(n, env)
            out[n][tag] = round(us, 2)
            print(f"{n:7d} envs  {tag:10s} {name:34s} {us:8.2f} us/step  {n / us * 1e6:.3e} env-steps/s", flush=True)
    us, _ = time_one(65536, {"ZBOT_W2_CTAS": "8"}, zero_actions=True)
    print(f"  65536 envs  w2x8 zero actions (no resets) {us:8.2f} us/step  {65536 / us * 1e6:.3e} env-steps/s")
    print(json.dumps(out))
