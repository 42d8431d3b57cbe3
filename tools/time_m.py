"""Back-to-back timing of the manager-task kernel:  python tools/time_m.py [envs ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zbot_lab_b200 import native  # noqa: E402
from zbot_lab_b200.stepper import NativeStepper  # noqa: E402

for n in [int(x) for x in sys.argv[1:]] or [4096, 65536]:
    terms = [(f, w, p) for _, f, w, p in native.M_FLAT_TERMS if f != "is_terminated"]
    st = NativeStepper(n, "cuda:0", native.make_m_cfg(n, terms, is_terminated_weight=-200.0, act_clip=0.04 * np.pi))
    st.reset_idx_m(None)
    g = torch.Generator(device="cuda:0").manual_seed(1)
    st.episode_length_buf[:] = torch.randint(0, 1000, (n,), device="cuda:0", generator=g)
    acts = torch.randn(16, n, 6, device="cuda:0", generator=g)
    for i in range(30):
        st.step(acts[i % 16])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        a.record()
        for i in range(200):
            st.step(acts[i % 16])
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / 200)
    print(f"manager task envs {n:6d}: {best:8.2f} us/step  {n / best:8.2f} M env-steps/s  resets/step {float(st.stats[16]):.0f}", flush=True)
    st.close()
