"""Step time with and without programmatic dependent launch (ZBOT_PDL), back to back and with a gap kernel in between:
   python tools/time_pdl.py [envs ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zbot_lab_b200.stepper import NativeStepper  # noqa: E402
from zbot_lab_b200.utils import synthetic as syn  # noqa: E402


def run(n, pdl, steps=300):
    os.environ["ZBOT_PDL"] = "1" if pdl else "0"
    st = NativeStepper(n, "cuda:0")
    st.reset_idx(None)
    rng = np.random.default_rng(0)
    st.set_sim_state({k: torch.from_numpy(v).cuda() for k, v in syn.synth_sim_state(rng, n).items()})
    g = torch.Generator(device="cuda:0").manual_seed(1)
    st.episode_length_buf[:] = torch.randint(0, 1000, (n,), device="cuda:0", generator=g)
    acts = torch.randn(16, n, 6, device="cuda:0", generator=g)
    outs = []
    for i in range(30):
        st.step(acts[i % 16])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        a.record()
        for i in range(steps):
            st.step(acts[i % 16])
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / steps)
    chk = (st.state.buf.double().sum().item(), st.stats.clone().cpu().numpy().sum())
    st.close()
    return best, chk


if __name__ == "__main__":
    for n in [int(x) for x in sys.argv[1:]] or [4096, 65536]:
        t0, c0 = run(n, False)
        t1, c1 = run(n, True)
        print(f"envs {n:6d}: plain {t0:7.2f} us/step   PDL {t1:7.2f} us/step   same result: {c0 == c1}", flush=True)
