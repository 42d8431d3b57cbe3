"""Step time with and without programmatic dependent launch (ZBOT_PDL; ZBOT_PDL_EARLY = the step kernel releases its
dependents at its start), back to back:
   python tools/time_pdl.py [envs ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zbot_lab_b200.stepper import NativeStepper  # noqa: E402
from zbot_lab_b200.utils import synthetic as syn  # noqa: E402


def run(n, pdl, steps=300, early=False, fused=False):
    os.environ["ZBOT_PDL"] = "1" if pdl else "0"
    os.environ["ZBOT_FUSED_STATS"] = "1" if fused else "0"
    os.environ["ZBOT_PDL_EARLY"] = "1" if early else "0"
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
        t2, c2 = run(n, True, early=True)
        t3, c3 = run(n, True, fused=True)
        t4, c4 = run(n, True, early=True, fused=True)
        t5, c5 = run(n, False, fused=True)
        print(f"envs {n:6d}: separate statistics kernel: plain {t0:7.2f} us/step   PDL {t1:7.2f}   PDL + early trigger {t2:7.2f}   "
              f"same result: {c0 == c1 == c2}", flush=True)
        print(f"envs {n:6d}: fused statistics:           plain {t5:7.2f} us/step   PDL {t3:7.2f}   PDL + early trigger {t4:7.2f}   "
              f"same state: {c0[0] == c3[0] == c4[0] == c5[0]}  stats sums {c0[1]!r} vs {c3[1]!r}", flush=True)
