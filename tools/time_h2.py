"""Packed-halves step kernel (ZBOT_STEP_VARIANT=h128x2, csrc/zbot_h2.h) against the default one-thread kernel:
one-step agreement from identical states, then back-to-back step time.   python tools/time_h2.py [envs ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import time_w2  # noqa: E402
from zbot_lab_b200 import native  # noqa: E402
from zbot_lab_b200.stepper import NativeStepper  # noqa: E402
from zbot_lab_b200.utils import synthetic as syn  # noqa: E402


def agree(n=2048 + 37):
    rng = np.random.default_rng(33)
    a = torch.from_numpy(rng.normal(0, 0.7, (4, n, 6)).astype(np.float32)).cuda()
    res = []
    for env in ({"ZBOT_W2": "0"}, {"ZBOT_STEP_VARIANT": "h128x2"}):
        for k in ("ZBOT_W2", "ZBOT_W2_CTAS", "ZBOT_STEP_VARIANT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        r = np.random.default_rng(5)
        st = NativeStepper(n, "cuda:0", native.make_cfg(n))
        st.reset_idx(None)
        st.set_sim_state({k: torch.from_numpy(v).cuda() for k, v in syn.synth_sim_state(r, n).items()})
        st.episode_length_buf[:] = torch.from_numpy(r.integers(0, 1000, n).astype(np.int64)).cuda()
        outs = []
        for t in range(4):
            o = st.step(a[t])
            outs.append([x.clone() for x in o] + [st.state.buf.clone(), st.episode_length_buf.clone()])
            if res:
                st.state.buf.copy_(res[0][t][4])
                st.episode_length_buf.copy_(res[0][t][5])
        print(st.kernel_name)
        res.append(outs)
        st.close()
    for t in range(4):
        o0, r0, te0, tr0, s0, ep0 = res[0][t]
        o1, r1, te1, tr1, s1, ep1 = res[1][t]
        same = te0 == te1
        d = (o0[same] - o1[same]).abs()
        dr = (r0[same] - r1[same]).abs()
        print(f"step {t}: term agree {float(same.float().mean()):.4f} trunc equal {torch.equal(tr0, tr1)} obs[:10] {float(d[:, :10].max()):.2e} "
              f"obs[10:16] {float(d[:, 10:16].max()):.2e} obs[16:] {float(d[:, 16:].max()):.2e} rew q999 {float(torch.quantile(dr, 0.999)):.2e} max {float(dr.max()):.2e}")


if __name__ == "__main__":
    agree()
    for n in [int(x) for x in sys.argv[1:]] or [65536, 32768, 131072, 16384]:
        for tag, env in (("default", {"ZBOT_W2": "0"}), ("h128x2", {"ZBOT_STEP_VARIANT": "h128x2"})):
            us, name = time_w2.time_one(n, env)
            print(f"{n:7d} envs {tag:8s} {name:32s} {us:8.2f} us/step {n / us * 1e6:.3e} env-steps/s", flush=True)
