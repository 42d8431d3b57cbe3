"""Back-to-back timing of one task's fused step:  python tools/time_task.py <walk|snake|v4|m> [envs ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zbot_lab_b200 import native  # noqa: E402
from zbot_lab_b200.stepper import NativeStepper  # noqa: E402


def make(task, n):
    if task == "m":
        terms = [(f, w, p) for _, f, w, p in native.M_FLAT_TERMS if f != "is_terminated"]
        st = NativeStepper(n, "cuda:0", native.make_m_cfg(n, terms, is_terminated_weight=-200.0, act_clip=0.04 * np.pi))
        st.reset_idx_m(None)
        st.state.set("joint_speed_limit", torch.rand(n, 1, device="cuda:0") * 0.7 + 0.3)
        return st
    tid = {"walk": native.TASK_WALKING_V2, "snake": native.TASK_SNAKE_V0, "v4": native.TASK_WALKING_V4}[task]
    st = NativeStepper(n, "cuda:0", native.make_cfg(n, task=tid))
    if task == "v4":
        st.reset_idx_v4(None)
        st.state.set("base_pos_y_err_sum", torch.rand(n, 1, device="cuda:0") * 3 + 3)
    else:
        st.reset_idx(None)
    if task == "snake":
        st.state.set("joint_speed_limit", (torch.rand(n, 1, device="cuda:0") * 1.8 + 0.2) * 3.14159265)
    return st


if __name__ == "__main__":
    task = sys.argv[1] if len(sys.argv) > 1 else "walk"
    for n in [int(x) for x in sys.argv[2:]] or [4096, 65536]:
        st = make(task, n)
        g = torch.Generator(device="cuda:0").manual_seed(1)
        st.episode_length_buf[:] = torch.randint(0, 790, (n,), device="cuda:0", generator=g)
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
        print(f"{task:5s} envs {n:6d} ZBOT_CTAS3={os.environ.get('ZBOT_CTAS3', 'auto'):4s}: {best:8.2f} us/step  {n / best:8.2f} M env-steps/s", flush=True)
        st.close()
