"""A few control steps of the fused kernel for ncu captures:  python tools/prof_step.py <envs> [steps] [task]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zbot_lab_b200 import native  # noqa: E402
from zbot_lab_b200.stepper import NativeStepper  # noqa: E402
from zbot_lab_b200.utils import synthetic as syn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
task = sys.argv[3] if len(sys.argv) > 3 else "walk"
tid = {"walk": native.TASK_WALKING_V2, "snake": native.TASK_SNAKE_V0, "v4": native.TASK_WALKING_V4, "m": native.TASK_WALKING_M}[task]
if task == "m":
    terms = [(f, w, p) for _, f, w, p in native.M_FLAT_TERMS if f != "is_terminated"]
    st = NativeStepper(n, "cuda:0", native.make_m_cfg(n, terms, is_terminated_weight=-200.0, act_clip=0.04 * np.pi))
else:
    st = NativeStepper(n, "cuda:0", native.make_cfg(n, task=tid))
if task == "m":
    st.reset_idx_m(None)
    st.state.set("joint_speed_limit", torch.rand(n, 1, device="cuda:0") * 0.7 + 0.3)
elif task == "v4":
    st.reset_idx_v4(None)
    st.state.set("base_pos_y_err_sum", torch.rand(n, 1, device="cuda:0") * 3 + 3)
else:
    st.reset_idx(None)
if task == "walk":
    rng = np.random.default_rng(0)
    st.set_sim_state({k: torch.from_numpy(v).cuda() for k, v in syn.synth_sim_state(rng, n).items()})
g = torch.Generator(device="cuda:0").manual_seed(1)
st.episode_length_buf[:] = torch.randint(0, 790, (n,), device="cuda:0", generator=g)
acts = torch.randn(4, n, 6, device="cuda:0", generator=g)
for i in range(steps):
    st.step(acts[i % 4])
torch.cuda.synchronize()
print("ok", n, steps, task)
