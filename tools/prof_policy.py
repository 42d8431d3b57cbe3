"""One launch of `zbot_policy_act_kernel` at 4096 envs for an ncu capture:
   ncu --set full --clock-control none --import-source on -k regex:zbot_policy_act -c 1 -o gpurun_out/policy python tools/prof_policy.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_rollout  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
r = bench_rollout.make(n, "cuda:0", False, True)
st, b, pol = r._fused, r.buf, r._policy_struct()
obs = r.env.get_observations()["policy"]
for _ in range(3):
    st.policy_act(pol, obs, b["obs"][0], b["act"][0], b["logp"][0], b["val"][0], b["mu"][0], b["sigma"][0])
torch.cuda.synchronize()
print("ok")
