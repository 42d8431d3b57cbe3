"""tcgen05 build of the act kernel (ZBOT_POLICY_TC=2) against the CUDA-core build (=0) and the mma.sync build (=1) on the same
observations: max differences and time per launch.   python tools/check_policy_tc5.py [envs]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_policy as T  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
outs, times = {}, {}
for tc in ("0", "1", "2"):
    os.environ["ZBOT_POLICY_TC"] = tc
    st, ac, pol, b = T._setup(n, seed=3)
    obs = torch.randn(n, 23, device="cuda:0")
    T._act(st, pol, obs, b)
    outs[tc] = {k: v.clone() for k, v in b.items()}
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        T._act(st, pol, obs, b)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(50):
            st.policy_act(pol, obs, b["obs_out"], b["act"], b["logp"], b["value"], b["mu"], b["sigma"], seed=11)
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    times[tc] = 1e3 * e0.elapsed_time(e1) / 500
    with torch.no_grad():
        mu64 = ac.actor.double()(obs.double())
        ac.float()
    print(f"ZBOT_POLICY_TC={tc}: {times[tc]:.2f} us per launch; |mu - float64| max {float((outs[tc]['mu'].double() - mu64).abs().max()):.3e}", flush=True)
    st.close()
for k in ("mu", "value", "act", "logp"):
    print(k, "tc5 vs cuda-core:", float((outs["2"][k] - outs["0"][k]).abs().max()), " mma.sync vs cuda-core:", float((outs["1"][k] - outs["0"][k]).abs().max()))
