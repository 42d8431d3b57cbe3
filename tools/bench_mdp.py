"""Times the MDP-only kernel (BASELINE.json configs[0] shape, scaled up) and reports its HBM
roofline fraction.  Algorithmic bytes per env-step (restated in DESIGN.md §4):
read 1220 B + write 350 B = 1570 B.   python tools/bench_mdp.py [envs] [steps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zbot_lab_b200.stepper import NativeStepper  # noqa: E402
from zbot_lab_b200.utils import synthetic as syn  # noqa: E402

MDP_ALGO_BYTES = 1570


def bench_mdp(n=65536, steps=50, warm=5, flush=True, dev="cuda:0", mode=None, sets=None):
    """mode "write" (default when flush=True): a 256 MB fill between steps (the L2 is left full of DIRTY lines the step then has
    to evict: their write-back shares the DRAM with the step's own traffic); "write+read": the fill, then a 256 MB read pass
    (inputs equally absent from the L2, which is left clean); "rotate": no flush, `sets` (default 6) input sets used round-robin
    (> L2 in total: an input has been evicted long before it comes round again); "none": two sets back to back."""
    mode = mode or ("write" if flush else "none")
    if mode == "graph-rotate":
        return bench_mdp_graph(n, steps, dev, sets or 6)
    sets = sets or (6 if mode == "rotate" else 2)
    st = NativeStepper(n, dev)
    st.mdp_init()
    rng = np.random.default_rng(0)
    org = torch.from_numpy(syn.env_origins_grid(n)).to(dev)
    S = [{k: torch.from_numpy(v).to(dev) for k, v in syn.synth_articulation_state(rng, n, org.cpu().numpy(), 0.002).items()}
         for _ in range(sets)]
    acts = torch.randn(4, n, 6, device=dev)
    st.mdp_observe(S[0], org)
    fl = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if mode in ("write", "write+read") else None
    fr = torch.zeros(64 << 20, dtype=torch.float32, device=dev) if mode == "write+read" else None
    for i in range(warm):
        st.mdp_step(S[i % sets], org, acts[i % 4])
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        if fl is not None:
            fl.fill_(i & 0xFF)
        if fr is not None:
            fr.sum()
        ev[i][0].record()
        st.mdp_step(S[i % sets], org, acts[i % 4])
        ev[i][1].record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    st.close()
    return ms


def bench_mdp_graph(n=65536, replays=20, dev="cuda:0", sets=6):
    """"Inputs larger than L2" instead of a flush: `sets` input sets (6 x 89 MB at 65536 envs = 534 MB against a 126 MB L2)
    used round-robin, 2 x sets steps captured as ONE CUDA graph (back-to-back launches, no host in the loop) and replayed;
    CUDA events around each replay.  The MDP state (19 MB) is the handle's own and is re-read every step, as in real use.
    A write-fill flush leaves the L2 full of dirty lines whose write-back (126 MB) shares the DRAM with the step's reads
    (108 MB at 65536 envs): that measures the flush as much as the kernel."""
    st = NativeStepper(n, dev)
    st.mdp_init()
    rng = np.random.default_rng(0)
    org = torch.from_numpy(syn.env_origins_grid(n)).to(dev)
    S = [{k: torch.from_numpy(v).to(dev) for k, v in syn.synth_articulation_state(rng, n, org.cpu().numpy(), 0.002).items()}
         for _ in range(sets)]
    acts = torch.randn(4, n, 6, device=dev)
    st.mdp_observe(S[0], org)
    per = 2 * sets
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for i in range(per):
            st.mdp_step(S[i % sets], org, acts[i % 4])
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(per):
            st.mdp_step(S[i % sets], org, acts[i % 4])
    g.replay()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(replays)]
    for a, b in ev:
        a.record()
        g.replay()
        b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / replays / per
    st.close()
    return ms


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    mode = sys.argv[3] if len(sys.argv) > 3 else "write"
    ms = bench_mdp(n, steps, mode=mode)
    peak = 6545.3
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    gbs = MDP_ALGO_BYTES * n / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "zbot_mdp_kernel<true>" if os.environ.get("ZBOT_MDP_PIPE") == "0" else "zbot_mdp_pipe_kernel",
                      "l2": mode, "envs": n, "ms_per_step": ms, "env_steps_per_s": n / (ms * 1e-3),
                      "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak, "algorithmic_bytes_per_env_step": MDP_ALGO_BYTES}))
