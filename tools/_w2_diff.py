import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from zbot_lab_b200 import native
from zbot_lab_b200.stepper import NativeStepper
from zbot_lab_b200.utils import synthetic as syn
from oracle import cpu_port
def make(n, env):
    for k in ("ZBOT_W2", "ZBOT_W2_CTAS"): os.environ.pop(k, None)
    os.environ.update(env)
    st = NativeStepper(n, "cuda:0", native.make_cfg(n))
    st.reset_idx(None)
    r = np.random.default_rng(77)
    st.set_sim_state({k: torch.from_numpy(v).cuda() for k, v in syn.synth_sim_state(r, n).items()})
    return st
n = 300
names = ["x6", "x6b", "x8", "x8b", "x5", "old"]
sts = [make(n, {"ZBOT_W2_CTAS": "6"}), make(n, {"ZBOT_W2_CTAS": "6"}), make(n, {"ZBOT_W2_CTAS": "8"}), make(n, {"ZBOT_W2_CTAS": "8"}),
       make(n, {"ZBOT_W2_CTAS": "5"}), make(n, {"ZBOT_W2": "0"})]
# CPU port (float32, one-chain formulation) from the same state
port = cpu_port.PortEnv(n, np.float32)
r = np.random.default_rng(77)
port.set_sim_state(syn.synth_sim_state(r, n))
rng = np.random.default_rng(5)
fields = ("root_pos", "root_quat", "root_lin_vel", "root_ang_vel", "joint_pos", "joint_vel")
for t in range(3):
    a = rng.normal(0, 1, (n, 6)).astype(np.float32)
    at = torch.from_numpy(a).cuda()
    outs = [[x.clone() for x in s.step(at)] for s in sts]
    po = port.step(a)
    torch.cuda.synchronize()
    ref = np.concatenate([port.field(k, w) for k, w in zip(fields, (3, 4, 3, 3, 6, 6))], 1)
    for nm, s, o in zip(names, sts, outs):
        g = torch.cat([s.state.get(k) for k in fields], 1).cpu().numpy()
        d = np.abs(g - ref)
        bad = (d.max(1) > 1e-2).sum()
        print(t, nm, "vs CPU port: max", float(d.max()), "median rowmax", float(np.median(d.max(1))), "rows>1e-2:", int(bad),
              "flags eq", bool((o[2].cpu().numpy().astype(bool) == po[2]).all() and (o[3].cpu().numpy().astype(bool) == po[3]).all()),
              "| equal to x6:", bool(torch.equal(s.state.buf, sts[0].state.buf)))
