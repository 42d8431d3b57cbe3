"""GPU tests of the dynamics model through physical invariants (the reference's simulator, PhysX, is closed and absent:
what CAN be checked without it is that the fused kernel's articulated-body + contact model obeys mechanics).
Every state evolution below is produced by the sm_100a kernel through the C ABI; the oracle is only the calculator of
momentum / energy from a state."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _stepper(n, task=0, **kw):
    from zbot_lab_b200 import native
    from zbot_lab_b200.stepper import NativeStepper
    st = NativeStepper(n, DEV, native.make_cfg(n, task=task, y_err_limit=1e9, termination_height=-1e9,
                                               contact_died_force=1e9, **kw))
    st.reset_idx(None)
    return st


def _state(st):
    g = lambda k: st.state.get(k).cpu().numpy().astype(np.float64)
    return {"root_pos": g("root_pos"), "root_quat": g("root_quat"), "root_lin_vel": g("root_lin_vel"),
            "root_ang_vel": g("root_ang_vel"), "joint_pos": g("joint_pos"), "joint_vel": g("joint_vel")}


def _momentum(state, model=None):
    """(kinetic energy, linear momentum, angular momentum about the origin) of a state: the oracle as a calculator."""
    from oracle.dyn_oracle import DynOracle, DynParams
    n = state["root_pos"].shape[0]
    o = DynOracle(n, DynParams(model=model, gravity=0.0), model=model)
    o.set_state(state)
    return o.energy_momentum()


def test_free_fall_is_exact_and_leaves_the_joints_alone(walk_kernel):
    """No contact, no drive: every body accelerates at -g, so v_z = -g t exactly for semi-implicit Euler, the fall
    distance is g dt^2 k (k + 1) / 2, and the joints of a uniformly accelerated articulated body do not move.
    (Each of the three walking step kernels.)"""
    n, steps = 256, 10
    st = _stepper(n, kp=0.0, kd=0.0, contact_alpha=0.0)
    assert st.kernel_name.startswith(walk_kernel), st.kernel_name
    rp = st.state.get("root_pos")
    rp[:, 2] += 2.0
    st.state.set("root_pos", rp)
    q0 = st.state.get("joint_pos").clone()
    z0 = st.state.get("root_pos")[:, 2].clone()
    for _ in range(steps):
        st.step(torch.zeros(n, 6, device=DEV))
    k, dt, g = 4 * steps, 0.005, 9.81
    vz = st.state.get("root_lin_vel")[:, 2]
    assert torch.allclose(vz, torch.full_like(vz, -g * dt * k), rtol=2e-5)
    dz = z0 - st.state.get("root_pos")[:, 2]
    assert torch.allclose(dz, torch.full_like(dz, g * dt * dt * k * (k + 1) / 2), rtol=1e-4)
    assert float((st.state.get("joint_pos") - q0).abs().max()) < 2e-5
    assert float(st.state.get("joint_vel").abs().max()) < 2e-3
    assert float(st.state.get("root_ang_vel").abs().max()) < 2e-3
    st.close()


def test_momentum_and_energy_drift_is_first_order_in_free_flight(walk_kernel):
    """g = 0, no contact, no drive, random spin and joint velocities.  Semi-implicit Euler in generalised coordinates
    conserves momentum and energy only to first order in dt: over the same 0.2 s the drift with dt = 2.5 ms must be
    about half the drift with dt = 5 ms (a wrong bias / Coriolis term would give an O(1), dt-independent error), and
    small in absolute terms."""
    n = 128
    drifts = []
    for sim_dt, steps in ((0.005, 10), (0.0025, 20)):
        rng = np.random.default_rng(0)
        st = _stepper(n, kp=0.0, kd=0.0, contact_alpha=0.0, gravity=0.0, sim_dt=sim_dt)
        assert st.kernel_name.startswith(walk_kernel), st.kernel_name
        rp = st.state.get("root_pos")
        rp[:, 2] += 2.0
        st.state.set("root_pos", rp)
        st.state.set("root_ang_vel", _t(rng.normal(0, 1.0, (n, 3)).astype(np.float32)))
        st.state.set("root_lin_vel", _t(rng.normal(0, 0.5, (n, 3)).astype(np.float32)))
        st.state.set("joint_vel", _t(rng.normal(0, 1.0, (n, 6)).astype(np.float32)))
        E0, P0, L0 = _momentum(_state(st))
        for _ in range(steps):
            st.step(torch.zeros(n, 6, device=DEV))
        E1, P1, L1 = _momentum(_state(st))
        drifts.append((np.abs(P1 - P0).mean(), np.abs(L1 - L0).mean(), np.abs(E1 - E0).mean() / np.abs(E0).mean()))
        st.close()
    (p5, l5, e5), (p2, l2, e2) = drifts
    assert p2 < 0.7 * p5 and l2 < 0.7 * l5 and e2 < 0.7 * e5, drifts
    assert p5 < 0.01 and l5 < 0.01 and e5 < 0.02, drifts


def test_static_stand_carries_the_weight(walk_kernel):
    """Zero actions from the default pose: the PD drive holds the pose, the two feet carry m g between them, the base
    stays at the reference's printed height 0.2545 m (…env_v2.py:403), nothing terminates.  (Each of the three kernels,
    through its own export hook.)"""
    from zbot_lab_b200 import native
    from zbot_lab_b200.stepper import NativeStepper
    n = 64
    st = NativeStepper(n, DEV)
    assert st.kernel_name.startswith(walk_kernel), st.kernel_name
    st.reset_idx(None)
    ex = st.alloc_export()
    for t in range(100):
        obs, rew, term, trunc = st.step(torch.zeros(n, 6, device=DEV), export=ex)
        assert not term.any()
    fz = ex["net_forces_w_history1"][:, 0, :, 2]              # newest slot, all 12 sensor bodies
    mg = 12 * 0.25042 * 9.81
    assert torch.allclose(fz.sum(-1), torch.full((n,), mg, device=DEV), rtol=2e-3)
    assert float(fz[:, [9, 10]].min()) > 0.15 * mg            # both feet loaded (sensor order: foot_0 = 9, foot_1 = 10)
    base_z = ex["body_link_pos_w1"][:, 6, 2]
    assert float((base_z - 0.2545).abs().max()) < 1e-3
    assert float(st.state.get("joint_vel").abs().max()) < 1e-3
    st.close()


def test_sliding_snake_decelerates_at_mu_g():
    """Coulomb friction, mu = 1: a snake pushed along the ground at 1.5 m/s loses g * mu * t of speed while it slides
    (regularised cone: only while |v_t| is well above the regularisation speed), then stops and stays put."""
    from zbot_lab_b200 import native
    n = 64
    st = _stepper(n, task=native.TASK_SNAKE_V0)
    st.state.set("joint_speed_limit", 3.14159)
    for _ in range(10):                                        # let the 12 spheres settle onto the plane
        st.step(torch.zeros(n, 6, device=DEV))
    v0 = 1.5
    lin = st.state.get("root_lin_vel")
    lin[:, 0] = v0                                             # along the chain axis (x): it can only slide, not roll
    st.state.set("root_lin_vel", lin)
    speeds = []
    for _ in range(5):
        st.step(torch.zeros(n, 6, device=DEV))
        speeds.append(float(st.state.get("root_lin_vel")[:, 0].mean()))
    dec = (v0 - speeds[3]) / (4 * 0.02)                        # average deceleration over the first 80 ms
    assert 0.8 * 9.81 < dec < 1.1 * 9.81, (dec, speeds)
    for _ in range(20):
        st.step(torch.zeros(n, 6, device=DEV))
    assert float(st.state.get("root_lin_vel").abs().max()) < 0.05
    st.close()


def test_long_random_rollout_stays_bounded(walk_kernel):
    """2000 control steps of random actions with every termination disabled: contact + stiff implicit drive never blow
    up (velocities bounded, quaternion normalised, no NaN).  (Each of the three walking step kernels.)"""
    n = 512
    st = _stepper(n)
    assert st.kernel_name.startswith(walk_kernel), st.kernel_name
    g = torch.Generator(device=DEV).manual_seed(3)
    for t in range(2000):
        st.step(torch.randn(n, 6, device=DEV, generator=g) * 2.0)
        if t % 250 == 249:
            assert torch.isfinite(st.state.buf).all()
            assert float(st.state.get("joint_vel").abs().max()) < 80.0
            assert float(st.state.get("root_lin_vel").abs().max()) < 20.0
            qn = st.state.get("root_quat").norm(dim=-1)
            assert float((qn - 1).abs().max()) < 1e-4
    st.close()
