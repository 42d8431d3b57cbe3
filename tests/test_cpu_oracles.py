"""CPU tests (no GPU): dynamics-oracle invariants, the float64 CPU instantiation of the kernel's
own arithmetic against the independent oracle, and the composed full-step oracle."""
import numpy as np
import pytest

from oracle.dyn_oracle import DynOracle, DynParams


def test_fk_known_answers_from_reference_comments():
    """…env_v2.py:403-404 and …env_v4.py:814-816 print these values at the default pose."""
    o = DynOracle(1)
    ls = o.link_state()
    assert ls["body_link_pos"][0, 6, 2] == pytest.approx(0.2545, abs=5e-5)
    assert np.allclose(ls["body_link_quat"][0, 6], [0.6003, -0.6003, -0.3735, -0.3739], atol=5e-5)
    assert ls["body_link_pos"][0, 0, 2] == pytest.approx(0.0, abs=1e-9)
    assert ls["body_link_pos"][0, 11, 2] == pytest.approx(5.3035e-2, abs=5e-7)


def test_free_motion_conserves_energy_and_momentum_first_order():
    """No gravity / contact / PD: drift must shrink ~linearly with dt (wrong bias terms give O(1))."""
    drift = []
    for dt in (2e-4, 1e-4):
        rng = np.random.default_rng(0)
        o = DynOracle(4, DynParams(gravity=0.0, contacts=False, pd=False, dt=dt))
        o.root_pos[:, 2] += 1.0
        o.root_ang_vel = rng.normal(0, 2, (4, 3))
        o.root_lin_vel = rng.normal(0, 1, (4, 3))
        o.qd = rng.normal(0, 3, (4, 6))
        E0, P0, L0 = o.energy_momentum()
        for _ in range(int(round(0.1 / dt))):
            o.substep(o.q)
        E1, P1, L1 = o.energy_momentum()
        drift.append((np.abs((E1 - E0) / E0).max(), np.abs(P1 - P0).max(), np.abs(L1 - L0).max()))
    for a, b in zip(drift[0], drift[1]):
        assert b < 0.7 * a + 1e-9          # halving dt roughly halves the drift
    assert drift[1][0] < 5e-4 and drift[1][1] < 1e-3 and drift[1][2] < 1e-3


def test_free_fall_momentum():
    o = DynOracle(2, DynParams(contacts=False, pd=False))
    o.root_pos[:, 2] += 5
    for _ in range(100):
        o.substep(o.q)
    _, P, _ = o.energy_momentum()
    mass = o.m.body_mass.sum()
    assert np.allclose(P[:, 2], -mass * o.P.gravity * 100 * o.P.dt, rtol=1e-9)
    assert np.allclose(P[:, :2], 0, atol=1e-12)


def test_static_stand_supports_weight_at_reference_height():
    o = DynOracle(1)
    for _ in range(600):
        o.substep(o.m.default_joint_pos[None])
    ls = o.link_state()
    assert ls["body_link_pos"][0, 6, 2] == pytest.approx(0.2545, abs=5e-4)   # reference prints 0.2545
    mg = o.m.body_mass.sum() * o.P.gravity
    assert o.body_force[0, :, 2].sum() == pytest.approx(mg, rel=1e-3)
    assert np.abs(o.body_force[0, 1:6]).max() == 0.0                         # only the feet touch
    assert abs(o.body_force[0, 0, 2] - o.body_force[0, 6, 2]) < 0.2 * mg


def test_port_f64_matches_independent_oracle():
    """The kernel's arithmetic (zbot_core.h, T=double, ABA in spatial algebra) equals the dense
    classical-Jacobian oracle to round-off, including contact and PD, over 40 substeps."""
    from oracle import cpu_port
    from zbot_lab_b200.utils import synthetic as syn
    n = 48
    rng = np.random.default_rng(3)
    st = syn.synth_sim_state(rng, n)
    st["root_ang_vel"] = rng.normal(0, 0.5, (n, 3)).astype(np.float32)
    o = DynOracle(n)
    o.set_state({k: v.astype(np.float64) for k, v in st.items()})
    sim = cpu_port.pack_sim(st, np.float64)
    tgt = o.m.default_joint_pos[None] + rng.uniform(-0.3, 0.3, (n, 6))
    for i in range(40):
        f, tau = cpu_port.substeps(sim, tgt, 1)
        o.substep(tgt)
        ref = np.concatenate([o.root_pos, o.root_quat, o.root_lin_vel, o.root_ang_vel, o.q, o.qd], -1)
        if i < 10:
            assert np.abs(sim - ref).max() < 1e-9, i
            assert np.abs(f[:, [0, 6]] - o.body_force[:, [0, 6]]).max() < 1e-7
            assert np.abs(tau - o.applied_torque).max() < 1e-9
    assert np.abs(sim - ref).max() < 1e-5       # round-off amplified by contact dynamics only


def test_port_mid_body_contact_matches_oracle():
    """A robot lying on its side: the merged-body spheres touch the ground (termination path)."""
    from oracle import cpu_port
    n = 4
    o = DynOracle(n)
    st = {"root_pos": np.tile([0.0, 0.0, 0.051], (n, 1)), "root_quat": np.tile([0.7071067811865476, 0.0, 0.7071067811865476, 0.0], (n, 1)),
          "root_lin_vel": np.zeros((n, 3)), "root_ang_vel": np.zeros((n, 3)),
          "joint_pos": np.zeros((n, 6)), "joint_vel": np.zeros((n, 6))}
    st["root_pos"][:, 2] += np.linspace(-0.002, 0.002, n)
    o.set_state(st)
    sim = cpu_port.pack_sim(st, np.float64)
    tgt = np.zeros((n, 6))
    touched = False
    for i in range(20):
        f, _ = cpu_port.substeps(sim, tgt, 1)
        o.substep(tgt)
        assert np.abs(f[:, 1:6] - o.body_force_pred[:, 1:6]).max() < 1e-6
        touched |= bool(np.abs(f[:, 1:6]).max() > 1.0)
        ref = np.concatenate([o.root_pos, o.root_quat, o.root_lin_vel, o.root_ang_vel, o.q, o.qd], -1)
        assert np.abs(sim - ref).max() < 1e-8
    assert touched


def test_full_step_oracle_matches_port_f64_with_resets():
    from oracle import cpu_port
    from oracle.full_step_oracle import FullStepOracle
    n = 24
    rng = np.random.default_rng(0)
    fo = FullStepOracle(n)
    fo.reset_all()
    pe = cpu_port.PortEnv(n, np.float64)
    fo.mdp.episode_length_buf[:4] = 990
    pe.ep_len[:4] = 990
    resets = 0
    for t in range(40):
        a = rng.normal(0, 0.6, (n, 6)).astype(np.float32)
        obs, rew, term, trunc, ids, log = fo.step(a)
        o2, r2, t2, tr2, rs, _ = pe.step(a)
        assert np.array_equal(term, t2) and np.array_equal(trunc, tr2)
        assert np.array_equal(fo.mdp.episode_length_buf, pe.ep_len)
        assert np.abs(obs - o2).max() < 5e-5 and np.abs(rew - r2).max() < 5e-5
        resets += len(ids)
    assert resets >= 4


def test_port_f32_tracks_f64_over_50_steps():
    """Stated horizon tolerance holds for the float32 instantiation on the CPU."""
    from oracle import cpu_port
    n = 256
    rng = np.random.default_rng(5)
    e32, e64 = cpu_port.PortEnv(n, np.float32), cpu_port.PortEnv(n, np.float64)
    alive = np.ones(n, bool)
    for t in range(50):
        a = rng.normal(0, 0.3, (n, 6)).astype(np.float32)
        _, _, t32, tr32, _, _ = e32.step(a)
        _, _, t64, tr64, _, _ = e64.step(a)
        alive &= ~(t32 | t64 | tr32 | tr64)
    dq = np.abs(e32.field("joint_pos", 6) - e64.field("joint_pos", 6))[alive]
    assert dq.max() < 5e-3 and np.median(dq.max(1)) < 1e-4


# ------------------------------------------------------------------------------------------------ snake task
def _snake_speed(rng, n):
    return ((rng.random(n) * 1.8 + 0.2) * np.pi).astype(np.float32)        # snake_v0.py:121


def test_snake_full_step_oracle_matches_port_f64_with_resets():
    """BASELINE.json configs[3]: composed oracle (independent float64 dynamics of the snake model + the
    reference-pinned snake MDP restatement) against the kernel's own arithmetic compiled for the host."""
    from oracle import cpu_port
    from oracle.full_step_oracle import SnakeFullStepOracle
    from zbot_lab_b200 import native
    n = 24
    rng = np.random.default_rng(0)
    speed = _snake_speed(rng, n)
    fo = SnakeFullStepOracle(n, speed)
    fo.reset_all()
    pe = cpu_port.PortEnv(n, np.float64, native.make_cfg(n, task=native.TASK_SNAKE_V0))
    pe.field("joint_speed_limit", 1)[:, 0] = speed
    fo.mdp.episode_length_buf[:4] = 790
    pe.ep_len[:4] = 790
    resets = 0
    for t in range(30):
        a = rng.normal(0, 1.0, (n, 6)).astype(np.float32)
        obs, rew, term, trunc, ids, _ = fo.step(a)
        o2, r2, t2, tr2, _, _ = pe.step(a)
        assert np.array_equal(term, t2) and np.array_equal(trunc, tr2)
        assert np.array_equal(fo.mdp.episode_length_buf, pe.ep_len)
        assert np.abs(obs - o2).max() < 1e-4 and np.abs(rew - r2).max() < 1e-5
        assert np.all(o2[:, 22] == speed)                       # survives resets
        resets += len(ids)
    assert resets >= 4


def test_snake_self_contact_and_x_drift_terminate():
    """Curl the chain so the end links overlap (filtered self-contact proxy > 1 N), and push the base off
    x = -0.318 m by more than 0.2 m: both terminations of snake_v0.py:222-240 fire, with the -20 penalty."""
    from oracle import cpu_port
    from zbot_lab_b200 import native
    n = 4
    pe = cpu_port.PortEnv(n, np.float64, native.make_cfg(n, task=native.TASK_SNAKE_V0))
    pe.field("joint_speed_limit", 1)[:] = np.pi
    pe.field("root_pos", 3)[1, 0] += 0.35                       # env 1: base_pos_x_err = 0.35
    a = np.zeros((n, 6), np.float32)
    _, rew, term, trunc, _, _ = pe.step(a)                      # stale base_pos is the START-of-step one
    assert term.tolist() == [False, True, False, False] and not trunc.any()
    assert rew[1] < -19.0 and abs(rew[0]) < 1.0
    # env 2: q = 3 rad on every joint folds the chain onto itself (sphere centres 0.04 m apart)
    pe.field("joint_pos", 6)[2] = 3.0
    pe.field("p_delta", 6)[2] = 3.0                             # PD target = the folded pose
    pe.field("root_pos", 3)[2, 2] += 0.3
    _, _, term, _, _, ex = pe.step(a, export=True)
    assert term[2] and not term[0]
    assert ex[2, 22] > 1.0 and ex[0, 22] == 0.0                 # exported self-contact force proxy


# ------------------------------------------------------------------------------------------------ walking v4
def _v4_port(n, dtype, rng):
    from oracle import cpu_port
    from zbot_lab_b200 import native
    pe = cpu_port.PortEnv(n, dtype, native.make_cfg(n, task=native.TASK_WALKING_V4))
    pe.field("feet_contact_forces_last", 2)[:] = 15.0                                   # …env_v4.py:637
    pe.field("carry_feet_fz", 2)[:] = np.stack([rng.uniform(-0.3, 0.3, n), rng.uniform(-0.1, 0.1, n)], -1)   # commands
    pe.field("carry_mid_max", 1)[:, 0] = rng.uniform(-3, 3, n)                          # target_heading_yaw
    tl = rng.uniform(3.0, 6.0, n)
    tl[: n // 4] = rng.integers(1, 12, n // 4) * 0.02 - 0.01                            # interval events inside the test
    pe.field("base_pos_y_err_sum", 1)[:, 0] = tl
    return pe


def test_v4_full_step_port_matches_pinned_oracle_on_exported_physics():
    """the v4 step's dones / rewards / command resampling / randomised resets / observations (kernel
    arithmetic, host build, float32) equal the reference-pinned v4 oracle evaluated on the articulation + contact
    view the step itself produced.  Resample masks, reset ids, counters bit-exact; floats <= 1e-5 relative."""
    from helpers import make_v4_oracle, v4_check_step
    n = 96
    rng = np.random.default_rng(42)
    pe = _v4_port(n, np.float32, rng)
    ep0 = rng.integers(0, 1000, n)
    ep0[:6] = 995
    pe.ep_len[:] = ep0
    o = make_v4_oracle(n, np.zeros((n, 3), np.float32))
    o.episode_length_buf[:] = ep0
    o.commands[:] = pe.field("carry_feet_fz", 2)
    o.target_heading_yaw[:] = pe.field("carry_mid_max", 1)[:, 0]
    o.interval_time_left[:] = pe.field("base_pos_y_err_sum", 1)[:, 0]
    o.feet_down_pos_last[:] = pe.field("feet_down_pos_last", 6).reshape(n, 2, 3)
    n_reset = n_int = n_term = 0
    for t in range(40):
        a = rng.normal(0, 1.0, (n, 6)).astype(np.float32)
        rnd = rng.random((n, 10)).astype(np.float32)
        obs, rew, term, trunc, rs, ex = pe.step(a, export=True, rnd=rnd)
        ids, iv, log = v4_check_step(o, a, rnd, ex, obs, rew, term, trunc, pe.ep_len,
                                     lambda k: pe.field(k, {"carry_feet_fz": 2, "carry_mid_max": 1, "base_pos_y_err_sum": 1,
                                                            "p_delta": 6, "feet_step_length": 2, "feet_contact_forces_last": 2,
                                                            "feet_down_pos_last": 6, "episode_sums": 16}[k]))
        if len(ids) > 0:
            for i, nm in enumerate(o.episode_sums):
                want = float(log["Episode_Reward/" + nm])
                got = float(np.mean(rs[ids, i], dtype=np.float32))
                assert abs(got - want) <= 1e-5 * max(1.0, abs(want)), nm
        n_reset += len(ids)
        n_int += len(iv)
        n_term += int(term.sum())
    assert n_reset >= 6 and n_int >= n // 4 and n_term > 0


def test_v4_port_f32_tracks_f64_and_resets_randomise_the_pose():
    n = 128
    rng = np.random.default_rng(7)
    e32, e64 = _v4_port(n, np.float32, np.random.default_rng(1)), _v4_port(n, np.float64, np.random.default_rng(1))
    for e in (e32, e64):
        e.ep_len[:8] = 990
    alive = np.ones(n, bool)
    for t in range(30):
        a = rng.normal(0, 0.3, (n, 6)).astype(np.float32)
        rnd = rng.random((n, 10)).astype(np.float32)
        _, _, t32, tr32, _, _ = e32.step(a, rnd=rnd)
        _, _, t64, tr64, _, _ = e64.step(a, rnd=rnd)
        assert np.array_equal(tr32, tr64)
        if t == 8:                                    # 990 + 9 = 999: the first 8 envs time out and get a random pose
            assert tr32[:8].all()
            p = e32.field("root_pos", 3)[:8]
            want = np.array([0.0, -0.06]) + (rnd[:8, :2] - 0.5)
            assert np.allclose(p[:, :2], want, atol=1e-6)
            yaw = rnd[:8, 2] * 6.28 - 3.14
            assert np.allclose(e32.field("root_quat", 4)[:8, 0], np.cos(yaw / 2), atol=1e-6)
            assert np.allclose(e32.field("base_heading_x_sum", 1)[:8, 0], yaw, atol=1e-6)        # current_yaw
        alive &= ~(t32 | t64 | tr32 | tr64)
    dq = np.abs(e32.field("joint_pos", 6) - e64.field("joint_pos", 6))[alive]
    assert alive.sum() > n // 3 and dq.max() < 5e-3


def test_two_lane_instantiation_equals_scalar_bit_for_bit():
    """physics_substep instantiated with T = F2 (two environments per call chain, the packed GPU kernel's arithmetic;
    csrc/zbot_pair.h) gives, lane by lane, exactly the scalar float32 result on the host -- walking and snake models,
    contacts included (the lane-generic contact law contributes exact zeros for an inactive lane)."""
    from oracle import cpu_port
    from zbot_lab_b200 import native
    from zbot_lab_b200.assets import zbot_d_6s as S
    from zbot_lab_b200.utils import synthetic as syn
    n = 64
    rng = np.random.default_rng(3)
    st = syn.synth_sim_state(rng, n)
    st["root_ang_vel"] = rng.normal(0, 0.5, (n, 3)).astype(np.float32)
    st["root_pos"][:8, 2] += 0.3                       # some pairs have one lane airborne, one in contact
    a = cpu_port.pack_sim(st, np.float32)
    b = a.copy()
    tgt = (np.asarray([0.312, 0.837, -2.02, 2.02, -0.837, -0.312], np.float32)[None] + rng.uniform(-0.3, 0.3, (n, 6))).astype(np.float32)
    for _ in range(24):
        f1, t1 = cpu_port.substeps(a, tgt, 1)
        f2, t2 = cpu_port.substeps_pair(b, tgt, 1)
        assert np.array_equal(a, b) and np.array_equal(f1, f2) and np.array_equal(t1, t2)
    m = S.model_f32()
    cfg = native.make_cfg(n, task=native.TASK_SNAKE_V0)
    sim = np.zeros((n, 25), np.float32)
    sim[:, 0:3], sim[:, 3:7] = m.default_root_pos, m.default_root_quat
    sim[:, 13:19], sim[:, 19:25] = rng.uniform(-0.5, 0.5, (n, 6)), rng.normal(0, 1, (n, 6))
    a, b = sim.copy(), sim.copy()
    tgt = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
    for _ in range(24):
        cpu_port.substeps(a, tgt, 1, cfg, snake=True)
        cpu_port.substeps_pair(b, tgt, 1, cfg, snake=True)
        assert np.array_equal(a, b)


# ------------------------------------------------------------------------------------------------ manager-based task
def _m_port(n, dtype, rng, terms, P=None):
    from helpers import M_PARAMS, m_native_cfg
    from oracle import cpu_port
    pe = cpu_port.PortEnv(n, dtype, m_native_cfg(n, terms, P or M_PARAMS))
    pe.field("carry_feet_fz", 2)[:] = np.stack([rng.uniform(-0.3, 0.3, n), rng.uniform(-0.1, 0.1, n)], -1)
    pe.field("carry_mid_max", 1)[:, 0] = rng.uniform(-0.2, 0.2, n)
    pe.field("base_pos_y_err_sum", 1)[:, 0] = rng.uniform(0.05, 0.3, n)          # command time_left
    pe.field("joint_speed_limit", 1)[:, 0] = rng.uniform(0.3, 1.0, n)           # per-env friction
    pe.field("joint_pos", 6)[:] += rng.uniform(-0.1, 0.1, (n, 6))
    return pe


@pytest.mark.parametrize("which", ["flat", "all"])
def test_m_full_step_port_matches_pinned_oracle_on_exported_physics(which):
    """the manager task's terminations / rewards / command resampling / randomised resets / observations
    (kernel arithmetic, host build, float32) equal the reference-pinned oracle evaluated on the view the step itself
    produced.  Flags, reset ids, counters, standing masks bit-exact; floats <= 1e-5 relative.  `flat` = the 11 terms of
    Zbot6BFlatEnvCfg, `all` = every RewTerm function the reference's rewards.py defines (16 slots + is_terminated)."""
    from helpers import M_ALL_TERMS, m_check_step, m_make_oracle
    from zbot_lab_b200 import native
    terms = M_ALL_TERMS if which == "all" else native.M_FLAT_TERMS
    n = 96
    rng = np.random.default_rng(42)
    pe = _m_port(n, np.float32, rng, terms)
    ep0 = rng.integers(0, 990, n)
    ep0[:6] = 996
    pe.ep_len[:] = ep0
    get = lambda k, w: pe.field(k, w).copy()
    o = m_make_oracle(n, terms, get, ep0)
    n_reset = n_term = n_res = 0
    for t in range(40):
        a = rng.normal(0, 1.5, (n, 6)).astype(np.float32)
        rnd = rng.random((n, 22)).astype(np.float32)
        obs, rew, term, trunc, rs, ex = pe.step(a, export=True, rnd=rnd)
        r = m_check_step(o, a, rnd, ex, obs, rew, term, trunc, pe.ep_len, get, rs)
        n_reset += len(r["reset_ids"])
        n_term += int(term.sum())
        n_res += len(r["resample_ids"])
    assert n_reset >= 6 and n_term > 0 and n_res >= n


def test_m_extra_cfg_features_port_matches_oracle():
    """The cfg features `ZbotLabRoughEnvCfg` defines and the registered cfgs switch off, all switched on: RewTerm
    `undesired_contacts` (zbotlab_env_cfg.py:367-371), DoneTerm `base_contact` = illegal_contact on `base` (:385-388),
    `heading_command=True` (:86-97) and the interval EventTerm `push_robot` (:253-258).  All four are Isaac Lab functions the
    reference does not vendor ([IL-upstream], unpinned): the oracle restates them from upstream knowledge, the kernel
    arithmetic (host build) must agree with it -- flags / masks / timers exact, floats <= 1e-5; a push adds exactly the drawn
    velocity to the whole robot."""
    from helpers import M_EXTRA_TERMS, M_PARAMS_EXTRA, m_check_step, m_make_oracle
    n = 128
    rng = np.random.default_rng(7)
    pe = _m_port(n, np.float32, rng, M_EXTRA_TERMS, M_PARAMS_EXTRA)
    twin = _m_port(n, np.float32, np.random.default_rng(7), M_EXTRA_TERMS, dict(M_PARAMS_EXTRA, push=None))
    pd = pe.field("p_delta", 6)
    pd[:, 0], pd[:, 1], pd[:, 2] = rng.uniform(-3, 3, n), rng.random(n) < 0.7, rng.uniform(0.02, 0.4, n)
    # a third of the robots start lying on their side: merged bodies on the ground -> undesired / illegal contacts
    lying = np.arange(n) % 3 == 0
    pe.field("root_pos", 3)[lying, 2] = 0.0485
    pe.field("root_quat", 4)[lying] = np.array([0.7071068, 0.0, 0.7071068, 0.0], np.float32)
    pe.field("joint_pos", 6)[lying] = 0.0
    ep0 = rng.integers(0, 990, n)
    pe.ep_len[:] = ep0
    get = lambda k, w: pe.field(k, w).copy()
    o = m_make_oracle(n, M_EXTRA_TERMS, get, ep0, M_PARAMS_EXTRA)
    n_ill = n_und = n_push = n_head = 0
    for t in range(30):
        a = rng.normal(0, 1.0, (n, 6)).astype(np.float32)
        rnd = rng.random((n, 22)).astype(np.float32)
        twin.state[:] = pe.state
        twin.ep_len[:] = pe.ep_len
        obs, rew, term, trunc, rs, ex = pe.step(a, export=True, rnd=rnd)
        twin.step(a, rnd=rnd)
        r = m_check_step(o, a, rnd, ex, obs, rew, term, trunc, pe.ep_len, get, rs)
        # the push: the twin (same state, same uniforms, no push_robot term) differs by exactly the drawn velocity
        dv = pe.field("root_lin_vel", 3) - twin.field("root_lin_vel", 3)
        assert np.abs(dv[:, :2] - o.push_dv).max() <= 1e-6 and np.abs(dv[:, 2]).max() == 0
        n_push += int((np.abs(o.push_dv).sum(1) > 0).sum())
        n_ill += int(r["illegal"].sum())
        n_und += int((r["values"]["undesired_contacts"] > 0).sum())
        n_head += int((o.is_heading & ~o.standing & (np.abs(o.cmd[:, 2]) > 0)).sum())
    assert n_ill > 0 and n_und > n_ill and n_push > n // 2 and n_head > n


def test_m_port_f32_tracks_f64_and_the_robot_stands():
    """float32 vs float64 host builds of the manager step over 50 control steps of small random actions; with zero
    actions the biped stands at the cfg's init height carrying m g on its two soles."""
    from zbot_lab_b200 import native
    n = 64
    rng = np.random.default_rng(5)
    e32, e64 = (_m_port(n, dt, np.random.default_rng(1), native.M_FLAT_TERMS) for dt in (np.float32, np.float64))
    for e in (e32, e64):
        e.field("base_pos_y_err_sum", 1)[:] = 100.0
        e.cfg.feet_close_min = 0.0          # the init stance is 0.12002 m wide against the 0.12 m limit: keep the horizon alive
    alive = np.ones(n, bool)
    for t in range(50):
        a = rng.normal(0, 0.3, (n, 6)).astype(np.float32)
        rnd = rng.random((n, 22)).astype(np.float32)
        _, _, t32, tr32, _, _ = e32.step(a, rnd=rnd)
        _, _, t64, tr64, _, _ = e64.step(a, rnd=rnd)
        alive &= ~(t32 | t64 | tr32 | tr64)
    dq = np.abs(e32.field("joint_pos", 6) - e64.field("joint_pos", 6))[alive]
    # kp 20 / kd 0.5 (ZBOT_6S_V2_CFG) is a softer, less damped drive than the direct tasks' 50 / 5: round-off grows faster
    assert alive.sum() > n // 3 and dq.max() < 5e-2 and np.median(dq) < 1e-4 and np.quantile(dq, 0.9) < 2e-3
    from oracle import cpu_port
    from oracle.m_mdp_oracle import split_view
    pe = cpu_port.PortEnv(4, np.float64, e64.cfg.__class__.from_buffer_copy(e64.cfg))
    pe.cfg.num_envs = 4
    pe.field("base_pos_y_err_sum", 1)[:] = 100.0
    for t in range(40):
        obs, rew, term, trunc, rs, ex = pe.step(np.zeros((4, 6)), export=True, rnd=np.full((4, 22), 0.5))
    v = split_view(ex)
    assert abs(v["root_pos"][0, 2] - 0.2545) < 1e-3 and abs(v["root_quat"][0, 0]) > 0.99999
    assert np.abs(v["feet_fz_hist"][0] - 0.5 * 3.00504 * 9.81).max() < 0.05          # each sole carries m g / 2


def test_m_model_known_answers_and_independent_dynamics():
    """ZBOT_6S_V2_CFG (assets/zbot_cfg.py:959-1005) on the decoded zbot_6s_v09.usd: with the base link at (0, 0, 0.2545),
    identity rotation, and the cfg's joint angles both soles lie flat on the ground (known answer: the init height),
    the feet's link y axes point up / down and x forward (the axes rewards.py:113-115, 134-136 hard-code); the full-inertia
    ABA of the kernel (host build, double) equals the dense classical-Jacobian oracle to round-off with contact, PD and
    friction 0.7."""
    from oracle import cpu_port
    from oracle.dyn_oracle import DynOracle, DynParams
    from helpers import m_native_cfg
    from zbot_lab_b200 import native
    from zbot_lab_b200.assets import zbot_6s as Z
    from zbot_lab_b200.assets import zbot_6s_v2 as V
    m = V.model_f32()
    lp, lq = V.default_link_poses()
    ib, i0, i1 = (V.link_index(k) for k in ("base", "foot0", "foot1"))
    assert np.allclose(lp[ib], [0, 0, 0.2545], atol=1e-6) and np.allclose(np.abs(lq[ib]), [1, 0, 0, 0], atol=1e-6)
    R0 = Z.quat_to_mat(m.default_root_quat)
    assert abs(m.default_root_pos[2]) < 2e-4 and R0[2, 2] > 0.99999                       # foot0 sole on the ground, flat
    up0, up1 = Z.quat_rotate(lq[i0], np.array([0.0, 1, 0])), Z.quat_rotate(lq[i1], np.array([0.0, -1, 0]))
    assert up0[2] > 0.99999 and up1[2] > 0.99999
    assert Z.quat_rotate(lq[i0], np.array([1.0, 0, 0]))[0] > 0.99999 and Z.quat_rotate(lq[i1], np.array([1.0, 0, 0]))[0] > 0.99999
    assert abs(lp[i0][1] + 0.06) < 1e-4 and abs(lp[i1][1] - 0.06) < 1e-4 and abs(lp[i1][2] - lp[i0][2]) < 1e-6
    assert abs(m.body_mass.sum() - 12 * 0.25042) < 1e-6
    n = 48
    rng = np.random.default_rng(3)
    cfg = m_native_cfg(n, native.M_FLAT_TERMS, friction=0.7)
    st = {"root_pos": np.tile(m.default_root_pos, (n, 1)) + np.c_[np.zeros((n, 2)), rng.uniform(0, 0.02, n)],
          "root_lin_vel": rng.normal(0, 0.1, (n, 3)), "root_ang_vel": rng.normal(0, 0.5, (n, 3)),
          "joint_pos": m.default_joint_pos[None] + rng.uniform(-0.2, 0.2, (n, 6)), "joint_vel": rng.normal(0, 0.5, (n, 6))}
    yaw = rng.uniform(-np.pi, np.pi, n)
    st["root_quat"] = Z.quat_mul(np.stack([np.cos(yaw / 2), 0 * yaw, 0 * yaw, np.sin(yaw / 2)], -1), np.tile(m.default_root_quat, (n, 1)))
    st = {k: np.float32(v).astype(np.float64) for k, v in st.items()}
    o = DynOracle(n, DynParams(m, mu=float(np.float32(0.7))), model=m)
    o.set_state(st)
    sim = cpu_port.pack_sim(st, np.float64)
    tgt = m.default_joint_pos[None] + rng.uniform(-0.3, 0.3, (n, 6))
    for i in range(30):
        f, tau = cpu_port.substeps(sim, tgt, 1, cfg, model="m")
        o.substep(tgt)
        ref = np.concatenate([o.root_pos, o.root_quat, o.root_lin_vel, o.root_ang_vel, o.q, o.qd], -1)
        if i < 10:
            assert np.abs(sim - ref).max() < 1e-9 and np.abs(f[:, [0, 6]] - o.body_force[:, [0, 6]]).max() < 1e-7
    assert np.abs(sim - ref).max() < 1e-5


def test_model_constants_match_the_references_usd_files():
    """Build container only: the constants embedded in assets/zbot_6s.py / zbot_6s_v2.py against the reference's binary USD
    crates, decoded with tools/usdc_dump.py -- link masses / CoMs / principal inertias, link placements, joint frames and axes,
    articulation roots (zbot_6s_new.usd: SURVEY App. A; zbot_6s_v09.usd: the manager-task robot)."""
    import os
    import sys
    ref = "/root/reference/source/zbot/zbot/assets/zbot_assets"
    if not os.path.isfile(os.path.join(ref, "zbot_6s_v09.usd")):
        pytest.skip("/root/reference is not present (GPU box)")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import usdc_dump as U
    from zbot_lab_b200.assets import zbot_6s as Z
    from zbot_lab_b200.assets import zbot_6s_v2 as V

    def load(name):
        c = U.Crate(os.path.join(ref, name))
        specs = {p: fs for p, fs, st in c.specs}
        return c, (lambda path, field="default": c.value(c.spec_fields(specs[path])[field]))

    # ---- zbot_6s_new.usd (walking robot)
    c, val = load("zbot_6s_new.usd")
    assert "PhysicsArticulationRootAPI" in val("/zbot/foot_0", "apiSchemas")["explicit"]
    for k, name in enumerate(Z.LINK_NAMES):
        assert np.allclose(val(f"/zbot/{name}.xformOp:translate"), (0, 0, Z.LINK_SPACING * k), atol=1e-6)
        assert abs(val(f"/zbot/{name}.physics:mass") - Z.LINK_MASS) < 1e-6
        com, diag, axes = (Z.A_COM, Z.A_DIAG_INERTIA, Z.A_PRINCIPAL_AXES) if name in Z.A_TYPE_LINKS else (Z.B_COM, Z.B_DIAG_INERTIA, Z.B_PRINCIPAL_AXES)
        assert np.allclose(val(f"/zbot/{name}.physics:centerOfMass"), com, atol=1e-7)
        assert np.allclose(val(f"/zbot/{name}.physics:diagonalInertia"), diag, rtol=1e-5)
        assert np.allclose(val(f"/zbot/{name}.physics:principalAxes"), axes, atol=1e-6)
    parents = ("foot_0", "a2", "a3", "base", "a5", "a6")
    for k, (par, sgn) in enumerate(zip(parents, Z.JOINT_AXIS_SIGN)):
        j = f"/zbot/{par}/joint{k + 1}"
        assert val(j + ".physics:axis") == "Z" and np.allclose(val(j + ".physics:localPos0"), (0, 0, 0.053), atol=1e-6)
        axis = Z.quat_rotate(np.array(val(j + ".physics:localRot0"), float), np.array([0.0, 0, 1]))
        assert np.allclose(axis, (sgn * Z.SIN45, 0, Z.SIN45), atol=1e-6)
    # ---- zbot_6s_v09.usd (manager-task robot)
    c, val = load("zbot_6s_v09.usd")
    assert "PhysicsArticulationRootAPI" in val("/zbot/base", "apiSchemas")["explicit"]
    usd_name = {n: n for n in V.LINK_NAMES}
    for name in V.LINK_NAMES:
        assert np.allclose(val(f"/zbot/{name}.xformOp:translate"), (0, V.LINK_Y[name], 0), atol=1e-6)
        q = np.array(val(f"/zbot/{name}.xformOp:orient"), float)
        want = np.array(V.BASE_LINK_QUAT) if name == "base" else np.array([1.0, 0, 0, 0])
        assert np.allclose(q, want, atol=2e-6)
        com, diag = (Z.A_COM, Z.A_DIAG_INERTIA) if name in V.A_TYPE_LINKS else (Z.B_COM, Z.B_DIAG_INERTIA)
        assert np.allclose(val(f"/zbot/{name}.physics:centerOfMass"), com, atol=1e-7)
        assert np.allclose(val(f"/zbot/{name}.physics:diagonalInertia"), diag, rtol=1e-5)
        assert abs(val(f"/zbot/{name}.physics:mass") - Z.LINK_MASS) < 1e-6
    # joints: position on the y axis, axis = joint-frame Y rotated by localRot1 (the child link frames are identity-oriented)
    joint_parent = {"joint1": "base", "joint2": "a2", "joint3": "a3", "joint7": "a7", "joint8": "a8", "joint9": "a9"}
    C = V.CHAIN_IN_WORLD
    m = V.model_f32()
    for k, jn in enumerate(V.CHAIN_JOINTS):
        j = f"/zbot/{joint_parent[jn]}/{jn}"
        assert val(j + ".physics:axis") == "Y"
        child = val(j + ".physics:body1", "targetPaths")["explicit"][0].split("/")[-1]
        assert abs(V.LINK_Y[child] - V.JOINT_Y[k]) < 1e-6 and np.allclose(val(j + ".physics:localPos1"), 0, atol=1e-7)
        axis_w = Z.quat_rotate(np.array(val(j + ".physics:localRot1"), float), np.array([0.0, 1, 0]))   # in the child = world frame
        axis_c = C.T @ axis_w
        # the three joints below the base are traversed child -> parent by the chain: the axis vector flips, the angle does not
        flip = -1.0 if jn in ("joint1", "joint2", "joint3") else 1.0
        assert np.allclose(flip * axis_c, m.joint_axis[k], atol=1e-6), (jn, axis_c, m.joint_axis[k])
        for lim in ("lowerLimit", "upperLimit"):
            assert abs(abs(val(j + ".physics:" + lim)) - 360.0) < 1e-6


def test_snake_model_constants_match_the_references_usd_file():
    """Build container only: assets/zbot_d_6s.py (derived from the ASCII sibling zbot_6s_v04.usda) against the binary
    zbot_6s_v03.usd the snake task actually names (assets/zbot_cfg.py:109-168): root a1, link placements and the link-frame
    rotations that the per-body mass tables are rotated with, revolute frames, mass properties."""
    import os
    import sys
    ref = "/root/reference/source/zbot/zbot/assets/zbot_assets"
    if not os.path.isfile(os.path.join(ref, "zbot_6s_v03.usd")):
        pytest.skip("/root/reference is not present (GPU box)")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import usdc_dump as U
    from zbot_lab_b200.assets import zbot_6s as Z
    from zbot_lab_b200.assets import zbot_d_6s as S
    c = U.Crate(os.path.join(ref, "zbot_6s_v03.usd"))
    specs = {p: fs for p, fs, st in c.specs}
    val = lambda path, field="default": c.value(c.spec_fields(specs[path])[field])
    assert "PhysicsArticulationRootAPI" in val("/zbot/a1", "apiSchemas")["explicit"]
    for name in S.LINK_NAMES:
        k = int(name[1])
        z = 0.106 * (k - 1) + (0.053 if name[0] == "b" else 0.0)
        assert np.allclose(val(f"/zbot/{name}.xformOp:translate"), (0, 0, z), atol=1e-6), name
        q, want = np.array(val(f"/zbot/{name}.xformOp:orient"), float), S._link_rot(name)
        assert min(np.abs(q - want).max(), np.abs(q + want).max()) < 2e-6, (name, q, want)
        com, diag = (Z.A_COM, Z.A_DIAG_INERTIA) if name[0] == "a" else (Z.B_COM, Z.B_DIAG_INERTIA)
        assert np.allclose(val(f"/zbot/{name}.physics:centerOfMass"), com, atol=1e-7)
        assert np.allclose(val(f"/zbot/{name}.physics:diagonalInertia"), diag, rtol=1e-5)
        assert abs(val(f"/zbot/{name}.physics:mass") - Z.LINK_MASS) < 1e-6
    for k in range(1, 7):
        j = f"/zbot/a{k}/joint{k}"
        assert val(j + ".physics:axis") == "Z" and np.allclose(val(j + ".physics:localPos0"), (0, 0, 0.053), atol=1e-6)
        assert np.allclose(val(j + ".physics:localRot0"), (0.92388, 0, 0.38268, 0), atol=2e-6)
        assert val(j + ".physics:body1", "targetPaths")["explicit"] == [f"/zbot/b{k}"]


def test_sole_geometry_matches_the_collision_meshes():
    """Build container only: the foot soles of the contact model (flat disc, radius 0.05 m, at the link-frame offsets the
    models use) against the flat faces of the reference's convex-hull collision meshes (`/zbot/<foot>/collisions.points`
    transformed by the mesh xform): zbot_6s_new.usd foot_0 z = 0 / foot_1 z = +0.053, zbot_6s_v09.usd foot0 y = -0.053 /
    foot1 y = +0.053."""
    import os
    import struct
    import sys
    ref = "/root/reference/source/zbot/zbot/assets/zbot_assets"
    if not os.path.isfile(os.path.join(ref, "zbot_6s_v09.usd")):
        pytest.skip("/root/reference is not present (GPU box)")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import usdc_dump as U
    from zbot_lab_b200.assets import zbot_6s as Z
    from zbot_lab_b200.assets import zbot_6s_v2 as V

    def flat_faces(fname, link):
        c = U.Crate(os.path.join(ref, fname))
        specs = {p: fs for p, fs, st in c.specs}
        rep = c.spec_fields(specs[f"/zbot/{link}/collisions.points"])["default"]
        off = rep & ((1 << 48) - 1)
        (n,) = struct.unpack_from("<Q", c.d, off)
        P = np.frombuffer(c.d, dtype="<f4", count=3 * n, offset=off + 8).reshape(n, 3).astype(np.float64)
        q = c.value(c.spec_fields(specs[f"/zbot/{link}/collisions.xformOp:orient"])["default"])
        Q = P @ Z.quat_to_mat(q).T
        out = []
        for ax in range(3):
            for v in (Q[:, ax].min(), Q[:, ax].max()):
                sel = np.abs(Q[:, ax] - v) < 1e-4
                if sel.sum() > 50:
                    pts = Q[sel]
                    out.append((ax, float(v), float(np.linalg.norm(pts - pts.mean(0), axis=1).max())))
        return out

    for fname, link, ax, pos in (("zbot_6s_new.usd", "foot_0", 2, Z.FOOT0_SOLE_Z), ("zbot_6s_new.usd", "foot_1", 2, Z.FOOT1_SOLE_Z),
                                 ("zbot_6s_v09.usd", "foot0", 1, V.SOLE_Y[0] - V.LINK_Y["foot0"]),
                                 ("zbot_6s_v09.usd", "foot1", 1, V.SOLE_Y[1] - V.LINK_Y["foot1"])):
        faces = flat_faces(fname, link)
        assert len(faces) == 1, (fname, link, faces)                 # each half-module has exactly one flat circular end
        a, v, r = faces[0]
        assert a == ax and abs(v - pos) < 2e-4 and abs(r - Z.FOOT_DISC_RADIUS) < 1.5e-3, (fname, link, faces)


def test_halves_elimination_equals_the_one_chain_substep():
    """csrc/zbot_halves.h (the arithmetic of the two-warps-per-32-envs GPU kernel): the articulated-body elimination run from
    BOTH feet towards body 3 -- side A through the reversed joints 0..2, side B through joints 5..3, one 6x6 solve at body
    3 -- is the same substep as the one-chain `physics_substep`: identical model, identical discretisation.  In float64 the
    two agree to round-off over 8 substeps with contact, PD saturation and random targets; in float32 both sit at the same
    distance from the float64 result."""
    from oracle import cpu_port
    from zbot_lab_b200.utils import synthetic as syn
    n = 600
    rng = np.random.default_rng(3)
    st = syn.synth_sim_state(rng, n)
    keys = ("root_pos", "root_quat", "root_lin_vel", "root_ang_vel", "joint_pos", "joint_vel")
    sim = np.concatenate([st[k].reshape(n, -1) for k in keys], 1).astype(np.float64)
    tgt = st["joint_pos"].astype(np.float64) + rng.normal(0, 0.6, (n, 6))     # large errors: the +-20 N m clamp is active
    a, b = sim.copy(), sim.copy()
    touched = 0
    for k in range(8):
        fa, ta = cpu_port.substeps(a, tgt, 1)
        fb, tb = cpu_port.substeps(b, tgt, 1, model="halves")
        assert np.abs(a - b).max() <= 1e-9 * (k + 1) and np.abs(ta - tb).max() <= 1e-9
        assert np.abs(fa - fb).max() <= 1e-7          # contact forces (hundreds of newtons), feet applied + merged-body predictor
        touched += int((np.abs(fa[:, 0, 2]) > 1.0).sum() + (np.abs(fa[:, 6, 2]) > 1.0).sum())
        b[:] = a                                      # one-step comparison each time
    assert touched > n                                # the feet were actually on the ground
    ref = sim.copy()
    cpu_port.substeps(ref, tgt, 4)
    x32, h32 = sim.astype(np.float32), sim.astype(np.float32)
    cpu_port.substeps(x32, tgt.astype(np.float32), 4)
    cpu_port.substeps(h32, tgt.astype(np.float32), 4, model="halves")
    dx, dh = np.abs(x32 - ref).max(1), np.abs(h32 - ref).max(1)
    assert np.median(dh) <= 2.0 * np.median(dx) + 1e-7 and np.quantile(dh, 0.99) <= 3.0 * np.quantile(dx, 0.99) + 1e-6


def test_packed_halves_substep_is_the_halves_substep_lane_by_lane():
    """csrc/zbot_h2.h (the arithmetic of the default GPU kernel beyond 9472 envs): both halves of the chain in the two FP32
    lanes of one thread -- lane x eliminates bodies 0, 1, 2, lane y bodies 6, 5, 4, body 3 and the kinematics sweep stay
    scalar.  On the host a lane pair is two plain floats, so without multiply-add contraction (the port's build) the packed
    substep must reproduce the float32 halves form BIT FOR BIT: state, applied torques, foot and merged-body contact forces,
    over 8 substeps with contact, PD saturation and joint wrap.  (And the halves form equals the one-chain substep:
    test_halves_elimination_equals_the_one_chain_substep.)"""
    from oracle import cpu_port
    from zbot_lab_b200.utils import synthetic as syn
    n = 700
    rng = np.random.default_rng(9)
    st = syn.synth_sim_state(rng, n)
    keys = ("root_pos", "root_quat", "root_lin_vel", "root_ang_vel", "joint_pos", "joint_vel")
    sim = np.concatenate([st[k].reshape(n, -1) for k in keys], 1).astype(np.float32)
    sim[:40, 13:19] += np.float32(2 * np.pi - 0.05) * np.sign(rng.normal(size=(40, 6))).astype(np.float32)   # joints near the +-2 pi wrap
    tgt = (sim[:, 13:19] + rng.normal(0, 0.6, (n, 6))).astype(np.float32)
    a, b = sim.copy(), sim.copy()
    touched = 0
    for k in range(8):
        fa, ta = cpu_port.substeps(a, tgt, 1, model="halves")
        fb, tb = cpu_port.substeps(b, tgt, 1, model="h2")
        assert np.array_equal(a, b) and np.array_equal(ta, tb) and np.array_equal(fa, fb), k
        touched += int((np.abs(fa[:, 0, 2]) > 1.0).sum() + (np.abs(fa[:, 6, 2]) > 1.0).sum() + (np.abs(fa[:, 1:6]).max((1, 2)) > 0.1).sum())
    assert touched > n
    # four substeps in one call (the state stays packed across substeps) == four calls
    c, d = sim.copy(), sim.copy()
    cpu_port.substeps(c, tgt, 4, model="h2")
    for k in range(4):
        cpu_port.substeps(d, tgt, 1, model="h2")
    assert np.array_equal(c, d)


def test_unlimited_revolute_joints_wrap_at_two_pi_like_physx():
    """PhysX keeps the position of a revolute joint without limits inside [-2 pi, 2 pi] (the reference's own note,
    assets/test_articulation.py:18-20): a joint spinning past +2 pi re-enters at -2 pi (shift by 4 pi, same physical angle).
    Kernel arithmetic (one-chain and two-halves formulations) and the float64 oracle agree, the pose of every link is
    continuous across the wrap, and the snake robot's limited joint6 does not wrap."""
    from oracle import cpu_port
    from zbot_lab_b200.assets import zbot_6s as Z
    n = 4
    two_pi = 2 * np.pi
    st = {"root_pos": np.tile([0.0, 0.0, 1.0], (n, 1)), "root_quat": np.tile([1.0, 0.0, 0.0, 0.0], (n, 1)),
          "root_lin_vel": np.zeros((n, 3)), "root_ang_vel": np.zeros((n, 3)),
          "joint_pos": np.zeros((n, 6)), "joint_vel": np.zeros((n, 6))}
    st["joint_pos"][:, 1] = two_pi - 0.01          # joint 2 about to cross +2 pi
    st["joint_pos"][:, 4] = -two_pi + 0.01         # joint 5 about to cross -2 pi
    st["joint_vel"][:, 1], st["joint_vel"][:, 4] = 4.0, -4.0
    tgt = st["joint_pos"] + 2.0 * np.sign(st["joint_vel"])       # drive pushes further the same way
    o = DynOracle(n)
    o.set_state(st)
    a, b = cpu_port.pack_sim(st, np.float64), cpu_port.pack_sim(st, np.float64)
    lp0, lq0 = Z.fk_links(a[0, :3], a[0, 3:7], a[0, 13:19], o.m)
    for i in range(6):
        cpu_port.substeps(a, tgt, 1)
        cpu_port.substeps(b, tgt, 1, model="halves")
        o.substep(tgt)
        ref = np.concatenate([o.root_pos, o.root_quat, o.root_lin_vel, o.root_ang_vel, o.q, o.qd], -1)
        assert np.abs(a - ref).max() < 1e-9 and np.abs(b - ref).max() < 1e-9, i
    assert np.all(a[:, 13 + 1] < -two_pi + 1.0) and np.all(a[:, 13 + 4] > two_pi - 1.0)          # wrapped to the other end
    assert np.all(np.abs(a[:, 13:19]) <= two_pi)
    lp1, lq1 = Z.fk_links(a[0, :3], a[0, 3:7], a[0, 13:19], o.m)
    unwrapped = a[0, 13:19].copy()
    unwrapped[1] += 2 * two_pi
    unwrapped[4] -= 2 * two_pi
    lp2, lq2 = Z.fk_links(a[0, :3], a[0, 3:7], unwrapped, o.m)
    assert np.abs(lp1 - lp2).max() < 1e-12 and np.abs(np.abs(np.sum(lq1 * lq2, -1)) - 1).max() < 1e-12   # same link poses
    # snake robot: joint6 carries +-720 deg limits in zbot_6s_v03.usd -> not wrapped; joint 2 is
    from zbot_lab_b200.assets import zbot_d_6s as S
    ms = S.model_f32()
    sims = np.zeros((1, 25))
    sims[0, :3], sims[0, 3:7] = [0, 0, 1.0], ms.default_root_quat
    sims[0, 13 + 5], sims[0, 19 + 5] = two_pi - 0.005, 4.0
    sims[0, 13 + 1], sims[0, 19 + 1] = two_pi - 0.005, 4.0
    tg = sims[:, 13:19] + 1.0
    for i in range(4):
        cpu_port.substeps(sims, tg, 1, cfg=cpu_port.make_cfg(1, task=cpu_port.TASK_SNAKE_V0) if hasattr(cpu_port, "TASK_SNAKE_V0") else None, snake=True)
    assert sims[0, 13 + 5] > two_pi and sims[0, 13 + 1] < -two_pi + 1.0


def _terrain_fixture(rows=3, cols=4, curriculum=True, seed=4):
    from zbot_lab_b200.terrain import Terrain, rough_terrains_cfg
    g = rough_terrains_cfg()
    g.num_rows, g.num_cols, g.border_width, g.curriculum = rows, cols, 2.0, curriculum
    return Terrain(g, seed=seed)


def test_generated_terrain_and_importer_bookkeeping():
    """zbot_lab_b200/terrain.py (restated ROUGH_TERRAINS_CFG / TerrainImporter [IL-upstream]): tile grid geometry, every
    sub-terrain family, difficulty growing with the level, spawn points on the field, bilinear sampling, level updates; and
    the reference's own curriculum rule (mdp/curriculums.py:26-55)."""
    from zbot_lab_b200.terrain import rough_terrains_cfg, sub_terrain, terrain_levels_vel
    t = _terrain_fixture(rows=4, cols=10)
    n = t.tile_cells
    assert n == 80 and t.heights.shape == (4 * 80 + 2 * 20 + 1, 10 * 80 + 2 * 20 + 1) and t.heights.dtype == np.float32
    assert set(t.col_kind) == set(rough_terrains_cfg().sub_terrains) and t.col_kind.count("pyramid_stairs") == 2
    assert np.all(t.heights[:20] == 0) and np.all(t.heights[:, -20:] == 0)                       # flat border
    for r in range(4):
        for c in range(10):
            o = t.origins[r, c]
            assert abs(float(t.height_at(o[0], o[1])) - o[2]) < 1e-6                             # spawn point sits on the field
    # level = difficulty: the stairs' platform rises with the row; inverted ones sink
    cs, ci = t.col_kind.index("pyramid_stairs"), t.col_kind.index("pyramid_stairs_inv")
    assert np.all(np.diff(t.origins[:, cs, 2]) > 0) and np.all(np.diff(t.origins[:, ci, 2]) < 0)
    rng = np.random.default_rng(0)
    st = sub_terrain("pyramid_stairs", dict(step_height_range=(0.05, 0.23), step_width=0.3, platform_width=3.0), 1.0, 80, 0.1, rng)
    assert abs(st.max() - 0.23 * ((80 - 30) // 2 // 3)) < 1e-9 and st[0, 0] == 0 and np.all(st == st.T)
    sl = sub_terrain("pyramid_slope", dict(slope_range=(0.0, 0.4), platform_width=2.0), 0.5, 80, 0.1, rng)
    assert abs(sl[40, 40] - 0.2 * 0.1 * 30) < 1e-9 and abs((sl[5, 40] - sl[4, 40]) / 0.1 - 0.2) < 1e-9
    bx = sub_terrain("random_grid", dict(grid_width=0.45, grid_height_range=(0.05, 0.2), platform_width=2.0), 1.0, 80, 0.1, rng)
    assert np.all(bx[30:50, 30:50] == 0) and abs(bx).max() <= 0.2 and len(np.unique(bx)) > 50
    ru = sub_terrain("random_uniform", dict(noise_range=(0.02, 0.10), noise_step=0.02), 1.0, 80, 0.1, rng)
    assert ru[40, 40] == 0 and 0.02 < np.ptp(ru) <= 0.16
    # bilinear sample between four cells
    x, y = t.x0 + 57.25 * t.cell, t.y0 + 33.5 * t.cell
    H = t.heights.astype(np.float64)
    want = (H[57, 33] * 0.75 + H[58, 33] * 0.25) * 0.5 + (H[57, 34] * 0.75 + H[58, 34] * 0.25) * 0.5
    assert abs(float(t.height_at(x, y)) - want) < 1e-5
    # TerrainImporter: initial placement and level updates
    lv, ty = t.initial_levels_types(1000, 5, np.random.default_rng(1))
    assert lv.min() == 0 and lv.max() == 3 and np.all(np.diff(ty) >= 0) and ty[0] == 0 and ty[-1] == 9 and np.bincount(ty).min() == 100
    new = t.update_levels(np.array([0, 0, 3, 3, 2]), np.array([0, 1, 1, 0, 0], bool), np.array([1, 0, 0, 1, 0], bool), np.array([9, 9, 1, 9, 9]))
    assert list(new) == [0, 1, 1, 2, 2]              # clipped at 0; solved the last level -> random restart; down; unchanged
    up, down = terrain_levels_vel(np.array([[4.1, 0], [0.5, 0.3], [2.0, 0.0], [0.0, 0.0]]), np.array([[0.1, 0], [0.1, 0], [0.1, 0.0], [0, 0]]), 8.0, 20.0)
    assert list(up) == [True, False, False, False] and list(down) == [False, True, False, False]


def test_m_rough_port_matches_oracles_on_the_height_field():
    """zbot-6b-walking-m-rough-v0 (kernel arithmetic, host build): (1) the substep with the height field under every contact
    candidate equals the independent float64 dynamics oracle with the same ground function; (2) the full step with the
    terrain curriculum -- world-height termination, level moves at reset, new origins -- equals the oracle on the exported
    view: flags, levels, origins exact."""
    from helpers import M_EXTRA_TERMS, M_PARAMS_EXTRA, m_check_step, m_make_oracle, m_native_cfg
    from oracle import cpu_port
    from oracle.dyn_oracle import DynOracle, DynParams
    from zbot_lab_b200 import native
    from zbot_lab_b200.assets import zbot_6s_v2 as V
    t = _terrain_fixture()
    n = 48
    rng = np.random.default_rng(11)
    lv, ty = t.initial_levels_types(n, 2, rng)
    org4 = np.zeros((n, 4), np.float32)
    org4[:, :3] = t.origins[lv, ty]
    # (1) dynamics on slopes / boxes: robots dropped next to their spawn points (off the flat platform)
    cfg = m_native_cfg(n, native.M_FLAT_TERMS)
    m = V.model_f32()
    pt = cpu_port.make_port_terrain(t, org4, False)
    st = {"root_pos": np.tile(m.default_root_pos, (n, 1)), "root_quat": np.tile(m.default_root_quat, (n, 1)),
          "root_lin_vel": rng.normal(0, 0.1, (n, 3)), "root_ang_vel": rng.normal(0, 0.3, (n, 3)),
          "joint_pos": np.tile(m.default_joint_pos, (n, 1)) + rng.uniform(-0.1, 0.1, (n, 6)), "joint_vel": rng.normal(0, 0.3, (n, 6))}
    st["root_pos"][:, 0] += rng.uniform(-2.5, 2.5, n)
    st["root_pos"][:, 1] += rng.uniform(-2.5, 2.5, n)
    gz = t.height_at(org4[:, 0] + st["root_pos"][:, 0], org4[:, 1] + st["root_pos"][:, 1]).astype(np.float64) - org4[:, 2]
    st["root_pos"][:, 2] += gz + 0.004
    o = DynOracle(n, DynParams(model=m, mu=float(cfg.contact_mu)), model=m)
    o64 = org4.astype(np.float64)
    o.ground = lambda x, y: t.height_at(x + o64[:, 0], y + o64[:, 1], np.float64) - o64[:, 2]
    o.set_state(st)
    sim = cpu_port.pack_sim(st, np.float64)
    tgt = st["joint_pos"].copy()
    touched = 0
    for i in range(12):
        f, tau = cpu_port.substeps(sim, tgt, 1, cfg=cfg, model="m", terrain=pt)
        o.substep(tgt)
        ref = np.concatenate([o.root_pos, o.root_quat, o.root_lin_vel, o.root_ang_vel, o.q, o.qd], -1)
        assert np.abs(sim - ref).max() < 1e-6, i          # the height field itself is float32: its samples carry ~1e-7
        touched += int((np.abs(f[:, [0, 6], 2]) > 1.0).sum())
        sim[:] = ref
    assert touched > n
    # (2) the full step with the curriculum
    P = dict(M_PARAMS_EXTRA, terrain={"origins": t.origins, "tile_size": 8.0, "curriculum": True}, cmd_ranges=((-0.3, 0.3), (-0.1, 0.1), (-1.0, 1.0)))
    pe = cpu_port.PortEnv(n, np.float32, m_native_cfg(n, M_EXTRA_TERMS, P))
    org4b = org4.copy()
    pe.terrain = cpu_port.make_port_terrain(t, org4b, True)
    pd = pe.field("p_delta", 6)
    pd[:, 3], pd[:, 4] = lv, ty
    pe.field("carry_feet_fz", 2)[:] = np.stack([rng.uniform(-0.3, 0.3, n), rng.uniform(-0.1, 0.1, n)], -1)
    pe.field("base_pos_y_err_sum", 1)[:, 0] = rng.uniform(0.05, 0.3, n)
    walked = np.arange(n) % 4 == 0                          # a quarter of the robots have walked more than half a tile: move up
    pe.field("root_pos", 3)[walked, 0] += 4.3
    ep0 = rng.integers(900, 999, n)
    pe.ep_len[:] = ep0
    get = lambda k, w: pe.field(k, w).copy()
    orc = m_make_oracle(n, M_EXTRA_TERMS, get, ep0, P)
    orc.levels[:], orc.types[:], orc.env_origins[:] = lv, ty, org4[:, :3]
    ups = downs = resets = 0
    for s in range(25):
        a = rng.normal(0, 0.5, (n, 6)).astype(np.float32)
        rnd = rng.random((n, 22)).astype(np.float32)
        obs, rew, term, trunc, rs, ex = pe.step(a, export=True, rnd=rnd)
        r = m_check_step(orc, a, rnd, ex, obs, rew, term, trunc, pe.ep_len, get, rs)
        assert np.array_equal(pe.field("p_delta", 6)[:, 3].astype(np.int64), orc.levels), "terrain levels"
        assert np.array_equal(org4b[:, :3], orc.env_origins), "env origins"
        if r["log"]:
            ups, downs = ups + r["log"]["#move_up"], downs + r["log"]["#move_down"]
        resets += len(r["reset_ids"])
    assert resets > n and ups > 0 and downs > 0
