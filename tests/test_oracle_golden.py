"""Pins ``oracle/mdp_oracle.py`` to the reference: the golden .npz files hold the outputs
of the reference's own unmodified code (tests/golden/make_golden.py)."""
import os
import sys

import numpy as np
import pytest

from helpers import GOLDEN_CASES, GOLDEN_DIR, load_golden, make_mdp_oracle, rel_err
from zbot_lab_b200.utils import synthetic as syn

RTOL = 1e-5  # BASELINE.json north_star: float32 reward/observation terms within 1e-5 relative


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_mdp_oracle_matches_reference_golden(name):
    g, case = load_golden(name)
    n, steps = int(g["n"]), int(g["steps"])
    o = make_mdp_oracle(n, case["origins"])
    o.episode_length_buf[:] = case["episode_length_buf0"]
    obs0 = o.observe(case["S0"])
    assert rel_err(obs0, g["obs0"]) <= RTOL
    for t, (a, S1) in enumerate(case["steps"]):
        obs, rew, term, trunc, ids, log = o.step(a, S1)
        k = t + 1
        # integer / index work: bit-exact
        assert np.array_equal(term, g[f"terminated{k}"])
        assert np.array_equal(trunc, g[f"truncated{k}"])
        assert np.array_equal(ids, g[f"reset_ids{k}"])
        assert np.array_equal(o.episode_length_buf, g[f"state{k}/episode_length_buf"])
        # float work: 1e-5 relative
        assert rel_err(obs, g[f"obs{k}"]) <= RTOL
        assert rel_err(rew, g[f"rew{k}"]) <= RTOL
        for key, v in o.mdp_state().items():
            if key == "episode_length_buf":
                continue
            assert rel_err(v, g[f"state{k}/{key}"]) <= RTOL, key
        if log is not None:
            for key, v in log.items():
                assert rel_err(np.float32(v), g[f"log{k}/{key}"]) <= RTOL, key


def test_reward_scales_match_reference_cfg():
    import os
    from helpers import GOLDEN_DIR
    from oracle.mdp_oracle import REWARD_SCALES_V2
    g = np.load(os.path.join(GOLDEN_DIR, "reward_scales_v2.npz"))
    assert list(g["names"]) == list(REWARD_SCALES_V2.keys())  # dict ORDER matters (SURVEY C-4)
    assert np.array_equal(g["values"], np.array(list(REWARD_SCALES_V2.values())))


def test_goldens_reproducible_from_reference_if_present():
    """When /root/reference exists (build container) regenerate one case and compare."""
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    import importlib.util, os
    from helpers import GOLDEN_DIR
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN_DIR, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    out = mg.run_case(1, 7, 4)
    g = np.load(os.path.join(GOLDEN_DIR, "mdp_v2_n7.npz"))
    for k, v in out.items():
        assert np.array_equal(np.asarray(v), g[k]), k


SNAKE_GOLDEN = ["snake_v0_n64", "snake_v0_n9"]


def _snake_oracle(n, case):
    from oracle.snake_mdp_oracle import SnakeMdpOracle
    from zbot_lab_b200.utils import synthetic as syn
    return SnakeMdpOracle(n, case["origins"], syn.snake_reset_tables(), case["joint_speed_limit"])


@pytest.mark.parametrize("name", SNAKE_GOLDEN)
def test_snake_mdp_oracle_matches_reference_golden(name):
    """BASELINE.json configs[3]: the snake task's MDP restatement against the reference's own code."""
    import os
    from helpers import GOLDEN_DIR
    from zbot_lab_b200.utils import synthetic as syn
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n, steps = int(g["n"]), int(g["steps"])
    case = syn.synth_snake_case(int(g["seed"]), n, steps)
    o = _snake_oracle(n, case)
    o.episode_length_buf[:] = case["episode_length_buf0"]
    assert rel_err(o.observe(case["S0"]), g["obs0"]) <= RTOL
    for t, (a, S1) in enumerate(case["steps"]):
        obs, rew, term, trunc, ids, log = o.step(a, S1)
        k = t + 1
        assert np.array_equal(term, g[f"terminated{k}"]) and np.array_equal(trunc, g[f"truncated{k}"])
        assert np.array_equal(ids, g[f"reset_ids{k}"])
        assert np.array_equal(o.episode_length_buf, g[f"state{k}/episode_length_buf"])
        assert rel_err(obs, g[f"obs{k}"]) <= RTOL and rel_err(rew, g[f"rew{k}"]) <= RTOL
        for key, v in o.mdp_state().items():
            if key != "episode_length_buf":
                assert rel_err(v, g[f"state{k}/{key}"]) <= RTOL, key
        if log is not None:
            for key, v in log.items():
                assert rel_err(np.float32(v), g[f"log{k}/{key}"]) <= RTOL, key


def test_snake_geometry_known_answers():
    """The reference hard-codes 0.318 (middle link x offset) and 0.636 (sum of the end-link CoM x) for the
    lying snake (zbot_direct_6dof_snake_v0.py:236, 333): the decoded model must reproduce them, and its
    `up_vec` / `heading_vec` conventions (:118-119) must map to world +z / +y at the default pose."""
    from zbot_lab_b200.assets import zbot_6s as Z, zbot_d_6s as S
    m = S.model_f32()
    lp, lq = Z.fk_links(m.default_root_pos, m.default_root_quat, m.default_joint_pos, m)
    assert abs(lp[6, 0] + 0.318) < 1e-6 and np.allclose(lp[:, 2], 0.05, atol=1e-6)
    com = lp + Z.quat_rotate(lq, m.link_com)
    assert abs(com[0, 0] + com[11, 0] + 0.636) < 1e-5
    assert np.allclose(Z.quat_rotate(lq[6], np.array([-1.0, 0, 0])), [0, 0, 1], atol=1e-7)
    assert np.allclose(Z.quat_rotate(lq[6], np.array([0, -1.0, 0])), [0, 1, 0], atol=1e-7)


# ------------------------------------------------------------------------------------------------ walking v4
V4_GOLDEN = ["v4_n64", "v4_n11"]


from helpers import make_v4_oracle  # noqa: E402


@pytest.mark.parametrize("name", V4_GOLDEN)
def test_v4_mdp_oracle_matches_reference_golden(name):
    """SURVEY §8 f1 (zbot-6b-walking-v4: commands, reset / interval resampling, randomised reset pose, 15 fresh
    reward terms): the numpy restatement against the outputs of the reference's own Zbot6SEnvV4 code.
    Flags, reset ids, interval-resample ids (the resample MASKS) and counters bit-exact; floats <= 1e-5."""
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n, steps = int(g["n"]), int(g["steps"])
    case = syn.synth_v4_case(int(g["seed"]), n, steps)
    o = make_v4_oracle(n, case["origins"])
    o.episode_length_buf[:] = case["episode_length_buf0"]
    o.commands[:] = case["commands0"]
    o.target_heading_yaw[:] = case["target_heading_yaw0"]
    o.interval_time_left[:] = case["interval_time_left0"]
    assert rel_err(o.observe(case["S0"]), g["obs0"]) <= 1e-5
    names = list(np.load(os.path.join(GOLDEN_DIR, "reward_scales_v4.npz"))["names"])
    assert names == list(o.reward_scales)
    n_reset = n_int = 0
    for t, (a, S1, rnd) in enumerate(case["steps"]):
        k = t + 1
        obs, rew, term, trunc, ids, iv, log = o.step(a, S1, rnd)
        assert np.array_equal(term, g[f"terminated{k}"]) and np.array_equal(trunc, g[f"truncated{k}"])
        assert np.array_equal(ids, g[f"reset_ids{k}"]) and np.array_equal(iv, g[f"interval_ids{k}"])
        assert np.array_equal(o.episode_length_buf, g[f"state{k}/episode_length_buf"])
        assert rel_err(rew, g[f"rew{k}"]) <= 1e-5
        assert rel_err(obs, g[f"obs{k}"]) <= 1e-5
        st = o.mdp_state()
        for nm in ("p_delta", "actions", "commands", "target_heading_yaw", "current_yaw", "feet_contact_forces_last",
                   "feet_down_pos_last", "feet_step_length", "interval_time_left"):
            assert rel_err(st[nm], g[f"state{k}/{nm}"]) <= 1e-5, (k, nm)
        for nm in names:
            assert rel_err(st["episode_sum/" + nm], g[f"state{k}/episode_sum/{nm}"]) <= 1e-5, nm
        if len(ids) > 0:
            for nm in names:
                want = float(g[f"log{k}/Episode_Reward/{nm}"])
                assert abs(float(log["Episode_Reward/" + nm]) - want) <= 1e-5 * max(1.0, abs(want)), nm
            assert log["Episode_Termination/died"] == g[f"log{k}/Episode_Termination/died"]
            assert log["Episode_Termination/time_out"] == g[f"log{k}/Episode_Termination/time_out"]
        n_reset += len(ids)
        n_int += len(iv)
    assert n_reset > 0 and n_int > 0


# ---------------------------------------------------------------------------------------------
# zbot-6b-walking-m-v0: every RewTerm / DoneTerm function of the reference's zbotlab_manager/mdp
# ---------------------------------------------------------------------------------------------
M_GOLDEN = ["m_v0_n64", "m_v0_n7"]


def _m_funcs():
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py")).read()
    ns = {}
    exec(src[src.index("M_FUNCS = ["):src.index("def run_m_case")], ns)
    return ns["M_FUNCS"]


@pytest.mark.parametrize("name", M_GOLDEN)
def test_m_oracle_terms_match_the_references_own_functions(name):
    """oracle/m_mdp_oracle.MTerms against tests/golden/m_v0_*.npz = outputs of the reference's unmodified
    zbotlab_manager/mdp/rewards.py functions (13 RewTerm functions) and terminations.py:feet_close on seeded synthetic data:
    values <= 1e-5 relative, the stateful terms' env attributes (feet_step_length, feet_down_pos_last,
    feet_contact_forces_last, feet_force_sum) and the feet_close mask exact."""
    from oracle.m_mdp_oracle import MTerms, synth_m_views
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    n, steps, seed = int(g["n"]), int(g["steps"]), int(g["seed"])
    s = {"feet_contact_forces_last": g["state0/feet_contact_forces_last"].copy(), "feet_down_pos_last": g["state0/feet_down_pos_last"].copy(),
         "feet_force_sum": g["state0/feet_force_sum"].copy(), "feet_step_length": np.zeros((n, 2), np.float32), "step_dt": 0.02}
    for t, (view, _a, _r) in enumerate(synth_m_views(seed, n, steps)):
        s["episode_length_buf"] = g[f"ep{t}"]
        cmd = g[f"cmd{t}"]
        for func, p in _m_funcs():
            val, ref = getattr(MTerms, func)(view, s, cmd, p), g[f"val{t}/{func}"]
            assert (np.abs(val - ref) / np.maximum(np.abs(ref), 1.0)).max() <= 1e-5, func
        close = np.linalg.norm(view["feet_pos"][:, 0] - view["feet_pos"][:, 1], axis=-1) < np.float32(0.12)
        assert np.array_equal(close, g[f"feet_close{t}"])
        for k in ("feet_contact_forces_last", "feet_down_pos_last", "feet_force_sum"):
            assert np.array_equal(s[k], g[f"state{t + 1}/{k}"]), k
        assert np.abs(s["feet_step_length"] - g[f"state{t + 1}/feet_step_length"]).max() <= 1e-6
    assert any(g[f"feet_close{t}"].any() for t in range(steps)) and any((g[f"val{t}/feet_gait"] > 0).any() for t in range(steps))


def test_m_goldens_regenerate_from_the_reference():
    """Build container only: re-run the reference's functions and compare with the committed fixtures."""
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("/root/reference is not present (GPU box)")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    for name, seed, n, steps in make_golden.M_CASES:
        new = make_golden.run_m_case(seed, n, steps)
        old = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
        for k in old.files:
            assert np.array_equal(np.asarray(new[k]), old[k]), (name, k)
