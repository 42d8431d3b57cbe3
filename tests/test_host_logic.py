"""CPU tests of the host side: the C-ABI library loads without a GPU and exports every symbol the
header declares, struct layouts agree, generated header is current, synthetic inputs are stable."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from zbot_lab_b200 import native
    lib = native.lib()
    hdr = open(os.path.join(ROOT, "include", "zbot_b200.h")).read()
    declared = set(re.findall(r"\b(zbot_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(native.EXPORTED_SYMBOLS), declared ^ set(native.EXPORTED_SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.zbot_abi_version() == native.ZBOT_ABI_VERSION
    assert b"sm_100a" in lib.zbot_build_info()


def test_cfg_struct_matches_c_defaults():
    from zbot_lab_b200 import native
    lib = native.lib()
    c = native.ZbotCfg()
    assert lib.zbot_default_cfg(C.byref(c), 4096) == 0
    assert bytes(c) == bytes(native.make_cfg(4096))     # same layout, same values, same term order
    assert c.num_terms == 13 and abs(c.term_weight[0] - 0.02) < 1e-9


def test_state_field_tables_match_library():
    from zbot_lab_b200 import native, stepper
    lib = native.lib()
    used = np.zeros(80, int)
    for k, w in stepper.STATE_FIELDS.items():
        w0 = lib.zbot_state_word(k.encode())
        assert w0 >= 0, k
        used[w0:w0 + w] += 1
    assert used.max() == 1
    used = np.zeros(72, int)
    for k, w in stepper.MDP_STATE_FIELDS.items():
        w0 = lib.zbot_mdp_state_word(k.encode())
        assert w0 >= 0, k
        used[w0:w0 + w] += 1
    assert used.max() == 1
    assert lib.zbot_state_word(b"nope") == -1


def test_generated_header_is_current():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen", os.path.join(ROOT, "tools", "gen_model_header.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    assert open(gen.HEADER).read() == gen.render()


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "zbot_lab_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(d, f)
                assert "cpu_port" not in src or f.endswith(".h") is False and "cpu_port" not in src, f


def test_env_origin_grid_matches_il_semantics():
    from oracle.il_semantics import env_origins_grid as ref
    from zbot_lab_b200.utils.synthetic import env_origins_grid
    for n in (1, 2, 7, 64, 4096):
        assert np.array_equal(env_origins_grid(n), ref(n))
    g = env_origins_grid(4096)
    assert g.shape == (4096, 3) and g[:, 0].max() == 126.0 and g[:, 1].min() == -126.0


def test_no_cuda_means_loud_failure():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from zbot_lab_b200.stepper import NativeStepper
    with pytest.raises(RuntimeError):
        NativeStepper(4, "cpu")


def test_manager_cfg_compiles_into_the_kernel_term_table_and_unknown_terms_fail_loudly():
    """zbot-6b-walking-m-v0: the cfg tree mirrors the reference (registry id / kwargs keys of
    zbotlab_manager/config/zbot6b_manager/__init__.py:14-22, term names / weights / params of zbotlab_env_cfg.py:240-393
    with the flat overrides); ManagerBasedRLEnv compiles it term by term into ZbotCfg; editing a weight, enabling a
    commented-out term or removing one is reflected; a term the kernel does not implement raises."""
    import zbot_lab_b200.tasks  # noqa: F401
    from zbot_lab_b200 import native
    from zbot_lab_b200.compat import gym_registry as gym
    from zbot_lab_b200.tasks.zbotlab_manager import mdp
    from zbot_lab_b200.tasks.zbotlab_manager.manager_env import ManagerBasedRLEnv

    spec = gym.spec("zbot-6b-walking-m-v0")
    assert set(spec.kwargs) == {"env_cfg_entry_point", "rsl_rl_cfg_entry_point"}
    agent = gym.load_cfg_from_registry("zbot-6b-walking-m-v0", "rsl_rl_cfg_entry_point")
    assert agent.num_steps_per_env == 24 and agent.policy.actor_hidden_dims == [128, 128, 128] and agent.algorithm.entropy_coef == 0.01

    def compile_(cfg, n=8):
        e = object.__new__(ManagerBasedRLEnv)
        e.cfg, e.num_envs, e.physics_dt, e.step_dt = cfg, n, 0.005, 0.02
        e.max_episode_length_s, e.max_episode_length = 20.0, 1000
        e._compile_cfg()
        return e, e._native_cfg()

    cfg = gym.load_cfg_from_registry("zbot-6b-walking-m-v0", "env_cfg_entry_point")
    e, c = compile_(cfg)
    assert c.task == native.TASK_WALKING_M and c.num_terms == 10 and c.max_episode_length == 1000
    assert list(c.term_id)[:10] == [34, 35, 8, 24, 7, 36, 1, 2, 38, 6]
    assert [round(w, 9) for w in list(c.term_weight)[:10]] == [1.0, 0.5, -1e-05, -2.5e-07, -0.01, 5.0, -1.0, -0.5, -6.5, -15.0]
    assert c.is_terminated_weight == -200.0 and abs(c.term_param[0][0] - 0.25) < 1e-7
    assert abs(c.kp - 20.0) < 1e-6 and abs(c.kd - 0.5) < 1e-6 and abs(c.termination_height - 0.2) < 1e-6 and abs(c.feet_close_min - 0.12) < 1e-6
    assert abs(c.act_scale - 0.04 * np.pi) < 1e-6 and abs(c.act_clip - 0.04 * np.pi) < 1e-6
    assert list(c.cmd_lo) == [np.float32(-0.1), 0.0, 0.0] and abs(c.cmd_rel_standing - 0.02) < 1e-7
    assert c.obs_noise_enable == 1 and c.obs_noise_hi[0] == np.float32(0.01) and c.obs_noise_hi[4] == 0.0 and c.obs_noise_hi[13] == 1.5
    # the reference's toggles: a weight edit, a commented-out term switched back on, a term removed
    cfg.rewards.feet_slide.weight = -0.2
    cfg.rewards.gait = mdp.RewardTermCfg(func=mdp.feet_gait, weight=0.5, params={
        "period": 2.0, "offset": [0.0, 0.5], "threshold": 0.55, "command_name": "base_velocity"})
    cfg.rewards.foot_forward = None
    cfg.observations.policy.enable_corruption = False
    e, c = compile_(cfg)
    ids = list(c.term_id)[:c.num_terms]
    assert 37 in ids and 2 not in ids and c.obs_noise_enable == 0
    assert abs(c.term_weight[ids.index(38)] + 0.2) < 1e-7 and list(c.term_param[ids.index(37)]) == [2.0, 0.0, 0.5, np.float32(0.55)]
    play = gym.load_cfg_from_registry("zbot-6b-walking-m-play-v0", "env_cfg_entry_point")
    e, c = compile_(play, 64)
    assert play.scene.num_envs == 64 and c.obs_noise_enable == 0 and list(c.cmd_hi)[0] == np.float32(0.3)
    # the terms ZbotLabRoughEnvCfg defines and the registered cfgs switch off (zbotlab_env_cfg.py:86-97, 253-258, 367-371, 385-388)
    full = gym.load_cfg_from_registry("zbot-6b-walking-m-v0", "env_cfg_entry_point")
    full.rewards.undesired_contacts = mdp.RewardTermCfg(func=mdp.undesired_contacts, weight=-1.0, params={
        "sensor_cfg": mdp.SceneEntityCfg("contact_forces", body_names="base|a.*|b.*"), "threshold": 1.0})
    full.terminations.base_contact = mdp.TerminationTermCfg(func=mdp.illegal_contact, params={
        "sensor_cfg": mdp.SceneEntityCfg("contact_forces", body_names="base"), "threshold": 1.0})
    full.events.push_robot = mdp.EventTermCfg(func=mdp.push_by_setting_velocity, mode="interval", interval_range_s=(10.0, 15.0),
                                              params={"velocity_range": {"x": (-0.5, 0.5), "y": (-0.5, 0.5)}})
    full.commands.base_velocity.heading_command = True
    full.commands.base_velocity.heading_control_stiffness = 0.5
    full.commands.base_velocity.ranges.heading = (-np.pi, np.pi)
    full.commands.base_velocity.ranges.ang_vel_z = (-1.0, 1.0)
    e, c = compile_(full)
    ids = list(c.term_id)[:c.num_terms]
    assert 43 in ids and c.term_param[ids.index(43)][0] == 1.0
    assert c.illegal_contact_threshold == 1.0 and c.illegal_contact_mask == 0b00100            # `base` sits in merged body 3
    assert c.cmd_heading == 1 and abs(c.cmd_heading_hi - np.pi) < 1e-6 and c.cmd_heading_stiffness == 0.5 and c.cmd_rel_heading == 1.0
    assert (c.push_interval_lo, c.push_interval_hi) == (10.0, 15.0) and list(c.push_lo) == [-0.5, -0.5] and list(c.push_hi) == [0.5, 0.5]
    assert ("base_contact", "illegal_contact") in e._done_names
    # not built -> loud: the per-env mass / CoM randomisation events (None in every registered cfg, rough_env_cfg.py:38-39)
    bad = gym.load_cfg_from_registry("zbot-6b-walking-m-v0", "env_cfg_entry_point")
    bad.events.add_base_mass = mdp.EventTermCfg(func=mdp.randomize_rigid_body_mass, mode="startup", params={
        "asset_cfg": mdp.SceneEntityCfg("robot", body_names="base"), "mass_distribution_params": (-1.0, 3.0), "operation": "add"})
    with pytest.raises(NotImplementedError):
        compile_(bad)
    bad = gym.load_cfg_from_registry("zbot-6b-walking-m-v0", "env_cfg_entry_point")
    bad.terminations.base_contact = mdp.TerminationTermCfg(func=mdp.illegal_contact, params={
        "sensor_cfg": mdp.SceneEntityCfg("contact_forces", body_names="foot.*"), "threshold": 1.0})
    with pytest.raises(NotImplementedError):
        compile_(bad)
    with pytest.raises(RuntimeError):
        mdp.feet_gait(None)


def test_graph_replay_runs_host_curricula_and_recaptures_on_change():
    """`env.step`'s Python does not run while a captured rollout graph is replayed, so
    `OnPolicyRunner.replay_rollout` must advance the step counter, evaluate the env's host curricula for the replayed
    steps, refresh the python-float `Curriculum/*` log entries and drop the graph when kernel parameters changed."""
    import torch
    from zbot_lab_b200.rl.ppo_runner import OnPolicyRunner

    class Env:
        num_envs, num_actions, max_episode_length = 4, 6, 1000
        def __init__(self):
            self.unwrapped = self
            self.common_step_counter = 0
            self.stage = 0
        def get_observations(self):
            return {"policy": torch.zeros(4, 23)}
        def advance_host_curricula(self, steps):
            before = self.stage
            for _ in range(steps):
                self.common_step_counter += 1
                if self.common_step_counter == 30:
                    self.stage = 1
            return self.stage != before
        def curriculum_log(self):
            return {"Curriculum/curriculum_stage": self.stage}

    class Graph:
        replays = 0
        def replay(self):
            Graph.replays += 1

    env = Env()
    r = OnPolicyRunner(env, {"num_steps_per_env": 24, "use_cuda_graph": False}, device="cpu")
    r._graph, r._obs_in = Graph(), torch.zeros(4, 23)
    r._graph_ep_infos = [{"Curriculum/curriculum_stage": 0} for _ in range(24)]
    r.replay_rollout()
    assert env.common_step_counter == 24 and r._graph is not None and r._graph_ep_infos[0]["Curriculum/curriculum_stage"] == 0
    r.replay_rollout()                              # step 30 falls inside this rollout: parameters changed
    assert env.common_step_counter == 48 and r._graph is None
    assert all(e["Curriculum/curriculum_stage"] == 1 for e in r._graph_ep_infos) and Graph.replays == 2


def test_runner_statistics_ring_guard():
    """log scalars are views into a 64-slot ring; rollouts longer than the ring clone them, and the
    captured graph (which needs slots T..2T-1) is only used when 2T fits."""
    import warnings
    import torch
    from zbot_lab_b200.rl.ppo_runner import OnPolicyRunner

    class St:
        stats_ring = torch.zeros(64, 32)

    class Env:
        num_envs, num_actions, max_episode_length = 4, 6, 1000
        _stepper = St()
        def __init__(self):
            self.unwrapped = self
        def get_observations(self):
            return {"policy": torch.zeros(4, 23)}

    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        r = OnPolicyRunner(Env(), {"num_steps_per_env": 40, "use_cuda_graph": True}, device="cpu")
    assert not r.use_cuda_graph and not r._clone_logs and any("statistics slots" in str(x.message) for x in w)
    r = OnPolicyRunner(Env(), {"num_steps_per_env": 24, "use_cuda_graph": True}, device="cpu")
    assert r.use_cuda_graph and not r._clone_logs
    r = OnPolicyRunner(Env(), {"num_steps_per_env": 100, "use_cuda_graph": False}, device="cpu")
    assert r._clone_logs
