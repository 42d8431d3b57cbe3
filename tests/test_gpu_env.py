"""GPU tests of the drop-in surface: registry -> gym.make -> ZbotDirectEnvV2 -> RslRlVecEnvWrapper."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(n, **over):
    import zbot_lab_b200.tasks  # noqa: F401
    from zbot_lab_b200.compat import gym_registry as gym
    cfg = gym.load_cfg_from_registry("zbot-6b-walking-v2", "env_cfg_entry_point")
    cfg.scene.num_envs = n
    cfg.sim.device = "cuda:0"
    cfg.seed = 3
    for k, v in over.items():
        setattr(cfg, k, v)
    return gym.make("zbot-6b-walking-v2", cfg=cfg, render_mode=None), cfg


def test_env_protocol_and_wrapper():
    from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper
    env, cfg = _make(64)
    assert env.step_dt == pytest.approx(0.02) and env.max_episode_length == 1000
    assert env.max_episode_length_s == 20.0 and env.num_envs == 64 and env.unwrapped is env
    w = RslRlVecEnvWrapper(env, clip_actions=None)
    assert w.num_actions == 6 and w.num_obs == 23
    obs = w.get_observations()["policy"]
    assert obs.shape == (64, 23) and obs.dtype == torch.float32
    # default pose known answers (…env_v2.py:403-404): base quat in obs[0:4], q - q_default = 0
    assert torch.allclose(obs[0, :4], torch.tensor([0.6003, -0.6003, -0.3735, -0.3739], device="cuda:0"), atol=1e-4)
    assert torch.all(obs[:, 4:22] == 0) and torch.all(obs[:, 22] == 1.0)
    assert env.episode_length_buf.dtype == torch.int64
    assert int(env.episode_length_buf.max()) < 1000 and int(env.episode_length_buf.max()) > 0   # randint spread
    w.episode_length_buf = torch.randint_like(w.episode_length_buf, high=1000)                   # train.py:205 path
    prev_obs = None
    for t in range(40):
        a = torch.randn(64, 6, device="cuda:0")
        obs, rew, dones, extras = w.step(a)
        obs = obs["policy"]
        assert obs.shape == (64, 23) and rew.shape == (64,) and dones.dtype == torch.long
        assert extras["time_outs"].dtype == torch.bool and "log" in extras
        assert set(extras["log"]) == {"Episode_Reward/" + k for k in cfg.reward_cfg["reward_scales"]} | {
            "Episode_Termination/body_contact", "Episode_Termination/time_out"}
        if prev_obs is not None:
            assert prev_obs.data_ptr() != obs.data_ptr()       # outputs rotate, rsl_rl keeps obs_t alive
        prev_obs = obs
    assert env.common_step_counter == 40
    # articulation view in world frame (origin added)
    p = env._robot.data.body_link_pos_w
    assert p.shape == (64, 12, 3)
    assert torch.allclose((p[:, 6, 2]), p[:, 6, 2].clamp(0.0, 0.4))
    w.close()


def test_env_matches_bare_stepper_and_cfg_weights_drive_the_kernel():
    from zbot_lab_b200.stepper import NativeStepper
    env, cfg = _make(32, check_all_envs_reset=False)
    st = NativeStepper(32, "cuda:0")
    st.reset_idx(None)
    env.reset()
    env.episode_length_buf = torch.zeros(32, dtype=torch.int64)
    g = torch.Generator(device="cuda:0").manual_seed(0)
    for t in range(10):
        a = torch.randn(32, 6, device="cuda:0", generator=g)
        o1, r1, te1, tr1, _ = env.step(a)
        o2, r2, te2, tr2 = st.step(a)
        assert torch.equal(o1["policy"], o2) and torch.equal(r1, r2)
        assert torch.equal(te1, te2.bool()) and torch.equal(tr1, tr2.bool())
    # a cfg with a different term subset / order / weights (names drive the kernel's table)
    env2, _ = _make(32, reward_cfg={"reward_scales": {"torques": -0.5, "base_vel_forward": 2.0}},
                    check_all_envs_reset=False)
    env2.reset()
    _, r, _, _, ex = env2.step(torch.zeros(32, 6, device="cuda:0"))
    assert set(ex["log"]) == {"Episode_Reward/torques", "Episode_Reward/base_vel_forward",
                              "Episode_Termination/body_contact", "Episode_Termination/time_out"}
    assert torch.isfinite(r).all()
    with pytest.raises(AttributeError):        # the reference's getattr(self, "_reward_" + name) (…env_v2.py:252)
        _make(4, reward_cfg={"reward_scales": {"no_such_term": 1.0}})
    env.close(); env2.close(); st.close()


def test_all_envs_reset_spreads_episode_lengths_on_torch_generator():
    """…env_v2.py:418-422: when every env resets in one step the counters are re-drawn with
    torch.randint_like on the (device) torch generator -- bit-exact against the same call."""
    env, _ = _make(16, check_all_envs_reset=True)
    env.reset()
    env.episode_length_buf = torch.full((16,), 998, dtype=torch.int64)
    torch.manual_seed(123)
    _, _, term, trunc, _ = env.step(torch.zeros(16, 6, device="cuda:0"))
    assert trunc.all()
    torch.manual_seed(123)
    want = torch.randint_like(env.episode_length_buf, high=1000)
    assert torch.equal(env.episode_length_buf, want)
    env.close()


def test_all_envs_reset_spread_on_the_device_for_large_env_counts():
    """Same event for N > 256 (default cfg): no host sync -- the statistics kernel sees `#reset == N` and writes the
    counters itself (in-kernel generator): spread over [0, max_episode_length), different per env, deterministic per
    seed, and a PARTIAL reset leaves the counters of the other envs at +1 / the reset ones at 0."""
    n = 1024
    outs = []
    for rep in range(2):
        env, _ = _make(n)
        assert not env._check_all_reset
        env.reset()
        env.episode_length_buf = torch.full((n,), 998, dtype=torch.int64)
        _, _, term, trunc, _ = env.step(torch.zeros(n, 6, device="cuda:0"))
        assert trunc.all()
        ep = env.episode_length_buf.clone()
        assert int(ep.min()) >= 0 and int(ep.max()) < 1000
        assert ep.unique().numel() > n // 3 and abs(float(ep.float().mean()) - 500.0) < 40.0
        outs.append(ep)
        # partial reset: only env 0 times out -> plain zero for it, +1 for the others
        before = env.episode_length_buf.clone()
        before[0] = 998
        env.episode_length_buf = before
        _, _, term, trunc, _ = env.step(torch.zeros(n, 6, device="cuda:0"))
        after = env.episode_length_buf
        done = term | trunc
        assert bool(trunc[0]) and not bool(done.all())
        assert torch.equal(after[~done], before[~done] + 1) and torch.all(after[done] == 0)
        env.close()
    assert torch.equal(outs[0], outs[1])


def test_in_kernel_generator_advances_under_cuda_graph_replay():
    """the stream position of the in-kernel generator lives in device memory (read by the step kernel,
    bumped by the statistics kernel), so replays of ONE captured graph draw fresh uniforms: the v4 task's randomised
    reset poses / resampled commands and the observation noise differ from replay to replay."""
    import zbot_lab_b200.tasks  # noqa: F401
    from zbot_lab_b200 import native
    from zbot_lab_b200.stepper import NativeStepper
    n = 512
    st = NativeStepper(n, "cuda:0", native.set_obs_noise(native.make_cfg(n, task=native.TASK_WALKING_V4, rng_seed=11),
                                                          {"joint_pos": (-0.01, 0.01)}))
    st.reset_idx_v4(None)
    a = torch.zeros(n, 6, device="cuda:0")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            st._slot = -1
            st.step(a)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    st._slot = -1
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        st.step(a)
    seen_obs, seen_cmd, seen_pos = [], [], []
    for rep in range(3):
        st.episode_length_buf[:] = 998          # every env times out in the replayed step -> reset pose + commands re-drawn
        g.replay()
        torch.cuda.synchronize()
        seen_obs.append(st.obs.clone())
        seen_cmd.append(st.state.get("carry_feet_fz").clone())
        seen_pos.append(st.state.get("root_pos").clone())
    for i in range(2):
        assert not torch.equal(seen_cmd[i], seen_cmd[i + 1])
        assert not torch.equal(seen_pos[i], seen_pos[i + 1])
        assert not torch.equal(seen_obs[i][:, 4:10], seen_obs[i + 1][:, 4:10])      # noise on the joint_pos columns
    st.close()


def test_ppo_runner_two_iterations_and_checkpoint(tmp_path):
    """train.py flow: gym.make -> RslRlVecEnvWrapper -> OnPolicyRunner.learn (config 5 shape, tiny)."""
    from zbot_lab_b200.compat import gym_registry as gym
    from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper
    from zbot_lab_b200.rl.ppo_runner import OnPolicyRunner
    env, _ = _make(256)
    agent_cfg = gym.load_cfg_from_registry("zbot-6b-walking-v2", "rsl_rl_cfg_entry_point")
    w = RslRlVecEnvWrapper(env, clip_actions=agent_cfg.clip_actions)
    r = OnPolicyRunner(w, agent_cfg.to_dict(), log_dir=str(tmp_path), device="cuda:0")
    hist = r.learn(num_learning_iterations=2, init_at_random_ep_len=True)
    assert len(hist) == 2 and all(np.isfinite(h["value_loss"]) for h in hist)
    assert "Episode_Reward/base_vel_forward" in hist[-1] and "Episode_Termination/time_out" in hist[-1]
    ck = tmp_path / "model_2.pt"
    assert ck.exists()
    r2 = OnPolicyRunner(w, agent_cfg.to_dict(), log_dir=None, device="cuda:0")
    r2.load(str(ck))
    pol = r2.get_inference_policy(device="cuda:0")
    obs = w.get_observations()
    assert pol(obs).shape == (256, 6) and pol(obs["policy"]).shape == (256, 6)
    w.close()


def test_cuda_graph_rollout_matches_eager_semantics(tmp_path):
    """The captured 24-step rollout graph advances the env exactly like 24 eager steps would: counters,
    finite storage, statistics slots, and a PPO update runs on the replayed buffers."""
    from zbot_lab_b200.compat import gym_registry as gym
    from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper
    from zbot_lab_b200.rl.ppo_runner import OnPolicyRunner
    env, _ = _make(512, check_all_envs_reset=False)
    acfg = gym.load_cfg_from_registry("zbot-6b-walking-v2", "rsl_rl_cfg_entry_point").to_dict()
    acfg["use_cuda_graph"] = True
    w = RslRlVecEnvWrapper(env)
    r = OnPolicyRunner(w, acfg, log_dir=None, device="cuda:0")
    w.episode_length_buf = torch.zeros(512, dtype=torch.int64)
    obs = w.get_observations()["policy"]
    obs = r.capture_rollout(obs)                       # 2 warm-up rollouts + 1 captured = 72 steps
    ep0 = env.episode_length_buf.clone()
    done_any = (r.buf["done"].sum(0) > 0)
    obs, infos = r.replay_rollout()
    torch.cuda.synchronize()
    assert len(infos) == 24 and torch.isfinite(r.buf["obs"]).all() and torch.isfinite(r.buf["rew"]).all()
    still = r.buf["done"].sum(0) == 0
    assert torch.equal(env.episode_length_buf[still], ep0[still] + 24)      # 24 control steps per replay
    assert obs.shape == (512, 23) and torch.isfinite(obs).all()              # static output buffer is readable
    hist = r.learn(num_learning_iterations=2)
    assert len(hist) == 2 and all(np.isfinite(h["surrogate_loss"]) for h in hist)
    w.close()


def test_snake_env_registry_protocol_and_log_keys():
    """zbot-6s-snake-v0 through the registry: reference id / entry-point keys (zbot6_direct/__init__.py:25-33),
    episode length 800 (16 s / 0.02 s), per-env joint_speed_limit in ((0.2..2.0) * pi) as the last observation
    column, log keys of snake_v0.py:277-293."""
    import zbot_lab_b200.tasks  # noqa: F401
    from zbot_lab_b200.compat import gym_registry as gym
    from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper
    cfg = gym.load_cfg_from_registry("zbot-6s-snake-v0", "env_cfg_entry_point")
    agent = gym.load_cfg_from_registry("zbot-6s-snake-v0", "rsl_rl_cfg_entry_point")
    assert agent.num_steps_per_env == 16 and agent.policy.actor_hidden_dims == [256, 256, 128]
    cfg.scene.num_envs = 128
    cfg.sim.device = "cuda:0"
    cfg.seed = 5
    env = gym.make("zbot-6s-snake-v0", cfg=cfg, render_mode=None)
    assert env.max_episode_length == 800 and env.step_dt == pytest.approx(0.02)
    w = RslRlVecEnvWrapper(env, clip_actions=None)
    obs = w.get_observations()["policy"]
    sp = obs[:, 22]
    assert float(sp.min()) >= 0.2 * np.pi - 1e-6 and float(sp.max()) <= 2.0 * np.pi + 1e-6 and float(sp.std()) > 0.5
    # default pose: base link a4 = root orientation (x) Rz(180): (0.707,0,-0.707,0)(x)(0,0,0,1) = (0,-0.707,0,0.707)
    assert torch.allclose(obs[0, :4].abs(), torch.tensor([0.0, 0.70710678, 0.0, 0.70710678], device="cuda:0"), atol=1e-5)
    assert torch.all(obs[:, 4:22] == 0)
    for t in range(30):
        obs, rew, dones, extras = w.step(torch.randn(128, 6, device="cuda:0"))
        assert torch.isfinite(obs["policy"]).all() and torch.isfinite(rew).all()
        assert torch.equal(obs["policy"][:, 22], sp)
        assert set(extras["log"]) == {"Episode_Reward/" + k for k in cfg.reward_cfg["reward_scales"]} | {
            "Episode_Termination/died", "Episode_Termination/time_out"}
    assert env.base_heading_y_sum.shape == (128,)
    w.close()


@pytest.mark.parametrize("n", [1000, 19001])   # 19001 > 148 * 128: the unroll-2 instantiation (one more variant family)
def test_step_host_zero_copy_equals_device_step(n):
    """env.step_host (pinned host actions in, packed host result out, zero-copy inside the kernel launch) returns
    bit-for-bit what env.step returns on the device for the same state and actions."""
    a_env, _ = _make(n, check_all_envs_reset=False)
    b_env, _ = _make(n, check_all_envs_reset=False)
    b_env.episode_length_buf = a_env.episode_length_buf.clone()
    g = torch.Generator().manual_seed(4)
    # the rows live inside a larger pinned buffer with canaries on both sides (ragged tail CTA, 25-word rows)
    big = torch.full((n * 25 + 128,), -7.0).pin_memory()
    h_out = big[64:64 + n * 25].view(n, 25)
    for t in range(25):
        h_act = torch.randn(n, 6, generator=g).pin_memory()
        obs_h, rew_h, term_h, trunc_h = b_env.step_host(h_act, h_out)
        obs, rew, term, trunc, _ = a_env.step(h_act.to("cuda:0"))
        assert torch.equal(obs["policy"].cpu(), obs_h) and torch.equal(rew.cpu(), rew_h)
        assert torch.equal(term.cpu(), term_h) and torch.equal(trunc.cpu(), trunc_h)
    assert torch.equal(a_env._stepper.state.buf, b_env._stepper.state.buf)
    assert torch.all(big[:64] == -7.0) and torch.all(big[64 + n * 25:] == -7.0)
    assert torch.equal(a_env._stepper.stats.cpu()[:22], b_env._stepper.stats.cpu()[:22])
    a_env.close()
    b_env.close()


def test_v4_env_registry_protocol_curriculum_and_log_keys():
    """zbot-6b-walking-v4 through the registry (reference id / kwargs keys, zbot6b_direct/__init__.py:91-99):
    24-wide observation, commands inside the cfg ranges, log keys of …env_v4.py:893-933, and the host curricula
    (my_curriculum thresholds, …env_v4.py:138-198) re-weighting the live kernel."""
    import zbot_lab_b200.tasks  # noqa: F401
    from zbot_lab_b200.compat import gym_registry as gym
    from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper
    cfg = gym.load_cfg_from_registry("zbot-6b-walking-v4", "env_cfg_entry_point")
    agent = gym.load_cfg_from_registry("zbot-6b-walking-v4", "rsl_rl_cfg_entry_point")
    assert agent.experiment_name == "zbot_6b_flat_direct_v4" and agent.policy.actor_hidden_dims == [256, 256, 128]
    cfg.scene.num_envs = 256
    cfg.sim.device = "cuda:0"
    cfg.seed = 9
    env = gym.make("zbot-6b-walking-v4", cfg=cfg, render_mode=None)
    w = RslRlVecEnvWrapper(env, clip_actions=None)
    assert w.num_obs == 24 and w.num_actions == 6 and env.max_episode_length == 1000
    obs = w.get_observations()["policy"]
    assert obs.shape == (256, 24) and torch.isfinite(obs).all()
    assert torch.allclose(obs[:, 22], torch.full((256,), 0.3, device="cuda:0"))          # velocity_range (0.3, 0.3), prob_pos 1
    assert float(obs[:, 23].abs().max()) <= 0.1 + 1e-5                                     # heading_err = commanded relative yaw
    assert float(obs[:, 0].std()) > 0.05                                                   # random yaw at reset
    for t in range(30):
        obs, rew, dones, extras = w.step(torch.randn(256, 6, device="cuda:0"))
        assert obs["policy"].shape == (256, 24) and torch.isfinite(obs["policy"]).all() and torch.isfinite(rew).all()
    assert set(extras["log"]) == {"Episode_Reward/" + k for k in cfg.reward_cfg["reward_scales"]} | {
        "Episode_Termination/died", "Episode_Termination/time_out", "Curriculum/curriculum_stage",
        "Curriculum/vel_lower_bound", "Curriculum/vel_upper_bound", "Curriculum/yaw_bound"}
    # my_curriculum stage 0 -> 1 at 12 episodes' worth of steps: new weights reach the kernel
    env.common_step_counter = 12 * env.max_episode_length - 1
    w.step(torch.zeros(256, 6, device="cuda:0"))
    assert env.curriculum_stage == 1 and env.reward_scales["airtime_variance"] == -10.0
    ids = [env._stepper.cfg.term_id[i] for i in range(env._stepper.cfg.num_terms)]
    from zbot_lab_b200 import native
    assert abs(env._stepper.cfg.term_weight[ids.index(native.V4_TERM_IDS["airtime_variance"])] + 10.0) < 1e-6
    env.common_step_counter = 24 * env.max_episode_length - 1
    w.step(torch.zeros(256, 6, device="cuda:0"))
    assert env.curriculum_stage == 2 and abs(env._stepper.cfg.ev_prob_pos - 0.8) < 1e-6
    w.close()


def test_observation_noise_is_additive_uniform_and_touches_nothing_else():
    """Opt-in ObservationManager-style corruption (manager-based task's PolicyCfg: base_quat +-0.01, joint_pos +-0.01,
    joint_vel +-1.5, zbotlab_manager/zbotlab_env_cfg.py): emitted obs = clean obs + U(n_min, n_max) per element; rewards,
    flags and the simulation state are bit-identical to the run without noise; the draw is reproducible per seed."""
    n = 4096
    noise = {"base_quat": (-0.01, 0.01), "joint_pos": (-0.01, 0.01), "joint_vel": (-1.5, 1.5)}
    clean, _ = _make(n, check_all_envs_reset=False)
    noisy, _ = _make(n, check_all_envs_reset=False, observation_noise=noise)
    noisy2, _ = _make(n, check_all_envs_reset=False, observation_noise=noise)
    for e in (noisy, noisy2):
        e.episode_length_buf = clean.episode_length_buf.clone()
    g = torch.Generator(device="cuda:0").manual_seed(2)
    for t in range(6):
        a = torch.randn(n, 6, device="cuda:0", generator=g)
        o0, r0, te0, tr0, _ = clean.step(a)
        o1, r1, te1, tr1, _ = noisy.step(a)
        o2, _, _, _, _ = noisy2.step(a)
        assert torch.equal(r0, r1) and torch.equal(te0, te1) and torch.equal(tr0, tr1)
        assert torch.equal(o1["policy"], o2["policy"])                       # same seed, same stream position
        d = o1["policy"] - o0["policy"]
        assert torch.all(d[:, 16:] == 0)                                     # actions, joint_speed_limit: no noise term
        for sl, w in ((slice(0, 10), 0.01), (slice(10, 16), 1.5)):
            x = d[:, sl]
            assert float(x.abs().max()) <= w * (1 + 1e-5) and abs(float(x.mean())) < 0.05 * w
            assert abs(float(x.std()) - w / 3 ** 0.5) < 0.05 * w             # uniform on [-w, w]: std = w / sqrt(3)
        if t > 0:
            assert not torch.equal(d, d_prev)                                # fresh draw every step
        d_prev = d.clone()
    assert torch.equal(clean._stepper.state.buf, noisy._stepper.state.buf)
    for e in (clean, noisy, noisy2):
        e.close()


def test_manager_env_registry_protocol_noise_friction_curriculum_and_ppo(tmp_path):
    """zbot-6b-walking-m-v0 through the registry (reference id / kwargs keys, zbotlab_manager/config/zbot6b_manager/
    __init__.py:14-22): ManagerBasedRLEnv over the fused step -- 25-wide policy group, standing robot at the cfg's init
    pose, per-env friction from the startup material event (64 buckets in [0.3, 1.0]), commands inside the ranges,
    the managers' log keys, the lin_vel_cmd_levels curriculum widening the live kernel's ranges, and a PPO iteration
    through the rsl_rl wrapper."""
    import zbot_lab_b200.tasks  # noqa: F401
    from zbot_lab_b200.compat import gym_registry as gym
    from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper
    from zbot_lab_b200.rl.ppo_runner import OnPolicyRunner
    cfg = gym.load_cfg_from_registry("zbot-6b-walking-m-v0", "env_cfg_entry_point")
    cfg.scene.num_envs = 256
    cfg.sim.device = "cuda:0"
    cfg.seed = 9
    env = gym.make("zbot-6b-walking-m-v0", cfg=cfg, render_mode=None)
    w = RslRlVecEnvWrapper(env, clip_actions=None)
    assert w.num_obs == 25 and w.num_actions == 6 and env.max_episode_length == 1000 and abs(env.step_dt - 0.02) < 1e-12
    mu = env.friction
    assert float(mu.min()) >= 0.3 and float(mu.max()) <= 1.0 and len(torch.unique(mu)) <= 64 and len(torch.unique(mu)) > 8
    obs = w.get_observations()["policy"]
    assert obs.shape == (256, 25) and torch.isfinite(obs).all()
    assert float(obs[:, 1:3].abs().max()) < 2e-3 and float(obs[:, 0].std()) > 0.05             # yaw-only random root pose
    assert float(obs[:, 4].abs().max()) <= 0.1 + 1e-6 and float(obs[:, 5:7].abs().max()) == 0   # ranges lin_vel_x (-0.1, 0.1)
    assert float(obs[:, 7:19].abs().max()) < 1e-5                                               # default joints, at rest
    stood = []
    for t in range(30):
        obs, rew, dones, extras = w.step(torch.zeros(256, 6, device="cuda:0"))
        assert obs["policy"].shape == (256, 25) and torch.isfinite(obs["policy"]).all() and torch.isfinite(rew).all()
        stood.append(int(dones.sum()))
    # zero actions = hold the current joint positions: the biped stands; what ends episodes here is the cfg's own margin --
    # the init stance is 0.12002 m wide against feet_close's 0.12 m, so a few hundredths of a millimetre of settling trip it
    assert sum(stood[:5]) == 0 and sum(stood) < 256 * 3
    want = {"Episode_Reward/" + k for k in ("track_lin_vel_xy_exp", "track_ang_vel_z_exp", "termination_penalty", "dof_torques_l2",
                                            "dof_acc_l2", "action_rate_l2", "foot_step_length", "foot_downward", "foot_forward",
                                            "feet_slide", "air_time_variance")}
    want |= {"Episode_Termination/time_out", "Episode_Termination/base_height", "Episode_Termination/feet_close",
             "Curriculum/lin_vel_cmd_levels"}
    assert set(extras["log"]) == want
    # random actions make envs fall; the log counts are consistent with the flags
    g = torch.Generator(device="cuda:0").manual_seed(3)
    n_term = 0
    for t in range(40):
        obs, rew, dones, extras = w.step(torch.randn(256, 6, device="cuda:0", generator=g) * 3.0)
        k = int(env.reset_terminated.sum())
        if k:
            lg = extras["log"]
            tot = float(lg["Episode_Termination/base_height"]) + float(lg["Episode_Termination/feet_close"])
            assert tot >= k - 1e-3 and tot <= 2 * k + 1e-3           # an env may trip both terms in one step
            assert abs(float(lg["Episode_Reward/termination_penalty"]) - (k / float(dones.sum())) * (-200.0 * 0.02 / 20.0)) < 1e-6
        n_term += k
    assert n_term > 0
    # lin_vel_cmd_levels (mdp/curriculums.py:57-83): at a multiple of max_episode_length, tracking reward above 80 % of
    # its weight widens lin_vel_x by 0.1 on both sides (clamped to limit_ranges)
    i = env._term_names.index("track_lin_vel_xy_exp")
    env.common_step_counter = 3 * env.max_episode_length
    env._stepper.stats_ring[max(env._stepper._slot, 0), 16] = 3.0         # some envs reset in that step
    env._stepper.stats_ring[max(env._stepper._slot, 0), i] = 0.5          # below 80 % of the weight: nothing moves
    env._lin_vel_cmd_levels()
    assert tuple(cfg.commands.base_velocity.ranges.lin_vel_x) == (-0.1, 0.1)
    env._stepper.stats_ring[max(env._stepper._slot, 0), i] = 0.95
    env.common_step_counter += 1                                           # not a multiple of max_episode_length
    env._lin_vel_cmd_levels()
    assert tuple(cfg.commands.base_velocity.ranges.lin_vel_x) == (-0.1, 0.1)
    env.common_step_counter = 4 * env.max_episode_length
    env._stepper.stats_ring[max(env._stepper._slot, 0), 16] = 0.0         # nobody reset in that step: the term is not evaluated
    env._lin_vel_cmd_levels()
    assert tuple(cfg.commands.base_velocity.ranges.lin_vel_x) == (-0.1, 0.1)
    env._stepper.stats_ring[max(env._stepper._slot, 0), 16] = 3.0
    env._lin_vel_cmd_levels()
    assert tuple(round(v, 6) for v in cfg.commands.base_velocity.ranges.lin_vel_x) == (-0.2, 0.2)
    assert abs(env._stepper.cfg.cmd_hi[0] - 0.2) < 1e-6 and abs(env._stepper.cfg.cmd_lo[0] + 0.2) < 1e-6
    for _ in range(3):                                                     # clamped to limit_ranges (-0.3, 0.3)
        env._lin_vel_cmd_levels()
    assert tuple(round(v, 6) for v in cfg.commands.base_velocity.ranges.lin_vel_x) == (-0.3, 0.3)
    obs, rew, dones, extras = w.step(torch.zeros(256, 6, device="cuda:0"))
    assert extras["log"]["Curriculum/lin_vel_cmd_levels"] == 0.3
    # PPO through the wrapper (agent cfg of the registry)
    acfg = gym.load_cfg_from_registry("zbot-6b-walking-m-v0", "rsl_rl_cfg_entry_point").to_dict()
    r = OnPolicyRunner(w, acfg, log_dir=str(tmp_path), device="cuda:0")
    hist = r.learn(num_learning_iterations=2, init_at_random_ep_len=True)
    assert len(hist) == 2 and all(np.isfinite(h["value_loss"]) for h in hist)
    assert "Episode_Reward/foot_step_length" in hist[-1]
    w.close()


def _make_cls(cls, n, scales=None, **over):
    import zbot_lab_b200.tasks  # noqa: F401
    from zbot_lab_b200.compat import gym_registry as gym
    cfg = gym.load_cfg_from_registry("zbot-6b-walking-v2", "env_cfg_entry_point")
    cfg.scene.num_envs, cfg.sim.device, cfg.seed = n, "cuda:0", 3
    cfg.check_all_envs_reset = False
    if scales is not None:
        cfg.reward_cfg = {"reward_scales": dict(scales)}
    for k, v in over.items():
        setattr(cfg, k, v)
    return cls(cfg=cfg, render_mode=None)


def test_subclass_reward_hooks_host_terms_and_step_view():
    """SURVEY §8b "subclass hooks": reward terms are discovered by `getattr(self, "_reward_" + name)` over the cfg dict
    (…env_v2.py:246-252).  (1) a subclass ADDS a term -> evaluated on the host after the kernel on the step view (end of
    physics, before the reset), weight * step_dt, with its own episode sum / log entry; (2) a subclass OVERRIDES every
    built-in term with the torch view of it -> the host-evaluated total equals the fused kernel's reward, which checks the
    15 `_reward_<name>` views against the kernel term by term; (3) an unknown key without a method raises like the reference;
    (4) `_contact_sensor.data`, `_get_dones / _get_rewards / _get_observations / _reset_idx` views."""
    import warnings
    from zbot_lab_b200.tasks.zbot6b_direct.host_terms import RewardTermViews
    from zbot_lab_b200.tasks.zbot6b_direct.walking_v2 import ZbotDirectEnvV2
    from zbot_lab_b200.tasks.zbot6b_direct.walking_v2_cfg import REWARD_SCALES_V2
    n = 512

    class WithHeight(ZbotDirectEnvV2):
        def _reward_base_height(self):
            return self._robot.data.body_link_pos_w[:, self.base_body_idx[0], 2]

    class AllOnHost(ZbotDirectEnvV2):
        pass
    # the three stateful terms stay in the kernel: their integrators / touchdown memory are advanced BY the kernel's term
    # evaluation, the built-in methods only view them
    stateful = {"base_heading_x_sum", "step_length", "base_pos_y_err_sum"}
    moved = [k for k in REWARD_SCALES_V2 if k not in stateful]
    for name in moved:
        setattr(AllOnHost, "_reward_" + name, (lambda nm: lambda self: getattr(RewardTermViews, "_reward_" + nm)(self))(name))

    plain = _make_cls(ZbotDirectEnvV2, n, capture_step_view=True)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        plus = _make_cls(WithHeight, n, {**REWARD_SCALES_V2, "base_height": 0.5})
        host = _make_cls(AllOnHost, n)
    assert sum("evaluated on the HOST" in str(x.message) for x in w) == 1 + len(moved)
    assert plus._stepper.cfg.num_terms == 13 and host._stepper.cfg.num_terms == 3
    with pytest.raises(AttributeError):
        _make_cls(ZbotDirectEnvV2, 64, {**REWARD_SCALES_V2, "no_such_term": 1.0})
    for e in (plain, plus, host):
        e.reset()
        e.episode_length_buf = torch.randint(0, 990, (n,), device="cuda:0", generator=torch.Generator("cuda:0").manual_seed(1))
    g = torch.Generator("cuda:0").manual_seed(5)
    resets = 0
    for t in range(40):
        a = torch.randn(n, 6, device="cuda:0", generator=g) * (2.0 if t % 10 == 9 else 0.7)
        eps0 = {k: v.clone() for k, v in plain._episode_sums.items()}
        o0, r0, te0, tr0, x0 = plain.step(a)
        # term by term: the kernel's contribution of this step (increment of the episode sum) == the torch view x scale
        keep = ~(te0 | tr0)
        plain._active_view = plain._view
        for name, wgt in REWARD_SCALES_V2.items():
            view = getattr(RewardTermViews, "_reward_" + name)(plain) * (wgt * 0.02)
            kern = plain._episode_sums[name] - eps0[name]
            assert float((view - kern)[keep].abs().max()) <= 2e-6, (t, name)
        plain._active_view = None
        o1, r1, te1, tr1, x1 = plus.step(a)
        o2, r2, te2, tr2, x2 = host.step(a)
        o0, o1, o2 = o0["policy"], o1["policy"], o2["policy"]
        assert torch.equal(o0, o1) and torch.equal(te0, te1) and torch.equal(tr0, tr1) and torch.equal(o0, o2) and torch.equal(te0, te2)
        base_z = plain._view.robot.body_link_pos_w[:, plain.base_body_idx[0], 2]          # end of physics, before the reset
        assert torch.allclose(r1 - r0, 0.5 * 0.02 * base_z, atol=2e-6)
        done = te0 | tr0
        resets += int(done.sum())
        # ten terms on the host + three in the kernel == thirteen in the kernel
        assert torch.allclose(r2, r0, rtol=1e-5, atol=3e-6), float((r2 - r0).abs().max())
    assert resets > 0
    log = x1["log"]
    assert "Episode_Reward/base_height" in log and float(log["Episode_Reward/base_height"]) > 0.0
    assert set(x2["log"]) >= {"Episode_Reward/" + k for k in REWARD_SCALES_V2}
    # contact sensor / hook views
    d = plain._contact_sensor.data
    assert d.net_forces_w_history.shape == (n, 5, 12, 3) and d.last_air_time.shape == (n, 12) and d.current_contact_time.shape == (n, 12)
    feet, _ = plain._contact_sensor.find_bodies("foot.*")
    assert float(d.net_forces_w_history[:, 0, feet, 2].max()) > 1.0
    te, tr = plain._get_dones()
    assert torch.equal(te, te0) and torch.equal(tr, tr0) and torch.equal(plain._get_rewards(), r0)
    assert torch.equal(plain._get_observations()["policy"], o0)
    ids = torch.tensor([3, 5, 100], device="cuda:0")
    plain._reset_idx(ids)
    assert torch.all(plain.episode_length_buf[ids] == 0) and float(plain._stepper.state.get("joint_vel")[ids].abs().max()) == 0.0
    fresh = _make_cls(ZbotDirectEnvV2, 64)
    fresh.reset()
    fresh.step(torch.zeros(64, 6, device="cuda:0"))
    with pytest.raises(RuntimeError):
        fresh._contact_sensor.data
    for e in (plain, plus, host, fresh):
        e.close()
