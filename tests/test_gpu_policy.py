"""GPU tests of the rollout halves (BASELINE configs[4]): `zbot_policy_act` (actor + Gaussian sample +
log-prob + critic + rollout-buffer stores in one launch) and `zbot_rollout_store`, through the C ABI, against the plain
PyTorch fp32 formulation of the same networks (`rl/ppo_runner.ActorCritic`, the rsl_rl ActorCritic of
`agents/rsl_rl_ppo_cfg.py:65-91`).  Tolerances: FP32 with a different summation order over K <= 128 -- 2e-5 absolute on
O(1) network outputs; the log-probability is recomputed from the STORED action / mean / std, so it is tight (1e-5 relative)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _setup(n, num_obs=23, num_actions=6, seed=0, task_cfg=None):
    from zbot_lab_b200 import native
    from zbot_lab_b200.rl.ppo_runner import ActorCritic
    from zbot_lab_b200.stepper import NativeStepper
    torch.manual_seed(seed)
    st = NativeStepper(n, DEV, task_cfg)
    st.reset_idx(None)
    ac = ActorCritic(num_obs, num_actions, init_noise_std=0.7).to(DEV)
    with torch.no_grad():
        ac.std.copy_(torch.linspace(0.2, 1.3, num_actions))
        for p in ac.parameters():            # biases away from zero, weights O(1/sqrt(K)) as nn.Linear initialises them
            if p.ndim == 1 and p is not ac.std:
                p.uniform_(-0.3, 0.3)
    pol = native.ZbotPolicy()
    for net, wn, bn in ((ac.actor, "actor_w", "actor_b"), (ac.critic, "critic_w", "critic_b")):
        lin = [m for m in net if isinstance(m, torch.nn.Linear)]
        for i, m in enumerate(lin):
            getattr(pol, wn)[i] = m.weight.data_ptr()
            getattr(pol, bn)[i] = m.bias.data_ptr()
    pol.std = ac.std.data_ptr()
    pol.num_obs, pol.num_actions, pol.hidden, pol.activation = num_obs, num_actions, 128, 0
    bufs = dict(obs_out=torch.full((n, num_obs), -7.0, device=DEV), act=torch.zeros(n, num_actions, device=DEV),
                logp=torch.zeros(n, device=DEV), value=torch.zeros(n, device=DEV), mu=torch.zeros(n, num_actions, device=DEV),
                sigma=torch.zeros(n, num_actions, device=DEV))
    return st, ac, pol, bufs


def _act(st, pol, obs, b, seed=11):
    st.policy_act(pol, obs, b["obs_out"], b["act"], b["logp"], b["value"], b["mu"], b["sigma"], seed=seed)
    torch.cuda.synchronize()


@pytest.mark.parametrize("kernel", ["tcgen05 (TMEM accumulators, 3 x TF32 split products, default)", "mma.sync (3 x TF32 split products)",
                                    "cuda-core FFMA2"])
@pytest.mark.parametrize("n,num_obs,num_actions", [(4096, 23, 6), (1000 + 17, 23, 6), (300, 48, 6), (33, 64, 8), (5, 7, 1)])
def test_policy_act_matches_torch_fp32(n, num_obs, num_actions, kernel, monkeypatch):
    """The three builds of the act kernel -- `zbot_policy_act_tc5_kernel` (tcgen05.mma kind::tf32 with the accumulators in TMEM:
    csrc/zbot_policy_tc5.cuh, the default), `zbot_policy_act_tc_kernel` (mma.sync TF32: csrc/zbot_policy_tc.cuh,
    ZBOT_POLICY_TC=1) -- both with every product split into lo*hi + hi*lo + hi*hi, FP32 accumulate -- and
    `zbot_policy_act_kernel<4>` (packed FP32 on the CUDA cores, ZBOT_POLICY_TC=0) against torch FP32 and float64: the same
    bounds for all, i.e. the tensor-core kernels are FP32 to round-off, not TF32."""
    monkeypatch.setenv("ZBOT_POLICY_TC", "2" if kernel.startswith("tcgen05") else "1" if kernel.startswith("mma") else "0")
    st, ac, pol, b = _setup(n, num_obs, num_actions, seed=n)
    obs = (torch.randn(n, num_obs, device=DEV) * 1.5).contiguous()
    _act(st, pol, obs, b)
    with torch.no_grad():
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        mu_ref = ac.actor(obs)
        val_ref = ac.critic(obs).squeeze(-1)
        mu64 = ac.actor.double()(obs.double())
        val64 = ac.critic.double()(obs.double()).squeeze(-1)
        ac.float()
        torch.backends.cuda.matmul.allow_tf32 = prev
    assert torch.equal(b["obs_out"], obs)                                        # the rollout buffer's copy: bit-exact
    assert torch.equal(b["sigma"], ac.std.detach().clamp(min=1e-6).expand(n, num_actions))
    # kernel vs torch fp32 (different summation order) and vs float64: the kernel is as close to float64 as cuBLAS fp32 is
    assert float((b["mu"] - mu_ref).abs().max()) < 2e-5, float((b["mu"] - mu_ref).abs().max())
    assert float((b["value"] - val_ref).abs().max()) < 2e-5
    e_k, e_t = float((b["mu"].double() - mu64).abs().max()), float((mu_ref.double() - mu64).abs().max())
    assert e_k < 1e-5 and e_k < 4 * e_t + 2e-6, (e_k, e_t)
    assert float((b["value"].double() - val64).abs().max()) < 1e-5
    # log-probability of the STORED action under Normal(mu, sigma), summed over actions (rsl_rl get_actions_log_prob)
    d = torch.distributions.Normal(b["mu"], b["sigma"], validate_args=False)
    lp = d.log_prob(b["act"]).sum(-1)
    assert torch.allclose(b["logp"], lp, rtol=1e-5, atol=2e-5), float((b["logp"] - lp).abs().max())
    st.close()


def test_policy_act_draws_are_standard_normal_and_follow_the_device_stream_position():
    n = 65536
    st, ac, pol, b = _setup(n)
    obs = torch.randn(n, 23, device=DEV)
    _act(st, pol, obs, b)
    z = ((b["act"] - b["mu"]) / b["sigma"]).double()
    assert torch.isfinite(z).all()
    m, s = z.mean(0), z.std(0)
    assert float(m.abs().max()) < 0.02 and float((s - 1).abs().max()) < 0.02, (m, s)
    kurt = ((z - m) ** 4).mean(0) / s ** 4
    assert float((kurt - 3).abs().max()) < 0.1, kurt
    c = torch.corrcoef(z.T)                                                   # independent across action dims
    assert float((c - torch.eye(6, device=DEV, dtype=c.dtype)).abs().max()) < 0.02
    assert abs(float(torch.corrcoef(torch.stack([z[:-1, 0], z[1:, 0]]))[0, 1])) < 0.02   # ... and across envs
    # same stream position + seed -> the same draws; another seed -> other draws
    a0 = b["act"].clone()
    _act(st, pol, obs, b)
    assert torch.equal(b["act"], a0)
    _act(st, pol, obs, b, seed=12)
    assert not torch.equal(b["act"], a0)
    # an env step advances the device counter: fresh draws (this is what keeps a replayed rollout graph stochastic)
    st.step(torch.zeros(n, 6, device=DEV))
    _act(st, pol, obs, b)
    assert float((b["act"] - a0).abs().min()) >= 0 and float(((b["act"] - a0).abs() > 1e-6).float().mean()) > 0.999
    st.close()


def test_tensor_core_and_cuda_core_act_kernels_draw_the_same_actions(monkeypatch):
    """Same generator, same slots: the two kernels differ only in the rounding of the network outputs, so means, values, actions
    and log-probabilities agree to FP32 round-off on the same observations and stream position."""
    n = 4096 + 33
    outs = []
    for tc in ("2", "0", "1"):
        monkeypatch.setenv("ZBOT_POLICY_TC", tc)
        st, ac, pol, b = _setup(n, seed=3)
        obs = torch.randn(n, 23, device=DEV)
        _act(st, pol, obs, b)
        outs.append({k: v.clone() for k, v in b.items()})
        st.close()
    a, c, m = outs
    for x in (a, m):
        assert torch.equal(x["obs_out"], c["obs_out"]) and torch.equal(x["sigma"], c["sigma"])
        for k, tol in (("mu", 5e-6), ("value", 5e-6), ("act", 5e-6), ("logp", 5e-5)):
            assert float((x[k] - c[k]).abs().max()) < tol, (k, float((x[k] - c[k]).abs().max()))
    # the two tensor-core kernels evaluate the same split products in the same k order: their hidden layers agree to the last
    # few bits (the accumulation inside one MMA may differ), far below the tolerance against the CUDA-core kernel
    assert float((a["mu"] - m["mu"]).abs().max()) < 2e-6


def test_policy_act_reads_the_live_weights_and_rejects_bad_shapes():
    st, ac, pol, b = _setup(64)
    obs = torch.randn(64, 23, device=DEV)
    _act(st, pol, obs, b)
    mu0 = b["mu"].clone()
    with torch.no_grad():
        ac.actor[6].bias.add_(1.0)            # in-place update, as Adam does
    _act(st, pol, obs, b)
    assert torch.allclose(b["mu"], mu0 + 1.0, atol=1e-6)
    pol.hidden = 256
    with pytest.raises(RuntimeError, match="3 x 128"):
        st.policy_act(pol, obs, b["obs_out"], b["act"], b["logp"], b["value"], b["mu"], b["sigma"])
    pol.hidden, pol.num_obs = 128, 65
    with pytest.raises((RuntimeError, ValueError)):
        st.policy_act(pol, obs, b["obs_out"], b["act"], b["logp"], b["value"], b["mu"], b["sigma"])
    pol.num_obs = 23
    with pytest.raises(ValueError):
        st.policy_act(pol, obs[:, :20].contiguous(), b["obs_out"], b["act"], b["logp"], b["value"], b["mu"], b["sigma"])
    st.close()


def test_rollout_store_is_exact():
    from zbot_lab_b200.stepper import NativeStepper
    n = 5000
    st = NativeStepper(n, DEV)
    g = torch.Generator(device=DEV).manual_seed(1)
    rew, val = torch.randn(n, device=DEV, generator=g), torch.randn(n, device=DEV, generator=g)
    term = (torch.rand(n, device=DEV, generator=g) < 0.2)
    trunc = (torch.rand(n, device=DEV, generator=g) < 0.2)
    ro, do = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    st.rollout_store(rew, term, trunc.view(torch.uint8), val, 0.99, ro, do)
    torch.cuda.synchronize()
    # fmaf(gamma, v, r) in float32 == the float64 expression rounded once
    ref = torch.where(trunc, (rew.double() + np.float32(0.99).astype(np.float64) * val.double()).float(), rew)
    assert torch.equal(ro, ref)
    assert torch.equal(do, (term | trunc).float())
    st.close()


def _runner(n, fused, graph, seed=3, steps=24):
    import zbot_lab_b200.tasks  # noqa: F401
    from zbot_lab_b200.compat import gym_registry as gym
    from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper
    from zbot_lab_b200.rl.ppo_runner import OnPolicyRunner
    cfg = gym.load_cfg_from_registry("zbot-6b-walking-v2", "env_cfg_entry_point")
    cfg.scene.num_envs, cfg.sim.device, cfg.seed = n, DEV, seed
    cfg.check_all_envs_reset = False
    env = gym.make("zbot-6b-walking-v2", cfg=cfg, render_mode=None)
    acfg = gym.load_cfg_from_registry("zbot-6b-walking-v2", "rsl_rl_cfg_entry_point").to_dict()
    acfg["use_cuda_graph"], acfg["fused_policy"], acfg["num_steps_per_env"] = graph, fused, steps
    torch.manual_seed(seed)
    w = RslRlVecEnvWrapper(env)
    r = OnPolicyRunner(w, acfg, log_dir=None, device=DEV)
    w.episode_length_buf = torch.zeros(n, dtype=torch.int64)
    return env, w, r


def test_fused_rollout_equals_the_torch_rollout_teacher_forced():
    """The fused rollout's buffers are what the torch formulation produces when it is fed the same actions: a second env
    instance steps through `buf.act` with the torch policy evaluating every stored observation."""
    n, T = 1024, 24
    env, w, r = _runner(n, fused=True, graph=False)
    assert r._fused is not None
    obs0 = w.get_observations()["policy"].clone()
    last, infos = r.collect_rollout(obs0)
    torch.cuda.synchronize()
    assert len(infos) == T
    env2, w2, r2 = _runner(n, fused=False, graph=False)
    assert r2._fused is None
    r2.policy.load_state_dict(r.policy.state_dict())
    obs = w2.get_observations()["policy"].clone()
    assert torch.equal(obs, obs0)
    b = r.buf
    with torch.no_grad():
        for t in range(T):
            assert torch.equal(b["obs"][t], obs), t
            d = r2.policy.dist(obs)
            assert float((d.mean - b["mu"][t]).abs().max()) < 2e-5
            assert float((r2.policy.evaluate(obs) - b["val"][t]).abs().max()) < 2e-5
            assert torch.allclose(torch.distributions.Normal(b["mu"][t], b["sigma"][t], validate_args=False)
                                  .log_prob(b["act"][t]).sum(-1), b["logp"][t], rtol=1e-5, atol=2e-5)
            o, rew, dones, ex = w2.step(b["act"][t])
            obs = o["policy"]
            rew = rew + r2.gamma * b["val"][t] * ex["time_outs"].float()
            assert torch.allclose(b["rew"][t], rew, rtol=0, atol=1e-6), t
            assert torch.equal(b["done"][t], dones.float()), t
    assert torch.equal(last, obs)
    w.close()
    w2.close()


def test_fused_rollout_graph_replay_draws_fresh_actions_and_trains():
    n = 512
    env, w, r = _runner(n, fused=True, graph=True)
    obs = r.capture_rollout(w.get_observations()["policy"])
    r.replay_rollout()
    torch.cuda.synchronize()
    a1 = r.buf["act"].clone()
    z1 = (r.buf["act"] - r.buf["mu"]) / r.buf["sigma"]
    r.replay_rollout()
    torch.cuda.synchronize()
    z2 = (r.buf["act"] - r.buf["mu"]) / r.buf["sigma"]
    assert not torch.equal(a1, r.buf["act"])
    assert float(((z1 - z2).abs() > 1e-4).float().mean()) > 0.99            # a replay advances the generator
    for z in (z1, z2):
        assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1) < 0.02
    hist = r.learn(num_learning_iterations=3)
    assert len(hist) == 3 and all(np.isfinite(h["surrogate_loss"]) and np.isfinite(h["value_loss"]) for h in hist)
    # the graph reads the live weights: a replay after the updates stores means of the UPDATED actor
    if r._graph is None:
        r.capture_rollout(obs)
    r.replay_rollout()
    torch.cuda.synchronize()
    with torch.no_grad():
        assert float((r.policy.actor(r.buf["obs"][5]) - r.buf["mu"][5]).abs().max()) < 2e-5
    w.close()
