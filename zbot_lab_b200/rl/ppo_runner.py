"""Minimal on-policy PPO runner with the ``rsl_rl.runners.OnPolicyRunner`` surface the reference's
training script uses (``scripts/rsl_rl/train.py:185-205``: ctor ``(env, cfg_dict, log_dir, device)``,
``learn(num_learning_iterations, init_at_random_ep_len)``, ``load``, ``save``,
``get_inference_policy``, ``add_git_repo_to_log``).  rsl_rl is not installed in this image; this is a
from-scratch restatement of standard PPO with the hyper-parameters of ``PPORunnerCfgV2``
(``agents/rsl_rl_ppo_cfg.py:65-91``): 24 steps/env, 3x128 ELU actor and critic, GAE(0.99, 0.95),
clip 0.2, 5 epochs x 4 mini-batches, adaptive learning rate on KL 0.01, grad-norm 1.0.

The UPDATE phase is plain PyTorch (autograd over cuBLAS GEMMs).  The ROLLOUT phase (BASELINE
configs[4]) runs, per step, three launches of the repo's own library: ``zbot_policy_act`` (actor + Gaussian sample +
log-prob + critic + rollout-buffer stores, reading the live nn.Linear weights in place), the fused env step, and
``zbot_rollout_store`` (time-out bootstrap + done flag) -- captured as one CUDA graph.  The torch formulation of the
same rollout (``fused_policy: False`` in the train cfg, and automatically for network shapes the kernel is not built
for or envs that are not this repo's) is kept as the semantic reference the tests compare with.
Multi-GPU: when ``torch.distributed`` is initialised, parameters are broadcast from rank 0 and the
flattened gradients are all-reduced once per mini-batch (what rsl_rl does under ``--distributed``).
"""
from __future__ import annotations

import json
import os
import time
from collections import deque

import types

import torch
import torch.distributed as dist
import torch.nn as nn

from ..envs.rsl_rl_wrapper import policy_obs


def _mlp(inp, hidden, out, act):
    layers, d = [], inp
    for h in hidden:
        layers += [nn.Linear(d, h), act()]
        d = h
    layers.append(nn.Linear(d, out))
    return nn.Sequential(*layers)


_ACT = {"elu": nn.ELU, "relu": nn.ReLU, "tanh": nn.Tanh, "selu": nn.SELU, "lrelu": nn.LeakyReLU}


class ActorCritic(nn.Module):
    def __init__(self, num_obs, num_actions, actor_hidden_dims=(128, 128, 128), critic_hidden_dims=(128, 128, 128),
                 activation="elu", init_noise_std=1.0, **_):
        super().__init__()
        act = _ACT[activation]
        self.actor = _mlp(num_obs, list(actor_hidden_dims), num_actions, act)
        self.critic = _mlp(num_obs, list(critic_hidden_dims), 1, act)
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))

    def dist(self, obs):
        mean = self.actor(obs)
        return torch.distributions.Normal(mean, self.std.clamp(min=1e-6).expand_as(mean), validate_args=False)

    def act_inference(self, obs):
        return self.actor(policy_obs(obs))

    def evaluate(self, obs):
        return self.critic(obs).squeeze(-1)


class OnPolicyRunner:
    def __init__(self, env, train_cfg: dict, log_dir: str | None = None, device="cpu"):
        self.env = env
        self.cfg = train_cfg
        self.alg_cfg = dict(train_cfg.get("algorithm", {}))
        self.policy_cfg = dict(train_cfg.get("policy", {}))
        self.device = torch.device(device)
        self.log_dir = log_dir
        self.num_steps = int(train_cfg.get("num_steps_per_env", 24))
        self.save_interval = int(train_cfg.get("save_interval", 100))
        obs = policy_obs(env.get_observations())
        self.num_obs, self.num_actions = obs.shape[1], env.num_actions
        self.policy = ActorCritic(self.num_obs, self.num_actions, **{k: v for k, v in self.policy_cfg.items()
                                                                   if k != "class_name"}).to(self.device)
        self.lr = float(self.alg_cfg.get("learning_rate", 1e-3))
        self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=self.lr)
        self.alg = types.SimpleNamespace(policy=self.policy, optimizer=self.optimizer)   # `runner.alg.policy` (play.py:161)
        self.gamma, self.lam = float(self.alg_cfg.get("gamma", 0.99)), float(self.alg_cfg.get("lam", 0.95))
        self.clip = float(self.alg_cfg.get("clip_param", 0.2))
        self.epochs = int(self.alg_cfg.get("num_learning_epochs", 5))
        self.minibatches = int(self.alg_cfg.get("num_mini_batches", 4))
        self.value_coef = float(self.alg_cfg.get("value_loss_coef", 1.0))
        self.entropy_coef = float(self.alg_cfg.get("entropy_coef", 0.0))
        self.max_grad_norm = float(self.alg_cfg.get("max_grad_norm", 1.0))
        self.desired_kl = self.alg_cfg.get("desired_kl", 0.01)
        self.schedule = self.alg_cfg.get("schedule", "adaptive")
        self.clipped_value = bool(self.alg_cfg.get("use_clipped_value_loss", True))
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if self.distributed:
            for p in self.policy.parameters():
                dist.broadcast(p.data, src=0)
        # rollout fusion: the whole `num_steps x (policy -> env.step -> store)` loop is captured
        # into ONE CUDA graph and replayed per iteration (no Python / launch overhead between the kernels)
        self.use_cuda_graph = bool(train_cfg.get("use_cuda_graph", self.device.type == "cuda"))
        self._graph = None
        # extras["log"] values are 0-dim VIEWS into the env's statistics ring (one slot per step): the per-step dicts of a
        # rollout stay valid only while the ring does not wrap.  Longer rollouts clone the scalars; the captured graph
        # additionally needs slots T..2T-1 (`_pin_env_cursors`), so it is only used when 2T fits.
        st = getattr(getattr(env, "unwrapped", env), "_stepper", None)
        ring = getattr(st, "stats_ring", None)
        self._ring_slots = int(ring.shape[0]) if ring is not None else None
        self._clone_logs = self._ring_slots is not None and self.num_steps > self._ring_slots
        if self.use_cuda_graph and self._ring_slots is not None and 2 * self.num_steps > self._ring_slots:
            import warnings
            warnings.warn(f"num_steps_per_env = {self.num_steps} needs {2 * self.num_steps} statistics slots for the captured "
                          f"rollout graph, the ring has {self._ring_slots}: running the rollout eagerly")
            self.use_cuda_graph = False
        self._fused = self._fused_policy_supported() if bool(train_cfg.get("fused_policy", True)) else None
        self._seed = int(train_cfg.get("seed", 0) or 0)
        self.current_learning_iteration = 0
        self.history: list[dict] = []
        self.git_status_repos: list[str] = []
        n, T, dev = env.num_envs, self.num_steps, self.device
        self.buf = {
            "obs": torch.zeros(T, n, self.num_obs, device=dev), "act": torch.zeros(T, n, self.num_actions, device=dev),
            "rew": torch.zeros(T, n, device=dev), "done": torch.zeros(T, n, device=dev),
            "val": torch.zeros(T, n, device=dev), "logp": torch.zeros(T, n, device=dev),
            "mu": torch.zeros(T, n, self.num_actions, device=dev), "sigma": torch.zeros(T, n, self.num_actions, device=dev),
        }

    # ------------------------------------------------------------------ rsl_rl surface
    def add_git_repo_to_log(self, repo_file_path):
        self.git_status_repos.append(repo_file_path)

    def get_inference_policy(self, device=None):
        self.policy.eval()
        if device is not None:
            self.policy.to(device)
        return self.policy.act_inference

    def save(self, path, infos=None):
        torch.save({"model_state_dict": self.policy.state_dict(), "optimizer_state_dict": self.optimizer.state_dict(),
                    "iter": self.current_learning_iteration, "infos": infos}, path)

    def load(self, path, load_optimizer=True):
        d = torch.load(path, map_location=self.device, weights_only=False)
        self.policy.load_state_dict(d["model_state_dict"])
        if load_optimizer and "optimizer_state_dict" in d:
            self.optimizer.load_state_dict(d["optimizer_state_dict"])
        self.current_learning_iteration = d.get("iter", 0)
        return d.get("infos")

    # ------------------------------------------------------------------ rollout + update
    def _fused_policy_supported(self):
        """The native stepper when the act / store halves of the rollout can run as the library's own kernels
        (``zbot_policy_act`` / ``zbot_rollout_store``): a CUDA env of this repo, 3 x 128 ELU actor and critic,
        num_obs <= 64, num_actions <= 8.  None otherwise (the torch formulation runs)."""
        if self.device.type != "cuda":
            return None
        st = getattr(getattr(self.env, "unwrapped", self.env), "_stepper", None)
        if st is None or not hasattr(st, "policy_act") or not hasattr(self.env, "env"):
            return None
        pc = self.policy_cfg
        if (list(pc.get("actor_hidden_dims", (128, 128, 128))) != [128, 128, 128]
                or list(pc.get("critic_hidden_dims", (128, 128, 128))) != [128, 128, 128]
                or pc.get("activation", "elu") != "elu" or self.num_obs > 64 or self.num_actions > 8):
            return None
        return st

    def _policy_struct(self):
        """``ZbotPolicy`` over the LIVE parameter storage (Adam and ``load_state_dict`` update it in place, so the
        pointers -- also those baked into a captured graph -- stay valid)."""
        from .. import native
        import ctypes as C
        pol = native.ZbotPolicy()
        for net, wn, bn in ((self.policy.actor, "actor_w", "actor_b"), (self.policy.critic, "critic_w", "critic_b")):
            lin = [m for m in net if isinstance(m, nn.Linear)]
            assert len(lin) == 4
            for i, m in enumerate(lin):
                assert m.weight.is_contiguous() and m.weight.dtype == torch.float32 and m.weight.device == self.device
                getattr(pol, wn)[i] = m.weight.data_ptr()
                getattr(pol, bn)[i] = m.bias.data_ptr()
        pol.std = self.policy.std.data_ptr()
        pol.num_obs, pol.num_actions, pol.hidden, pol.activation = self.num_obs, self.num_actions, 128, 0
        return pol

    @torch.no_grad()
    def _collect_rollout_fused(self, obs):
        """Per step: ``zbot_policy_act`` -> the fused env step (+ its statistics pass) -> ``zbot_rollout_store``."""
        b, ep_infos, st, T = self.buf, [], self._fused, self.num_steps
        pol = self._policy_struct()
        inner, clip = self.env.env, getattr(self.env, "clip_actions", None)
        gamma = 0.0 if getattr(self.env.unwrapped.cfg, "is_finite_horizon", False) else self.gamma
        if obs.dtype != torch.float32 or not obs.is_contiguous():
            obs = obs.float().contiguous()
        for t in range(T):
            st.policy_act(pol, obs, b["obs"][t], b["act"][t], b["logp"][t], b["val"][t], b["mu"][t], b["sigma"][t],
                          seed=self._seed)
            act = b["act"][t] if clip is None else torch.clamp(b["act"][t], -clip, clip)
            obs_dict, rew, term, trunc, infos = inner.step(act)
            obs = policy_obs(obs_dict)
            st.rollout_store(rew, term, trunc, b["val"][t], gamma, b["rew"][t], b["done"][t])
            if "log" in infos:
                log = infos["log"]
                if self._clone_logs:
                    log = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in log.items()}
                ep_infos.append(log)
        return obs, ep_infos

    @torch.no_grad()
    def collect_rollout(self, obs):
        """``num_steps`` x (act -> env.step -> store); returns the last observation and episode infos."""
        if self._fused is not None:
            return self._collect_rollout_fused(obs)
        b, ep_infos = self.buf, []
        for t in range(self.num_steps):
            d = self.policy.dist(obs)
            # mean + std * N(0,1): torch.normal(mean, std_tensor) validates std with a host sync, which would
            # break CUDA-graph capture of the rollout
            act = d.mean + d.stddev * torch.randn_like(d.mean)
            b["obs"][t], b["act"][t] = obs, act
            b["val"][t], b["logp"][t] = self.policy.evaluate(obs), d.log_prob(act).sum(-1)
            b["mu"][t], b["sigma"][t] = d.mean, d.stddev
            obs, rew, dones, infos = self.env.step(act)
            obs = policy_obs(obs)
            rew = rew.clone()
            if "time_outs" in infos:   # bootstrap on truncation
                rew += self.gamma * b["val"][t] * infos["time_outs"].to(rew.dtype)
            b["rew"][t], b["done"][t] = rew, dones.to(rew.dtype)
            if "log" in infos:
                log = infos["log"]
                if self._clone_logs:   # the ring wraps within this rollout: detach the scalars from their slots
                    log = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in log.items()}
                ep_infos.append(log)
        return obs, ep_infos

    # ------------------------------------------------------------------ CUDA-graph rollout
    def _pin_env_cursors(self):
        """Make the env's ring cursors start every rollout from the same position, so the buffers a captured
        graph writes (output ring, statistics slots 0..T-1) are the ones Python hands out afterwards."""
        u = self.env.unwrapped
        u._out_i = len(u._out) - 1
        u._stepper._slot = self.num_steps - 1
        u._check_all_reset = False          # a host sync cannot live inside a graph

    def capture_rollout(self, obs):
        """Warm up on a side stream, then capture one rollout; returns the static (obs_in, last_obs, ep_infos)."""
        assert self.device.type == "cuda"
        self._obs_in = obs.clone()
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._pin_env_cursors()
                last, _ = self.collect_rollout(self._obs_in)
                self._obs_in.copy_(last)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self._pin_env_cursors()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            last, ep_infos = self.collect_rollout(self._obs_in)
            self._obs_in.copy_(last)
        self._graph_ep_infos = ep_infos
        return self._obs_in

    def replay_rollout(self):
        """Replay the captured rollout.  `env.step`'s Python does not run during a replay, so what it does on the host is
        done here: the global step counter advances, and the env's HOST curricula (v4 `my_curriculum` / `range_curriculum`,
        …env_v4.py:138-265; manager `lin_vel_cmd_levels`, mdp/curriculums.py:57-83) are evaluated for the replayed steps.
        The kernel parameters are by-value arguments baked into the graph, so when a curriculum changed them the graph
        is dropped and re-captured before the next rollout (a few times per training run)."""
        self._graph.replay()
        u = self.env.unwrapped
        adv = getattr(u, "advance_host_curricula", None)
        if adv is None:
            u.common_step_counter += self.num_steps
        elif adv(self.num_steps):
            self._graph = None                      # stale parameters inside the captured launches: re-capture
        cur = getattr(u, "curriculum_log", None)
        if cur is not None:                         # python-float log entries were frozen at capture time
            fresh = cur()
            for e in self._graph_ep_infos:
                e.update(fresh)
        return self._obs_in, self._graph_ep_infos

    def _returns(self, last_obs):
        b, T = self.buf, self.num_steps
        with torch.no_grad():
            last_val = self.policy.evaluate(last_obs)
        adv = torch.zeros_like(b["rew"])
        gae = torch.zeros_like(last_val)
        for t in reversed(range(T)):
            nv = last_val if t == T - 1 else b["val"][t + 1]
            nd = 1.0 - b["done"][t]
            delta = b["rew"][t] + self.gamma * nv * nd - b["val"][t]
            gae = delta + self.gamma * self.lam * nd * gae
            adv[t] = gae
        ret = adv + b["val"]
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        return ret, adv

    def update(self, last_obs):
        b = self.buf
        ret, adv = self._returns(last_obs)
        flat = lambda x: x.reshape(-1, *x.shape[2:])
        obs, act, logp_old, val_old = flat(b["obs"]), flat(b["act"]), flat(b["logp"]), flat(b["val"])
        mu_old, sig_old, ret, adv = flat(b["mu"]), flat(b["sigma"]), flat(ret), flat(adv)
        total = obs.shape[0]
        mb = total // self.minibatches
        stats = {"value_loss": 0.0, "surrogate_loss": 0.0, "kl": 0.0}
        for _ in range(self.epochs):
            perm = torch.randperm(total, device=self.device)
            for i in range(self.minibatches):
                idx = perm[i * mb:(i + 1) * mb]
                d = self.policy.dist(obs[idx])
                logp = d.log_prob(act[idx]).sum(-1)
                value = self.policy.evaluate(obs[idx])
                with torch.no_grad():
                    kl = torch.sum(torch.log(d.stddev / sig_old[idx] + 1e-5)
                                   + (sig_old[idx] ** 2 + (mu_old[idx] - d.mean) ** 2) / (2 * d.stddev ** 2) - 0.5, -1).mean()
                    if self.distributed:
                        dist.all_reduce(kl)
                        kl /= dist.get_world_size()
                    if self.schedule == "adaptive" and self.desired_kl is not None:
                        if kl > 2.0 * self.desired_kl:
                            self.lr = max(1e-5, self.lr / 1.5)
                        elif 0.0 < kl < 0.5 * self.desired_kl:
                            self.lr = min(1e-2, self.lr * 1.5)
                        for g in self.optimizer.param_groups:
                            g["lr"] = self.lr
                ratio = torch.exp(logp - logp_old[idx])
                surr = torch.max(-adv[idx] * ratio, -adv[idx] * ratio.clamp(1 - self.clip, 1 + self.clip)).mean()
                if self.clipped_value:
                    vc = val_old[idx] + (value - val_old[idx]).clamp(-self.clip, self.clip)
                    vloss = torch.max((value - ret[idx]) ** 2, (vc - ret[idx]) ** 2).mean()
                else:
                    vloss = ((ret[idx] - value) ** 2).mean()
                loss = surr + self.value_coef * vloss - self.entropy_coef * d.entropy().sum(-1).mean()
                self.optimizer.zero_grad(set_to_none=True)
                loss.backward()
                if self.distributed:
                    grads = [p.grad for p in self.policy.parameters() if p.grad is not None]
                    flat_g = torch.cat([g.reshape(-1) for g in grads])
                    dist.all_reduce(flat_g)
                    flat_g /= dist.get_world_size()
                    o = 0
                    for g in grads:
                        g.copy_(flat_g[o:o + g.numel()].view_as(g))
                        o += g.numel()
                nn.utils.clip_grad_norm_(self.policy.parameters(), self.max_grad_norm)
                self.optimizer.step()
                stats["value_loss"] += float(vloss.detach())
                stats["surrogate_loss"] += float(surr.detach())
                stats["kl"] += float(kl)
        k = self.epochs * self.minibatches
        return {n: v / k for n, v in stats.items()}

    def learn(self, num_learning_iterations: int, init_at_random_ep_len: bool = False):
        if init_at_random_ep_len:   # scripts/rsl_rl/train.py:205
            self.env.episode_length_buf = torch.randint_like(self.env.episode_length_buf,
                                                             high=int(self.env.max_episode_length))
        obs = policy_obs(self.env.get_observations()).to(self.device)
        self.policy.train()
        rewbuf, cur_rew = deque(maxlen=100), torch.zeros(self.env.num_envs, device=self.device)
        if self.log_dir and (not self.distributed or dist.get_rank() == 0):
            os.makedirs(self.log_dir, exist_ok=True)
        start = self.current_learning_iteration
        graphed = self.use_cuda_graph and self.device.type == "cuda"
        for it in range(start, start + num_learning_iterations):
            if graphed and self._graph is None:     # first iteration, or a host curriculum changed the kernel parameters
                obs = self.capture_rollout(obs)
            t0 = time.perf_counter()
            if graphed:
                obs, ep_infos = self.replay_rollout()
            else:
                obs, ep_infos = self.collect_rollout(obs)
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            t1 = time.perf_counter()
            losses = self.update(obs)
            t2 = time.perf_counter()
            self.current_learning_iteration = it + 1
            rec = {"iteration": it, "collection_s": t1 - t0, "learn_s": t2 - t1,
                   "fps": self.num_steps * self.env.num_envs / (t2 - t0),
                   "mean_step_reward": float(self.buf["rew"].mean()), "lr": self.lr, **losses}
            if ep_infos:   # average each key over the rollout, as rsl_rl's logger does
                for key in ep_infos[0]:
                    vals = [torch.as_tensor(e[key], dtype=torch.float32).reshape(-1).to(self.device) for e in ep_infos]
                    rec[key] = float(torch.cat(vals).mean())
            self.history.append(rec)
            if self.log_dir and (not self.distributed or dist.get_rank() == 0):
                with open(os.path.join(self.log_dir, "progress.jsonl"), "a") as f:
                    f.write(json.dumps(rec) + "\n")
                if (it + 1) % self.save_interval == 0:
                    self.save(os.path.join(self.log_dir, f"model_{it + 1}.pt"))
        if self.log_dir and (not self.distributed or dist.get_rank() == 0):
            self.save(os.path.join(self.log_dir, f"model_{self.current_learning_iteration}.pt"))
        return self.history
