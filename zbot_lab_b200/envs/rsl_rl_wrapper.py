"""``RslRlVecEnvWrapper`` -- the rsl_rl VecEnv adapter the reference wraps its env in
(``scripts/rsl_rl/train.py:181``; semantics: SURVEY.md Appendix B.6).  ``isaaclab_rl`` is not
installed in this image, so the adapter is restated here with the same surface."""
from __future__ import annotations

import torch


class ObsDict(dict):
    """Observation groups keyed by name -- the minimal ``TensorDict`` surface rsl_rl >= 3.0 relies on
    (``obs["policy"]``, ``.to(device)``); ``tensordict`` is not installed in this image."""

    def to(self, device):
        return ObsDict({k: v.to(device) for k, v in self.items()})

    @property
    def batch_size(self):
        return next(iter(self.values())).shape[:1]

    @property
    def device(self):
        return next(iter(self.values())).device


def policy_obs(obs) -> torch.Tensor:
    """Accept both conventions: a plain tensor (rsl_rl 2.x) or an observation dict (rsl_rl >= 3.0)."""
    return obs["policy"] if isinstance(obs, dict) else obs


class RslRlVecEnvWrapper:
    def __init__(self, env, clip_actions: float | None = None):
        self.env = env
        self.clip_actions = clip_actions
        u = env.unwrapped
        self.num_envs = u.num_envs
        self.device = u.device
        self.max_episode_length = u.max_episode_length
        self.num_actions = int(u.single_action_space.shape[0])
        self.num_obs = int(u.single_observation_space["policy"].shape[0])
        self.num_privileged_obs = 0
        self._obs = None
        self.env.reset()

    def __str__(self):
        return f"<{type(self).__name__}{self.env}>"

    @property
    def cfg(self):
        return self.unwrapped.cfg

    @property
    def unwrapped(self):
        return self.env.unwrapped

    @property
    def render_mode(self):
        return self.env.render_mode

    @property
    def observation_space(self):
        return self.env.observation_space

    @property
    def action_space(self):
        return self.env.action_space

    @property
    def episode_length_buf(self) -> torch.Tensor:
        return self.unwrapped.episode_length_buf

    @episode_length_buf.setter
    def episode_length_buf(self, value: torch.Tensor):
        self.unwrapped.episode_length_buf = value

    def seed(self, seed: int = -1) -> int:
        return self.unwrapped.seed(seed)

    def get_observations(self) -> ObsDict:
        """rsl_rl >= 3.0 convention (scripts/rsl_rl/train.py:59 requires 3.0.1): the observation groups."""
        return ObsDict({"policy": self.unwrapped._stepper.observe().clone()})

    def reset(self):
        obs_dict, extras = self.env.reset()
        return ObsDict(obs_dict), extras

    def step(self, actions: torch.Tensor):
        if self.clip_actions is not None:
            actions = torch.clamp(actions, -self.clip_actions, self.clip_actions)
        obs_dict, rew, terminated, truncated, extras = self.env.step(actions)
        dones = (terminated | truncated).to(dtype=torch.long)
        extras["observations"] = obs_dict
        if not self.unwrapped.cfg.is_finite_horizon:
            extras["time_outs"] = truncated
        return ObsDict(obs_dict), rew, dones, extras

    def close(self):
        return self.env.close()
