"""``RslRlVecEnvWrapper`` -- the rsl_rl VecEnv adapter the reference wraps its env in
(``scripts/rsl_rl/train.py:181``; semantics: SURVEY.md Appendix B.6).  ``isaaclab_rl`` is not
installed in this image, so the adapter is restated here with the same surface."""
from __future__ import annotations

import torch


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

    def get_observations(self):
        obs_dict = {"policy": self.unwrapped._stepper.observe().clone()}
        return obs_dict["policy"], {"observations": obs_dict}

    def reset(self):
        obs_dict, _ = self.env.reset()
        return obs_dict["policy"], {"observations": obs_dict}

    def step(self, actions: torch.Tensor):
        if self.clip_actions is not None:
            actions = torch.clamp(actions, -self.clip_actions, self.clip_actions)
        obs_dict, rew, terminated, truncated, extras = self.env.step(actions)
        dones = (terminated | truncated).to(dtype=torch.long)
        extras["observations"] = obs_dict
        if not self.unwrapped.cfg.is_finite_horizon:
            extras["time_outs"] = truncated
        return obs_dict["policy"], rew, dones, extras

    def close(self):
        return self.env.close()
