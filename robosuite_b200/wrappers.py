"""Batched counterpart of robosuite's GymWrapper (robosuite/wrappers/gym_wrapper.py:26-180).

Same key selection and flattening rule (`object-state` first, then `robot{i}_proprio-state`), same 5-tuple `step` return,
but every array carries a leading environment axis and lives on the GPU, and finished episodes are reset inside `step`
(gymnasium VectorEnv "next-step autoreset is too late for a fused simulator": the observation returned for a finished
environment is the first observation of its next episode, the final one is in `info["final_observation"]`)."""
import numpy as np


class BatchedGymWrapper:
    def __init__(self, env, keys=None, flatten_obs=True, auto_reset=True):
        self.env = env
        self.name = env.robot_name + "_" + type(env).__name__.replace("Batched", "")
        self.reward_range = (0, env.reward_scale)
        if keys is None:  # gym_wrapper.py:52-61
            keys = []
            if env.use_object_obs:
                keys += ["object-state"]
            keys += ["robot0_proprio-state"]
        self.keys = keys
        self.flatten_obs = flatten_obs
        self.auto_reset = auto_reset
        self.num_envs = env.num_envs
        obs = env._get_observations()
        self.obs_dim = int(sum(obs[k].shape[1] for k in self.keys if k in obs))
        low, high = env.action_spec
        self.action_low, self.action_high = np.asarray(low, dtype=np.float32), np.asarray(high, dtype=np.float32)
        self.single_observation_shape = (self.obs_dim,)
        self.single_action_shape = self.action_low.shape

    def _flatten_obs(self, obs_dict):
        import torch

        return torch.cat([obs_dict[k].reshape(self.num_envs, -1) for k in self.keys if k in obs_dict], dim=1)

    def _filter_obs(self, obs_dict):
        return {k: obs_dict[k] for k in self.keys if k in obs_dict}

    def _format(self, obs_dict):
        return self._flatten_obs(obs_dict) if self.flatten_obs else self._filter_obs(obs_dict)

    def reset(self, seed=None, options=None):
        if seed is not None:
            if not isinstance(seed, int):
                raise TypeError("Seed must be an integer type!")
            self.env.rng.manual_seed(seed)
        return self._format(self.env.reset()), {}

    def step(self, action):
        """-> (obs, reward [N], terminated [N] bool, truncated [N] bool, info).  `terminated` is the reference's `done`
        (horizon reached, environments/base.py:513-514); the reference never truncates."""
        import torch

        ob_dict, reward, done, info = self.env.step(action)
        obs = self._format(ob_dict)
        terminated = done.clone()
        if self.auto_reset and not self.env.ignore_done and self.env._max_steps_since_reset >= self.env.horizon and bool(done.any()):
            info = dict(info)
            info["final_observation"] = obs.clone() if self.flatten_obs else {k: v.clone() for k, v in obs.items()}
            obs = self._format(self.env.reset(mask=done))
            # environments that were not reset keep counting from their own timestep
            self.env._max_steps_since_reset = int(self.env.timestep.max())
        return obs, reward, terminated, torch.zeros_like(terminated), info

    def compute_reward(self, achieved_goal=None, desired_goal=None, info=None):
        return self.env.reward()

    def close(self):
        self.env.close()
