"""Batched counterpart of robosuite's GymWrapper (robosuite/wrappers/gym_wrapper.py:26-180).

Same key selection and flattening rule (`object-state` first, then `robot{i}_proprio-state`), same 5-tuple `step` return,
but every array carries a leading environment axis and lives on the GPU, and finished episodes are reset inside `step`
(gymnasium VectorEnv "next-step autoreset is too late for a fused simulator": the observation returned for a finished
environment is the first observation of its next episode, the final one is in `info["final_observation"]`).  `step` never
synchronises with the device: which episodes ended is known on the host (horizon-only termination) and the masked reset is a
device-side launch sequence (`b2s_reset_envs`)."""
import numpy as np


class _Box:
    """duck-typed stand-in for gymnasium.spaces.Box (bounds, shape, dtype, contains, sample) when gymnasium is not installed"""

    def __init__(self, low, high, dtype=np.float32):
        self.low, self.high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
        self.shape, self.dtype = self.low.shape, np.dtype(dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def sample(self, rng=None):
        rng = rng or np.random.default_rng()
        lo, hi = np.where(np.isfinite(self.low), self.low, -1.0), np.where(np.isfinite(self.high), self.high, 1.0)
        return rng.uniform(lo, hi).astype(self.dtype)

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


def _make_spaces(obs_dim, act_low, act_high, num_envs):
    """(single_observation_space, single_action_space, observation_space, action_space) as in gym_wrapper.py:70-85, batched like a
    gymnasium VectorEnv; real gymnasium spaces when the package is importable"""
    hi = np.full((obs_dim,), np.inf, dtype=np.float32)
    try:
        from gymnasium import spaces
        from gymnasium.vector.utils import batch_space

        so, sa = spaces.Box(-hi, hi, dtype=np.float32), spaces.Box(act_low, act_high, dtype=np.float32)
        return so, sa, batch_space(so, num_envs), batch_space(sa, num_envs)
    except ImportError:
        so, sa = _Box(-hi, hi), _Box(act_low, act_high)
        rep = lambda b: _Box(np.tile(b.low, (num_envs, 1)), np.tile(b.high, (num_envs, 1)))
        return so, sa, rep(so), rep(sa)


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
        # gymnasium VectorEnv surface (gym_wrapper.py:70-85 per environment, batched along the leading axis)
        (self.single_observation_space, self.single_action_space, self.observation_space,
         self.action_space) = _make_spaces(self.obs_dim, self.action_low, self.action_high, self.num_envs)
        self.metadata, self.render_mode, self.spec = {"autoreset_mode": "same_step"}, None, None

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
        env = self.env
        if self.auto_reset and not env.ignore_done and env._max_steps_since_reset >= env.horizon:
            # the horizon is the only termination rule (environments/base.py:513-514), so the host mirror of the episode clocks says which
            # environments finished without reading `done` back; the reset itself is a masked device-side launch sequence.  Without the mirror
            # (a caller reset by device mask) the masked reset is enqueued anyway: it is a no-op for an all-false mask
            hd = env.host_done()
            if hd is None or hd.any():
                info = dict(info)
                info["final_observation"] = obs.clone() if self.flatten_obs else {k: v.clone() for k, v in obs.items()}
                info["sim_warn"] = info["sim_warn"].clone()  # the reset clears the flags of the finished episodes
                obs = self._format(env.reset(mask=done, host_mask=hd))
        return obs, reward, terminated, torch.zeros_like(terminated), info

    def compute_reward(self, achieved_goal=None, desired_goal=None, info=None):
        return self.env.reward()

    def close(self):
        self.env.close()
