"""Batched environment API: the reference's MujocoEnv surface (`make`, `reset`, `step`, `_get_observations`,
`action_spec`, `reward`, `_check_success`) over N environments living on one GPU.

Behavioural spec, file:line in the reference:
  make / registry            robosuite/environments/base.py:23-56
  reset                      robosuite/environments/base.py:277-347 (+ robots/robot.py:234-300)
  step (25-substep loop)     robosuite/environments/base.py:467-521
  _get_observations          robosuite/environments/base.py:429-465
  action_spec                robosuite/environments/robot_env.py:271-285
The substep loop, controller, observation sampling and task outputs run inside ONE CUDA kernel per control step
(csrc/b2s_kernel.cuh); this module only assembles tensors.
"""
import os
from collections import OrderedDict

import numpy as np

from .. import controller_config as cc
from ..engine import BatchedSim, CtrlCfg
from ..mjcf.compiler import Model, compile_mjcf, load_model

REGISTERED_ENVS = {}
_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets", "models")

# observation scalar ops (must match enum OB_* in csrc/b2s_types.cuh)
OB_QPOS, OB_COS_QPOS, OB_SIN_QPOS, OB_QVEL, OB_QACC, OB_SITE_POS, OB_BODY_POS, OB_BODY_QUAT_XYZW, OB_SITE_QUAT_XYZW, \
    OB_BODY_MINUS_SITE, OB_SITE_MINUS_SITE, OB_BODY_QUAT_REL_SITE_XYZW, OB_ZERO, OB_BODY_MINUS_BODY, OB_REL_POS_LAG, \
    OB_REL_QUAT_LAG = range(16)


def register_env(cls):
    REGISTERED_ENVS[cls.__name__.replace("Batched", "")] = cls
    return cls


def make(env_name, *args, **kwargs):
    """suite.make(env_name, robots=..., num_envs=N, **kw) (environments/base.py:23-42)"""
    if env_name not in REGISTERED_ENVS:
        raise Exception("Environment {} not found. Make sure it is a registered environment among: {}".format(
            env_name, ", ".join(REGISTERED_ENVS)))
    return REGISTERED_ENVS[env_name](*args, **kwargs)


def load_task_model(task, robot, xml=None):
    """Compiled model for task/robot: from a composed MJCF string (reference composer output) when given, else the
    packaged compiled fixture (mesh files do not travel to GPU boxes)."""
    if xml is not None:
        return compile_mjcf(xml)
    path = os.path.join(_ASSETS, f"{task}_{robot}.npz")
    if not os.path.exists(path):
        raise FileNotFoundError(f"no packaged model for {task}/{robot}; pass the composed MJCF via xml=")
    return load_model(path)


class ObsBuilder:
    """Collects (name, modality, op table rows) in the reference's observable order."""

    def __init__(self):
        self.items = []  # (name, modality, [(op,a,b),...])

    def add(self, name, modality, rows):
        self.items.append((name, modality, rows))

    def add_rel_pose(self, obj_key, eef_site, eef_body, modality, pf="robot0_"):
        """`{obj}_to_{pf}eef_pos` (3) and `{obj}_to_{pf}eef_quat` (4): pose of the object in the gripper frame, from the
        `{obj}_pos` / `{obj}_quat` values of the previous sample (manipulation_env.py:268-329).  The slots of those two
        observables are resolved in tables(), so they may be added afterwards, as the reference orders them."""
        b = (eef_site << 8) | (eef_body << 16)
        self.items.append((f"{obj_key}_to_{pf}eef_pos", modality, [("lagpos", obj_key, b | k) for k in range(3)]))
        self.items.append((f"{obj_key}_to_{pf}eef_quat", modality, [("lagquat", obj_key, b | k) for k in range(4)]))

    def tables(self):
        """rows ordered modality by modality (first-seen order), as _get_observations concatenates them"""
        mods = []
        for _, mod, _ in self.items:
            if mod not in mods:
                mods.append(mod)
        ops, slices, mod_slices = [], OrderedDict(), OrderedDict()
        for mod in mods:
            start = len(ops)
            for name, m2, rows in self.items:
                if m2 != mod:
                    continue
                slices[name] = (len(ops), len(ops) + len(rows))
                ops += rows
            mod_slices[mod + "-state"] = (start, len(ops))
        for i, row in enumerate(ops):
            if row[0] in ("lagpos", "lagquat"):
                ps, qs = slices[row[1] + "_pos"][0], slices[row[1] + "_quat"][0]
                ops[i] = (OB_REL_POS_LAG if row[0] == "lagpos" else OB_REL_QUAT_LAG, ps | (qs << 12), row[2])
        arr = np.array(ops, dtype=np.int32).reshape(-1, 3)
        return arr[:, 0], arr[:, 1], arr[:, 2], slices, mod_slices


# bits of sim.warn / info["sim_warn"] (robosuite_b200/csrc: b2s_kernel.cuh, b2s_collide.cuh, b2s_solver.cuh)
SIM_WARN_BITS = {1: "singular mass matrix", 2: "non-finite state in the integrator", 4: "contact capacity overflow (maxcon)",
                 8: "constraint-row capacity overflow (maxefc)", 16: "singular Newton Hessian",
                 32: "diverged state (non-finite / huge qpos, qvel or qacc): data reset to the model defaults, as mj_checkPos/Vel/Acc do",
                 64: "unit-queue watchdog fired: the control step is incomplete (mode 2 only; a library bug, please report)"}


class BatchedMujocoEnv:
    """N copies of one task on one GPU.  All returned arrays are torch.cuda tensors with leading dim N."""

    maxcon = None  # per-environment contact / constraint-row capacity (None: engine defaults 32 / 64); overflow sets warn bit 4
    maxefc = None
    # capacities of the tail kernel's small tier (contacts, rows): what all but ~0.1 % of this task's environment-substeps stay within
    # under random actions (measured: tools/probe_instr.py); None = no tiering
    tier_small = None

    def __init__(self, robots="Panda", num_envs=1, device=0, controller_configs=None, control_freq=20, horizon=500,
                 ignore_done=False, reward_scale=1.0, reward_shaping=False, use_object_obs=True, seed=None,
                 initialization_noise="default", precision="f32", xml=None, has_renderer=False,
                 has_offscreen_renderer=False, use_camera_obs=False, hard_reset=False, lite_physics=True, model=None,
                 kernel_mode="pipeline", sim_cls=None, **kwargs):
        import torch

        if has_renderer or has_offscreen_renderer or use_camera_obs:
            raise NotImplementedError("rendering / camera observations are out of scope of the batched engine")
        if not lite_physics:
            raise NotImplementedError("only lite_physics=True semantics (environments/base.py:494-503) are implemented")
        self.robot_name = robots if isinstance(robots, str) else robots[0]
        self.num_envs = int(num_envs)
        self.control_freq = control_freq
        self.horizon = horizon
        self.ignore_done = ignore_done
        self.reward_scale = reward_scale
        self.reward_shaping = reward_shaping
        self.use_object_obs = use_object_obs
        self.initialization_noise = {"magnitude": 0.02, "type": "gaussian"} if initialization_noise == "default" \
            else (initialization_noise or {"magnitude": 0.0, "type": "gaussian"})
        self.model = model if model is not None else self._load_model(xml)
        self.model_timestep = self.model.opt_timestep
        self.control_timestep = 1.0 / control_freq
        if control_freq <= 0:
            raise ValueError("Control frequency {} is invalid".format(control_freq))
        self.n_substeps = int(self.control_timestep / self.model_timestep)
        caps = {k: v for k, v in (("maxcon", kwargs.get("maxcon", self.maxcon)), ("maxefc", kwargs.get("maxefc", self.maxefc)),
                                  ("tier_small", kwargs.get("tier_small", self.tier_small))) if v}
        # sim_cls: test hook (tests/oracle_sim.py drives the same host code on the CPU oracle); the product path is BatchedSim
        self.sim = (sim_cls or BatchedSim)(self.model, self.num_envs, device=device, precision=precision, **caps)
        self.device = self.sim.torch_device
        self.dtype = self.sim.dtype
        self.composite_controller_config = cc.load_composite_controller_config(controller_configs, self.robot_name)
        self.gripper_type = "panda" if self.robot_name == "Panda" else "rethink"
        self._ctrl_cfg = cc.resolve(self.model, self.composite_controller_config, CtrlCfg, gripper=self.gripper_type)
        self.sim.ctrl_config(self._ctrl_cfg)
        self._setup_references()
        ob = ObsBuilder()
        self._setup_observables(ob)
        op, a, b, self._obs_slices, self._modality_slices = ob.tables()
        self.obs_dim = len(op)
        self.sim.obs_config(op, a, b)
        self._setup_task()
        self.sim.set_export(False)
        # "pipeline": phase kernels + global collision work lists (fastest in steady state); "fused": one kernel per step
        self.sim.set_mode(1 if kernel_mode == "pipeline" else 0)
        self.rng = torch.Generator(device=self.device)
        self.seed = seed
        if seed is not None:
            self.rng.manual_seed(int(seed))
        self.timestep = torch.zeros(self.num_envs, dtype=torch.long, device=self.device)
        self.done = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        self.cur_time = 0.0
        self._max_steps_since_reset = 0  # host-side upper bound of `timestep` (avoids a device sync per step)
        self._host_steps = np.zeros(self.num_envs, dtype=np.int64)  # host mirror of `timestep` (None once a caller resets by device mask)
        self.reset()

    # ---- to be provided by tasks
    def _load_model(self, xml):
        raise NotImplementedError

    def _setup_references(self):
        m = self.model
        jn = m.names["joint"]
        pf = "robot0_"
        self.robot_joints = [i for i, n in enumerate(jn) if n and n.startswith(pf) and int(m.jnt_type[i]) == 3]
        self._ref_joint_pos_indexes = [int(m.jnt_qposadr[j]) for j in self.robot_joints]
        self._ref_joint_vel_indexes = [int(m.jnt_dofadr[j]) for j in self.robot_joints]
        self.gripper_joints = [i for i, n in enumerate(jn) if n and n.startswith("gripper0_")]
        self._ref_gripper_joint_pos_indexes = [int(m.jnt_qposadr[j]) for j in self.gripper_joints]
        self._ref_gripper_joint_vel_indexes = [int(m.jnt_dofadr[j]) for j in self.gripper_joints]
        self.eef_site_id = m.names["site"].index("gripper0_right_grip_site")
        self.eef_body_id = m.names["body"].index("robot0_right_hand")

    def _setup_observables(self, ob):
        """robot proprio observables in the reference's order (robots/robot.py:347-392, 412-484)"""
        mod = "robot0_proprio"
        qp, qv = self._ref_joint_pos_indexes, self._ref_joint_vel_indexes
        ob.add("robot0_joint_pos", mod, [(OB_QPOS, i, 0) for i in qp])
        ob.add("robot0_joint_pos_cos", mod, [(OB_COS_QPOS, i, 0) for i in qp])
        ob.add("robot0_joint_pos_sin", mod, [(OB_SIN_QPOS, i, 0) for i in qp])
        ob.add("robot0_joint_vel", mod, [(OB_QVEL, i, 0) for i in qv])
        ob.add("robot0_joint_acc", mod, [(OB_QACC, i, 0) for i in qv])
        ob.add("robot0_eef_pos", mod, [(OB_SITE_POS, self.eef_site_id, k) for k in range(3)])
        ob.add("robot0_eef_quat", mod, [(OB_BODY_QUAT_XYZW, self.eef_body_id, k) for k in range(4)])
        ob.add("robot0_eef_quat_site", mod, [(OB_SITE_QUAT_XYZW, self.eef_site_id, k) for k in range(4)])
        ob.add("robot0_gripper_qpos", mod, [(OB_QPOS, i, 0) for i in self._ref_gripper_joint_pos_indexes])
        ob.add("robot0_gripper_qvel", mod, [(OB_QVEL, i, 0) for i in self._ref_gripper_joint_vel_indexes])

    def _setup_task(self):
        pass

    def _robot_reset_qpos(self, n):
        """[n, nq] float64 qpos0 with the arm at init_qpos + noise and the gripper at its init pose
        (robots/robot.py:247-259, gripper init_qpos)"""
        import torch

        from .lift import GRIPPER_INIT_QPOS, PANDA_INIT_QPOS, SAWYER_INIT_QPOS

        dev = self.device
        q = self._dev_const("qpos0", self.model.qpos0).repeat(n, 1)
        init = PANDA_INIT_QPOS if self.robot_name == "Panda" else SAWYER_INIT_QPOS
        mag = float(self.initialization_noise["magnitude"])
        if self.initialization_noise["type"] == "gaussian":
            noise = torch.randn((n, len(init)), generator=self.rng, device=dev, dtype=torch.float64) * mag
        else:
            noise = (torch.rand((n, len(init)), generator=self.rng, device=dev, dtype=torch.float64) * 2 - 1) * mag
        q[:, self._dev_index("arm_qpos", self._ref_joint_pos_indexes)] = self._dev_const("arm_init", init) + noise
        q[:, self._dev_index("grip_qpos", self._ref_gripper_joint_pos_indexes)] = self._dev_const("grip_init", GRIPPER_INIT_QPOS[self.robot_name])
        return q

    def _dev_const(self, key, value, dtype=None):
        """device-resident copy of a host constant, uploaded once (the reset path must not touch the host: it runs inside step())"""
        import torch

        c = self.__dict__.setdefault("_dev_consts", {})
        if key not in c:
            c[key] = torch.as_tensor(np.asarray(value), device=self.device, dtype=dtype or torch.float64)
        return c[key]

    def _dev_index(self, key, idx):
        import torch

        return self._dev_const("idx_" + key, np.asarray(idx, dtype=np.int64), dtype=torch.long)

    @staticmethod
    def _place_free_body(q, adr, x, y, z, yaw):
        """free-joint qpos <- position + rotation about z"""
        import torch

        q[:, adr] = x
        q[:, adr + 1] = y
        q[:, adr + 2] = z
        q[:, adr + 3] = torch.cos(yaw / 2)
        q[:, adr + 4] = 0
        q[:, adr + 5] = 0
        q[:, adr + 6] = torch.sin(yaw / 2)

    def _sample_reset_state(self, n):
        raise NotImplementedError

    def _randomize_model(self, mask):
        """per-reset placements the reference writes into MODEL constants (Door: door.py:417-427); mask: bool [N] on the device or None"""

    def reward(self, action=None):
        raise NotImplementedError

    def _check_success(self):
        raise NotImplementedError

    # ---- API
    @property
    def action_dim(self):
        return int(self._ctrl_cfg.action_dim)

    @property
    def action_spec(self):
        """(low, high) bounds (robot_env.py:271-285): OSC input limits + gripper [-1, 1]"""
        c = self._ctrl_cfg
        if c.kind in (2, 3, 4):  # joint-space controllers: per-joint input limits
            n = c.n_arm
            return (np.array(list(c.jv_in_min)[:n] + [-1.0] * (c.action_dim - n)),
                    np.array(list(c.jv_in_max)[:n] + [1.0] * (c.action_dim - n)))
        low = np.array(list(c.input_min)[:6] + [-1.0] * (c.action_dim - 6))
        high = np.array(list(c.input_max)[:6] + [1.0] * (c.action_dim - 6))
        return low, high

    def _fingerpad_geoms(self):
        """left / right fingerpad geom id lists (models/grippers/*_gripper.py `_important_geoms`)"""
        gn = self.model.names["geom"]
        if self.gripper_type == "panda":
            l, r = ["gripper0_right_finger1_pad_collision"], ["gripper0_right_finger2_pad_collision"]
        else:
            l, r = ["gripper0_right_l_fingerpad_g0"], ["gripper0_right_r_fingerpad_g0"]
        return [gn.index(x) for x in l], [gn.index(x) for x in r]

    def reset(self, mask=None, host_mask=None):
        """Re-initialise all (or masked) environments: robot init pose + noise, gripper open, task objects sampled,
        controllers rebuilt (goal <- current eef pose), observations force-updated (environments/base.py:277-347).
        Everything runs on the device without a host round trip: an initial state is sampled for every environment (a few small
        tensor ops) and `b2s_reset_envs` applies it to the masked ones, so a per-step auto-reset costs three short launches.
        host_mask: the same mask as a numpy bool array when the caller has it (keeps the host mirror of the episode clocks exact)."""
        import torch

        q = self._sample_reset_state(self.num_envs).to(self.dtype).contiguous()
        self._randomize_model(None if mask is None else mask.to(device=self.device).bool())
        if mask is None:
            self.timestep.zero_()
            self.done.zero_()
            self._max_steps_since_reset = 0
            self._host_steps = np.zeros(self.num_envs, dtype=np.int64)
            self.sim.reset_envs(None, q)
        else:
            self._reset_mask8 = mask.to(device=self.device, dtype=torch.uint8).contiguous()  # a copy (the caller may pass `self.done`), kept alive
            mask = self._reset_mask8.bool()
            self.timestep.masked_fill_(mask, 0)
            self.done.masked_fill_(mask, False)
            self.sim.reset_envs(self._reset_mask8, q)
            if host_mask is not None and self._host_steps is not None:
                self._host_steps[np.asarray(host_mask, dtype=bool)] = 0
                self._max_steps_since_reset = int(self._host_steps.max())
            else:
                self._host_steps = None  # episode clocks now only known on the device; `_max_steps_since_reset` stays an upper bound
        self._reset_qpos = q
        self.cur_time = 0.0
        return self._get_observations()

    def set_episode_steps(self, steps):
        """Set the per-environment episode clocks (e.g. to stagger the episode phases of a long-running vector environment)."""
        import torch

        h = np.asarray(steps.cpu() if torch.is_tensor(steps) else steps, dtype=np.int64).reshape(self.num_envs)
        self._host_steps = h.copy()
        self.timestep[:] = torch.as_tensor(h, device=self.device)
        self._max_steps_since_reset = int(h.max())

    def host_done(self):
        """numpy bool [N]: which environments have reached the horizon, from the host mirror of the episode clocks; None if unknown"""
        if self._host_steps is None or self.ignore_done:
            return None
        return self._host_steps >= self.horizon

    def reset_to(self, qpos, qvel=None):
        """Put every environment into the given state and do what reset() does afterwards (forward, controllers rebuilt,
        observation cache emptied and force-updated): `set_state` + the tail of environments/base.py:277-347.
        qpos: [nq] or [N, nq]"""
        import torch

        def dev(x):
            return x.to(device=self.device, dtype=self.dtype) if torch.is_tensor(x) else torch.as_tensor(np.asarray(x), dtype=self.dtype, device=self.device)

        q = dev(qpos)
        self.sim.qpos[:] = q if q.ndim == 2 else q.unsqueeze(0).expand(self.num_envs, -1)
        if qvel is None:
            self.sim.qvel[:] = 0
        else:
            v = dev(qvel)
            self.sim.qvel[:] = v if v.ndim == 2 else v.unsqueeze(0).expand(self.num_envs, -1)
        self.sim.qacc[:] = 0
        self.sim.qacc_warmstart[:] = 0
        self.sim.ctrl[:] = 0
        self.sim.time[:] = 0
        self.timestep[:] = 0
        self.done[:] = False
        self.sim.obs_fresh[:] = 1
        self.sim.warn[:] = 0
        self._max_steps_since_reset = 0
        self._host_steps = np.zeros(self.num_envs, dtype=np.int64)
        self.sim.forward()
        self.sim.ctrl_reset(None)
        self.cur_time = 0.0
        return self._get_observations()

    def step(self, action):
        """One control step = n_substeps x {step1, controller, step2} in one kernel launch (base.py:467-521)."""
        import torch

        # an env can only be done once the longest-running one has reached the horizon: no device sync before that
        if not self.ignore_done and self._max_steps_since_reset >= self.horizon:
            hd = self.host_done()
            if bool(hd.any()) if hd is not None else bool(self.done.any()):
                raise ValueError("executing action in terminated episode")
        self._max_steps_since_reset += 1
        if self._host_steps is not None:
            self._host_steps += 1
        action = torch.as_tensor(action, dtype=self.dtype, device=self.device).contiguous()
        assert action.shape == (self.num_envs, self.action_dim), "environment got invalid action dimension -- expected {}, got {}".format(
            (self.num_envs, self.action_dim), tuple(action.shape))
        self.timestep += 1
        self.sim.env_step(action, self.n_substeps)
        self.cur_time += self.control_timestep
        reward = self.reward(action)
        self.done = (self.timestep >= self.horizon) & (not self.ignore_done)
        # (an environment whose state diverged was reset to the model defaults by the engine, like mj_resetData after mj_checkPos / Vel / Acc;
        # as in the reference the episode simply continues from there - info["sim_warn"] bit 32 tells the caller)
        # per-environment engine flags since the last reset, as a device tensor (no host sync here; see SIM_WARN_BITS): a non-zero entry means
        # the episode is no longer a faithful MuJoCo rollout (capacity overflow, singular mass matrix / Hessian, diverged state)
        return self._get_observations(), reward, self.done, {"sim_warn": self.sim.warn}

    def check_sim_warnings(self):
        """Host-side check of the engine flags (one device sync): raises SimulationError naming the flags and how many
        environments carry them.  The reference surfaces the same conditions as mujoco warnings / MujocoException."""
        from ..errors import SimulationError

        w = self.sim.warn
        if bool((w != 0).any()):
            bits = {name: int(((w & bit) != 0).sum()) for bit, name in SIM_WARN_BITS.items()}
            raise SimulationError("engine flags raised: " + ", ".join(f"{k} in {v} envs" for k, v in bits.items() if v))

    def _get_observations(self):
        """OrderedDict of per-observable tensors plus the per-modality concatenations (base.py:429-465)."""
        obs = self.sim.obs
        out = OrderedDict()
        for name, (a, b) in self._obs_slices.items():
            out[name] = obs[:, a:b]
        for name, (a, b) in self._modality_slices.items():
            out[name] = obs[:, a:b]
        return out

    def flat_obs(self):
        """[N, obs_dim] tensor aliasing the kernel's observation buffer (GymWrapper-style flattening)"""
        return self.sim.obs

    def get_state(self):
        return self.sim.get_state()

    def close(self):
        self.sim.close()
