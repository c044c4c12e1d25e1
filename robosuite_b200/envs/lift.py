"""Lift task (robosuite/environments/manipulation/lift.py) on the batched engine."""
import math

import numpy as np

from .base import (OB_BODY_MINUS_SITE, OB_BODY_POS, OB_BODY_QUAT_XYZW, BatchedMujocoEnv, load_task_model,
                   register_env)

PANDA_INIT_QPOS = np.array([0, np.pi / 16.0, 0.00, -np.pi / 2.0 - np.pi / 3.0, 0.00, np.pi - 0.2, np.pi / 4])
SAWYER_INIT_QPOS = np.array([0.00, -1.18, 0.00, 2.18, 0.00, 0.57, -1.57])
GRIPPER_INIT_QPOS = {"Panda": [0.020833, -0.020833], "Sawyer": [0.020833, -0.020833]}


@register_env
class BatchedLift(BatchedMujocoEnv):
    """suite.make("Lift", robots="Panda", num_envs=N): table arena + one cube, sparse/shaped lifting reward."""

    # capacities: the small tier holds every contact / row count seen in 10^5 random-action environment-substeps (max 12 / 47 observed,
    # mean 4 / 21); the large tier exists for the 1-in-10^6 pile-ups (an arm lying on the table next to the cube)
    maxcon, maxefc = 48, 128
    tier_small = (12, 44)

    table_offset = (0.0, 0.0, 0.8)  # lift.py:146

    def _load_model(self, xml):
        return load_task_model("Lift", self.robot_name, xml)

    def _setup_references(self):
        super()._setup_references()
        m = self.model
        self.cube_body_id = m.names["body"].index("cube_main")
        self.cube_joint = m.names["joint"].index("cube_joint0")
        self.cube_qadr = int(m.jnt_qposadr[self.cube_joint])
        self.cube_geoms = [i for i, n in enumerate(m.names["geom"]) if n and n.startswith("cube_g")]
        self.cube_half_height = float(m.geom_size[self.cube_geoms[0], 2])

    def _setup_observables(self, ob):
        super()._setup_observables(ob)
        if self.use_object_obs:  # lift.py:356-399
            b, s = self.cube_body_id, self.eef_site_id
            ob.add("cube_pos", "object", [(OB_BODY_POS, b, k) for k in range(3)])
            ob.add("cube_quat", "object", [(OB_BODY_QUAT_XYZW, b, k) for k in range(4)])
            ob.add("gripper_to_cube_pos", "object", [(OB_BODY_MINUS_SITE, (b << 8) | s, k) for k in range(3)])

    def _setup_task(self):
        left, right = self._fingerpad_geoms()
        self.sim.task_config(self.cube_body_id, self.eef_site_id, left, right, self.cube_geoms)

    def _sample_reset_state(self, n):
        """robot: init_qpos + N(0, 0.02^2) (robots/robot.py:247-259); cube: x,y ~ U[-0.03,0.03], yaw ~ U[0,2pi),
        z = table + 0.01 + half height (lift.py:311-336, placement_samplers.py:221-309)"""
        import torch

        q = self._robot_reset_qpos(n)
        u = torch.rand((n, 3), generator=self.rng, device=self.device, dtype=torch.float64)
        a = self.cube_qadr
        q[:, a] = self.table_offset[0] + (u[:, 0] * 2 - 1) * 0.03
        q[:, a + 1] = self.table_offset[1] + (u[:, 1] * 2 - 1) * 0.03
        q[:, a + 2] = self.table_offset[2] + 0.01 + self.cube_half_height
        yaw = u[:, 2] * 2 * math.pi
        q[:, a + 3] = torch.cos(yaw / 2)
        q[:, a + 4] = 0
        q[:, a + 5] = 0
        q[:, a + 6] = torch.sin(yaw / 2)
        return q

    def _check_success(self):
        """cube higher than the table top + 0.04 (lift.py:433-444); uses the pose of the last step1 like the reference"""
        return self.sim.task_out[:, 0] > self.table_offset[2] + 0.04

    def reward(self, action=None):
        """lift.py:224-273: 2.25 if lifted, else (shaping) reaching 1 - tanh(10 d) + 0.25 grasp; scaled by scale/2.25"""
        import torch

        t = self.sim.task_out
        success = self._check_success()
        r = torch.where(success, torch.full_like(t[:, 0], 2.25), torch.zeros_like(t[:, 0]))
        if self.reward_shaping:
            shaped = 1 - torch.tanh(10.0 * t[:, 1]) + 0.25 * t[:, 2]
            r = torch.where(success, r, shaped)
        if self.reward_scale is not None:
            r = r * (self.reward_scale / 2.25)
        return r
