"""Door task (robosuite/environments/manipulation/door.py) on the batched engine."""
import math

import numpy as np

from .base import (OB_BODY_MINUS_SITE, OB_BODY_POS, OB_QPOS, OB_SITE_MINUS_SITE, OB_SITE_POS, BatchedMujocoEnv,
                   load_task_model, register_env)


@register_env
class BatchedDoor(BatchedMujocoEnv):
    """suite.make("Door", robots="Panda", num_envs=N): hinged door with a spring-loaded latch (use_latch=True, the
    reference default).  The reference draws the door pose per reset and writes it into model.body_pos/body_quat
    (door.py:303-318, 417-427); model constants are shared by all environments of a batch here, so the door stands at
    `door_placement` (default: centre of the reference's sampling range) in every environment."""

    table_offset = (-0.2, -0.35, 0.8)  # door.py:177
    maxcon, maxefc = 48, 160
    tier_small = (8, 32)  # small tail tier: see BatchedMujocoEnv.tier_small

    def __init__(self, *args, door_placement=None, **kwargs):
        # (x, y, yaw) relative to table_offset; sampler ranges x [0.07, 0.09], y [-0.01, 0.01], yaw [-pi/2 - 0.25, -pi/2]
        self.door_placement = door_placement if door_placement is not None else (0.08, 0.0, -math.pi / 2 - 0.125)
        super().__init__(*args, **kwargs)

    def _load_model(self, xml):
        m = load_task_model("Door", self.robot_name, xml)
        b = m.names["body"].index("Door_main")
        x, y, yaw = self.door_placement
        # z: reference_pos z - bottom_offset z (door.xml bottom_site at -0.3), placement_samplers.py:277-279
        m.body_pos[b] = [self.table_offset[0] + x, self.table_offset[1] + y, self.table_offset[2] + 0.3]
        m.body_quat[b] = [math.cos(yaw / 2), 0.0, 0.0, math.sin(yaw / 2)]
        return m

    def _setup_references(self):
        super()._setup_references()
        m = self.model
        bn, jn = m.names["body"], m.names["joint"]
        self.door_body_id = bn.index("Door_door")
        self.frame_body_id = bn.index("Door_frame")
        self.latch_body_id = bn.index("Door_latch")
        self.door_handle_site_id = m.names["site"].index("Door_handle")
        self.hinge_qpos_addr = int(m.jnt_qposadr[jn.index("Door_hinge")])
        self.use_latch = "Door_latch_joint" in jn
        if self.use_latch:
            self.handle_qpos_addr = int(m.jnt_qposadr[jn.index("Door_latch_joint")])

    def _setup_observables(self, ob):
        super()._setup_observables(ob)
        if self.use_object_obs:  # door.py:345-398, in the reference's order
            d, h, s = self.door_body_id, self.door_handle_site_id, self.eef_site_id
            ob.add("door_pos", "object", [(OB_BODY_POS, d, k) for k in range(3)])
            ob.add("handle_pos", "object", [(OB_SITE_POS, h, k) for k in range(3)])
            ob.add("hinge_qpos", "object", [(OB_QPOS, self.hinge_qpos_addr, 0)])
            ob.add("door_to_eef_pos", "object", [(OB_BODY_MINUS_SITE, (d << 8) | s, k) for k in range(3)])
            ob.add("handle_to_eef_pos", "object", [(OB_SITE_MINUS_SITE, (h << 8) | s, k) for k in range(3)])
            if self.use_latch:
                ob.add("handle_qpos", "object", [(OB_QPOS, self.handle_qpos_addr, 0)])

    def _setup_task(self):
        left, right = self._fingerpad_geoms()
        self.sim.task_config(self.door_body_id, self.eef_site_id, left, right, [])
        h, s = self.door_handle_site_id, self.eef_site_id
        self.sim.task_table([(OB_SITE_MINUS_SITE, (h << 8) | s, k) for k in range(3)])  # _gripper_to_handle after the step

    def _sample_reset_state(self, n):
        return self._robot_reset_qpos(n)  # door closed, latch at rest (qpos0)

    def _check_success(self):
        """hinge opened beyond 0.3 rad (door.py:429-437); qpos after the step, as the reference reads it"""
        return self.sim.qpos[:, self.hinge_qpos_addr] > 0.3

    def reward(self, action=None):
        """door.py:219-266: 1 if opened; shaping: 0.25 (1 - tanh(10 |handle - eef|)) + latch rotation term"""
        import torch

        success = self._check_success()
        r = success.to(self.dtype)
        if self.reward_shaping:
            shaped = 0.25 * (1 - torch.tanh(10.0 * torch.linalg.norm(self.sim.task_vec, dim=1)))
            if self.use_latch:
                hq = self.sim.qpos[:, self.handle_qpos_addr]
                shaped = shaped + torch.clamp(0.25 * torch.abs(hq / (0.5 * np.pi)), -0.25, 0.25)
            r = torch.where(success, r, shaped.to(self.dtype))
        if self.reward_scale is not None:
            r = r * (self.reward_scale / 1.0)
        return r
