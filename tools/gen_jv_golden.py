"""Golden vectors for JOINT_VELOCITY (BASELINE config 3) from the REFERENCE's own class,
robosuite.controllers.parts.generic.joint_vel.JointVelocityController, driven on the duck-typed sim of gen_osc_golden.py.

The class cannot be constructed as written at the reference's commit: joint_vel.py:127 assigns `self.torque_compensation = kwargs.get(
"use_torque_compensation", True)` onto the read-only property of the base class (controller.py:303-311, which returns the gravity
compensation torques qfrc_bias[qvel_index]), and run_controller uses the same name both as the flag (`if self.torque_compensation:`,
:192) and as the torque vector added to the PID output (:194).  The generator therefore runs the class's own unmodified __init__ / set_goal /
run_controller / reset_goal code under a subclass that only replaces that one attribute by a descriptor giving both readings at once: the
setter stores the flag of line 127, the getter returns qfrc_bias[qvel_index] as an array whose truth value is that flag.  No other line of
the reference is bypassed; this is the reading "PID velocity controller plus gravity compensation torques" of the comment at :191 and of
the sibling controllers (joint_pos.py, joint_tor.py).  Output: tests/golden/jv_golden.npz (build container only).

Usage: python tools/gen_jv_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from tools.gen_mjcf_fixtures import install_stubs  # noqa: E402

install_stubs()
import mujoco  # noqa: E402,F401  (the stub)

from oracle.pyoracle import Oracle  # noqa: E402
from robosuite_b200 import controller_config as cc  # noqa: E402
from robosuite_b200.mjcf.compiler import load_model, pack_model  # noqa: E402
from tools.gen_osc_golden import _Sim  # noqa: E402

SAWYER_INIT = [0.00, -1.18, 0.00, 2.18, 0.00, 0.57, -1.57]


class _FlaggedTorques(np.ndarray):
    """qfrc_bias[qvel_index] whose truth value is the use_torque_compensation flag"""

    def __new__(cls, values, flag):
        obj = np.asarray(values, dtype=np.float64).view(cls)
        obj._flag = bool(flag)
        return obj

    def __array_finalize__(self, obj):
        self._flag = getattr(obj, "_flag", True)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):  # arithmetic yields plain arrays: only the attribute itself carries the flag
        inputs = tuple(np.asarray(x) if isinstance(x, _FlaggedTorques) else x for x in inputs)
        return getattr(ufunc, method)(*inputs, **kwargs)

    def __bool__(self):
        return self._flag


class _TorqueCompensation:
    def __set__(self, obj, value):
        obj.__dict__["_use_torque_compensation"] = bool(value)

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        return _FlaggedTorques(obj.sim.data.qfrc_bias[obj.qvel_index], obj.__dict__.get("_use_torque_compensation", True))


def make_controller(sim, model, arm_j, part, **extra):
    from robosuite.controllers.parts.generic.joint_vel import JointVelocityController

    class JV(JointVelocityController):
        torque_compensation = _TorqueCompensation()

    return JV(sim, joint_indexes={"joints": arm_j, "qpos": [int(model.jnt_qposadr[j]) for j in arm_j],
                                  "qvel": [int(model.jnt_dofadr[j]) for j in arm_j]},
              actuator_range=(model.actuator_ctrlrange[:7, 0], model.actuator_ctrlrange[:7, 1]),
              part_name="right", naming_prefix="robot0_", policy_freq=20, **part, **extra)


def main():
    from tests.util import dedegenerate_sawyer

    model = dedegenerate_sawyer(load_model(os.path.join(ROOT, "tests", "golden", "models", "Stack_Sawyer.npz")))
    o = Oracle(pack_model(model))
    sim = _Sim(o, model)
    arm_j = [i for i, n in enumerate(model.names["joint"]) if n and n.startswith("robot0_") and int(model.jnt_type[i]) == 3]
    part = dict(cc._DEFAULT_JOINT_VELOCITY)
    part.pop("type")
    part.pop("interpolation")
    rng = np.random.default_rng(321)
    n_env, n_steps, nsub = 4, 6, 25
    rec = dict(qpos0=[], actions=[], torques=[], raw_torques=[], goal_vel=[], ctrl=[], qpos=[], qvel=[], saturated=[])
    lo, hi = model.actuator_ctrlrange[:, 0], model.actuator_ctrlrange[:, 1]
    for e in range(n_env):
        q = np.array(model.qpos0)
        q[:7] = np.array(SAWYER_INIT) + rng.normal(0, 0.02, 7)
        q[7:9] = [0.020833, -0.020833]
        q[9:12] = [0.05, 0.05, 0.83]
        q[16:19] = [-0.05, -0.05, 0.835]
        o.reset_data()
        o.qpos[:] = q
        o.qvel[:] = 0
        o.forward()
        ctl = make_controller(sim, model, arm_j, part)
        ctl.reset_goal()
        acts, tqs, raws, gvs, ctrls, qs, vs, sats = [], [], [], [], [], [], [], []
        grip = np.zeros(2)
        for t in range(n_steps):
            # env 3 asks for velocities the torque limits cannot deliver: exercises the saturation flag / anti-windup (joint_vel.py:186-201)
            a = rng.uniform(-1.5, 1.5, 8) if e < 3 else np.concatenate([np.sign(rng.uniform(-1, 1, 7)) * 1.5, [1.0]])
            for sub in range(nsub):
                o.step1()
                if sub == 0:
                    ctl.set_goal(np.clip(a[:7], -1, 1))  # Robot.control clips the action to the controller's input range first
                    grip = np.clip(grip + np.array([1.0, -1.0]) * 0.2 * np.sign(a[7:8]), -1.0, 1.0)  # rethink_gripper.py:43-58
                tau = ctl.run_controller()
                o.ctrl[:7] = np.clip(tau, lo[:7], hi[:7])
                o.ctrl[7:9] = np.clip(0.5 * (hi[7:9] + lo[7:9]) + 0.5 * (hi[7:9] - lo[7:9]) * grip, lo[7:9], hi[7:9])
                tqs.append(np.array(tau))
                gvs.append(np.array(ctl.goal_vel))
                ctrls.append(np.array(o.ctrl))
                sats.append(bool(ctl.saturated))
                o.step2()
            acts.append(a)
            qs.append(np.array(o.qpos))
            vs.append(np.array(o.qvel))
        rec["qpos0"].append(q)
        for k, v in zip(("actions", "torques", "goal_vel", "ctrl", "qpos", "qvel", "saturated"), (acts, tqs, gvs, ctrls, qs, vs, sats)):
            rec[k].append(np.array(v))
    rec.pop("raw_torques")
    out = os.path.join(ROOT, "tests", "golden", "jv_golden.npz")
    np.savez_compressed(out, **{k: np.array(v) for k, v in rec.items()}, nsub=nsub)
    print("wrote", out, {k: np.array(v).shape for k, v in rec.items()}, "saturated substeps:", int(np.sum(rec["saturated"])))


if __name__ == "__main__":
    main()
