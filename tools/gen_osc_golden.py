"""Golden vectors for the controller half of the path, produced by the REFERENCE's own Python:
robosuite.controllers.parts.arm.osc.OperationalSpaceController (osc.py:225-495) driven on a duck-typed `sim`
whose dynamics quantities (site pose, Jacobians, M, qfrc_bias) come from this repo's CPU oracle, i.e. both sides
see identical inputs and only the controller arithmetic differs.  Runs only in the build container
(/root/reference); output committed as tests/golden/osc_golden.npz.

Usage: python tools/gen_osc_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from tools.gen_mjcf_fixtures import install_stubs  # noqa: E402

install_stubs()
import mujoco  # noqa: E402  (the stub)

from oracle.pyoracle import CtrlCfg, Oracle  # noqa: E402
from robosuite_b200 import controller_config as cc  # noqa: E402
from robosuite_b200.mjcf.compiler import load_model, pack_model  # noqa: E402
from tests.util import lift_states  # noqa: E402


class _Model:
    def __init__(self, model):
        self._m = model
        self.nv = model.nv
        self._model = None

    def joint_id2name(self, i):
        return self._m.names["joint"][i]

    def site_name2id(self, n):
        return self._m.names["site"].index(n)


class _Data:
    def __init__(self, o, model):
        self.o, self._m = o, model
        self.qM = None

    qpos = property(lambda s: s.o.qpos)
    qvel = property(lambda s: s.o.qvel)
    qfrc_bias = property(lambda s: s.o.qfrc_bias)
    site_xpos = property(lambda s: s.o.site_xpos)
    site_xmat = property(lambda s: s.o.site_xmat)

    def _jac(self, name):
        sid = self._m.names["site"].index(name)
        return self.o.jac(self.o.site_xpos[sid], self._m.site_bodyid[sid])

    def get_site_jacp(self, name):
        return self._jac(name)[0]

    def get_site_jacr(self, name):
        return self._jac(name)[1]

    def get_site_xvelp(self, name):
        return self._jac(name)[0] @ self.o.qvel

    def get_site_xvelr(self, name):
        return self._jac(name)[1] @ self.o.qvel


class _Sim:
    def __init__(self, o, model):
        self.o = o
        self.model = _Model(model)
        self.data = _Data(o, model)

    def forward(self):
        self.o.forward()


def main():
    from robosuite.controllers.parts.arm.osc import OperationalSpaceController

    model = load_model(os.path.join(ROOT, "tests", "golden", "models", "Lift_Panda.npz"))
    o = Oracle(pack_model(model))
    sim = _Sim(o, model)
    mujoco.mj_fullM = lambda mm, dst, qM: dst.__setitem__(slice(None), o.M)
    cfg = cc.resolve(model, cc.default_composite_config(), CtrlCfg)
    arm_j = [i for i, n in enumerate(model.names["joint"]) if n and n.startswith("robot0_joint")]
    part = dict(cc._DEFAULT_OSC_POSE)
    part.pop("type")
    rng = np.random.default_rng(123)
    n_env, n_steps, nsub = 6, 4, 25
    q0, v0 = lift_states(model, n_env, seed=7)
    rec = dict(qpos0=q0, actions=[], torques=[], goal_pos=[], goal_ori=[], ctrl=[], qpos=[], qvel=[])
    for e in range(n_env):
        o.reset_data()
        o.qpos[:] = q0[e]
        o.forward()
        ctl = OperationalSpaceController(
            sim, ref_name="gripper0_right_grip_site",
            joint_indexes={"joints": arm_j, "qpos": [int(model.jnt_qposadr[j]) for j in arm_j],
                           "qvel": [int(model.jnt_dofadr[j]) for j in arm_j]},
            actuator_range=(model.actuator_ctrlrange[:7, 0], model.actuator_ctrlrange[:7, 1]),
            part_name="right", naming_prefix="robot0_", policy_freq=20, ndim=7, **part)
        bs = model.names["site"].index("robot0_right_center")
        ctl.update_origin(np.array(o.site_xpos[bs]), np.array(o.site_xmat[bs]).reshape(3, 3))
        ctl.reset_goal()
        acts, tqs, gps, gos, ctrls, qs, vs = [], [], [], [], [], [], []
        grip = np.zeros(2)
        for t in range(n_steps):
            a = rng.uniform(-1, 1, 7)
            if t == 1:
                a[:6] *= 1.7  # exercise the input clipping
            for sub in range(nsub):
                o.step1()
                ctl.update_origin(np.array(o.site_xpos[bs]), np.array(o.site_xmat[bs]).reshape(3, 3))
                if sub == 0:
                    ctl.set_goal(a[:6])
                    grip = np.clip(grip + np.array([-1.0, 1.0]) * 0.2 * np.sign(a[6:7]), -1.0, 1.0)
                tau = ctl.run_controller()
                lo, hi = model.actuator_ctrlrange[:, 0], model.actuator_ctrlrange[:, 1]
                o.ctrl[:7] = np.clip(tau, lo[:7], hi[:7])
                o.ctrl[7:9] = np.clip(0.5 * (hi[7:9] + lo[7:9]) + 0.5 * (hi[7:9] - lo[7:9]) * grip, lo[7:9], hi[7:9])
                if sub in (0, 1, nsub - 1):
                    tqs.append(np.array(tau))
                    gps.append(np.array(ctl.goal_pos))
                    gos.append(np.array(ctl.goal_ori))
                    ctrls.append(np.array(o.ctrl))
                o.step2()
            acts.append(a)
            qs.append(np.array(o.qpos))
            vs.append(np.array(o.qvel))
        for k, v in zip(("actions", "torques", "goal_pos", "goal_ori", "ctrl", "qpos", "qvel"),
                        (acts, tqs, gps, gos, ctrls, qs, vs)):
            rec[k].append(np.array(v))
    out = os.path.join(ROOT, "tests", "golden", "osc_golden.npz")
    np.savez_compressed(out, **{k: np.array(v) for k, v in rec.items()}, nsub=nsub)
    print("wrote", out, {k: np.array(v).shape for k, v in rec.items()})


if __name__ == "__main__":
    main()
