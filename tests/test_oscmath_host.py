"""CPU: the controller arithmetic the device runs (robosuite_b200/csrc/b2s_oscmath.h, one thread per environment in
ctrl_osc_kernel) compiled for the host and checked against the oracle's OSC (oracle/o_ctrl.c), which is pinned to the reference's
OperationalSpaceController (tests/test_oracle_ctrl.py).  Covers ordinary poses, coupled / uncoupled mode, Sawyer, and poses AT a
kinematic singularity, where numpy's pinv cut-off (control_utils.py:74-76) decides the result."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.util import ROOT, lift_states, load

_SO = os.path.join(ROOT, "tests", "csrc", "libosc_host.so")


def _lib():
    src = os.path.join(ROOT, "tests", "csrc", "osc_host.cpp")
    hdr = os.path.join(ROOT, "robosuite_b200", "csrc", "b2s_oscmath.h")
    if not os.path.exists(_SO) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(_SO):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", _SO, src])
    L = C.CDLL(_SO)
    dp = C.POINTER(C.c_double)
    L.osc_host_torques.argtypes = [C.c_int] + [dp] * 14 + [dp, dp, C.c_double, C.c_int, dp]
    L.osc_host_torques.restype = None
    return L


def _p(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def host_torques(L, model, o, cfg):
    """inputs exactly as ctrl_osc_kernel gathers them from the engine's arrays after step1"""
    na = cfg.n_arm
    dofs = [cfg.arm_dof[k] for k in range(na)]
    qadr = [cfg.arm_qpos[k] for k in range(na)]
    eb, bb = int(model.site_bodyid[cfg.eef_site]), int(model.site_bodyid[cfg.base_site])
    st = o.ctrl_state
    kp = np.array([cfg.kp[k] for k in range(6)])
    kd = 2 * np.sqrt(kp) * np.array([cfg.damping_ratio[k] for k in range(6)])
    keep = []
    args = []
    for a in (o.cdof[dofs], o.site_xpos[cfg.eef_site], o.site_xmat[cfg.eef_site], o.site_xpos[cfg.base_site], o.site_xmat[cfg.base_site],
              np.array(st.goal_pos[:3]), np.array(st.goal_ori[:9]), o.cvel[eb], o.cvel[bb], o.M[np.ix_(dofs, dofs)], o.qfrc_bias[dofs],
              o.qpos[qadr], o.qvel[dofs], np.array(st.initial_joint[:na]), kp, kd):
        arr, ptr = _p(a)
        keep.append(arr); args.append(ptr)
    tau = np.zeros(na)
    L.osc_host_torques(na, *args[:14], args[14], args[15], float(cfg.null_kp), int(cfg.uncouple_pos_ori), tau.ctypes.data_as(C.POINTER(C.c_double)))
    return tau


def _setup(name, uncouple=True):
    from oracle.pyoracle import CtrlCfg as OCfg
    from oracle.pyoracle import Oracle
    from robosuite_b200 import controller_config as cc
    from robosuite_b200.mjcf.compiler import pack_model

    model = load(name)
    robot = "Sawyer" if "Sawyer" in name else "Panda"
    comp = cc.load_composite_controller_config(None, robot)
    cfg = cc.resolve(model, comp, OCfg, gripper="rethink" if robot == "Sawyer" else "panda")
    cfg.uncouple_pos_ori = int(uncouple)
    o = Oracle(pack_model(model))
    o.ctrl_setup(cfg)
    return model, o, cfg


@pytest.mark.parametrize("name,uncouple", [("Lift_Panda", True), ("Lift_Panda", False), ("Lift_Sawyer", True)])
def test_host_build_of_device_controller_matches_oracle(name, uncouple):
    L = _lib()
    model, o, cfg = _setup(name, uncouple)
    rng = np.random.default_rng(0)
    worst = 0.0
    for trial in range(6):
        o.reset_data()
        if "Panda" in name:
            q, _ = lift_states(model, 1, seed=trial)
            o.qpos[:] = q[0]
        else:
            o.qpos[:] = model.qpos0
            o.qpos[:7] = np.array([0, -1.18, 0.0, 2.18, 0.0, 0.57, -1.57]) + rng.normal(0, 0.05, 7)
        o.forward(); o.ctrl_reset()
        for t in range(3):  # a few control steps so that velocities, goals and the nullspace term are all non-trivial
            o.env_step(rng.uniform(-1, 1, cfg.action_dim), 10)
        o.step1()
        tau = host_torques(L, model, o, cfg)   # before ctrl_run: same inputs
        o.ctrl_run(None)
        ref = np.array(o.ctrl_state.torques[:cfg.n_arm])
        worst = max(worst, np.abs(tau - ref).max() / max(np.abs(ref).max(), 1e-9))
    print(name, "uncouple" if uncouple else "coupled", "host build of device controller vs oracle: rel err %.3g" % worst)
    assert worst < 1e-9


def test_singular_pose_takes_the_exact_pinv_path_like_numpy():
    """Panda with joints 2, 4 and 6 at 0 (arm stretched straight up: joints 1, 3, 5, 7 share one axis, 2, 4, 6 are parallel): J
    loses rank and lambda_full^-1 = J M^-1 J^T has an eigenvalue ~1e-17 relative: numpy's pinv drops it (control_utils.py:74-76).
    The fast path (Cholesky) would divide by it.  Every joint value is exactly representable in fp32 as well."""
    L = _lib()
    for uncouple in (True, False):
        model, o, cfg = _setup("Lift_Panda", uncouple)
        o.reset_data()
        q, _ = lift_states(model, 1, seed=3)
        o.qpos[:] = q[0]
        o.qpos[:7] = [0.25, 0.0, 0.125, 0.0, 0.0625, 0.0, 0.5]
        o.qvel[:7] = [0.1, -0.2, 0.15, 0.05, -0.1, 0.2, -0.3]
        o.forward(); o.ctrl_reset()
        st = o.ctrl_state
        st.goal_pos[0] += 0.02; st.goal_pos[2] -= 0.03
        o.step1()
        dofs = [cfg.arm_dof[k] for k in range(7)]
        jp, jr = o.jac(o.site_xpos[cfg.eef_site], int(model.site_bodyid[cfg.eef_site]))
        J = np.vstack([jp[:, dofs], jr[:, dofs]])
        lam_inv = J @ np.linalg.inv(o.M[np.ix_(dofs, dofs)]) @ J.T
        ev = np.linalg.eigvalsh(lam_inv)
        assert ev[0] < 1e-13 * ev[-1], ev  # the pose really is singular to working precision
        tau = host_torques(L, model, o, cfg)
        o.ctrl_run(None)
        ref = np.array(st.torques[:7])
        assert np.all(np.isfinite(tau))
        err = np.abs(tau - ref).max() / max(np.abs(ref).max(), 1e-9)
        print("singular pose, uncouple=%s: host build vs oracle rel err %.3g, |tau| max %.3g" % (uncouple, err, np.abs(ref).max()))
        assert err < 1e-8
