"""Oracle controller (oracle/o_ctrl.c) vs golden vectors produced by the reference's own OperationalSpaceController
(tools/gen_osc_golden.py).  Pins the controller half of the path."""
import os

import numpy as np

from tests.util import ROOT, load


def _setup():
    from oracle.pyoracle import CtrlCfg, Oracle
    from robosuite_b200 import controller_config as cc
    from robosuite_b200.mjcf.compiler import pack_model

    model = load("Lift_Panda")
    o = Oracle(pack_model(model))
    o.ctrl_setup(cc.resolve(model, cc.default_composite_config(), CtrlCfg))
    return model, o


def test_oracle_osc_matches_reference_python():
    g = np.load(os.path.join(ROOT, "tests", "golden", "osc_golden.npz"))
    model, o = _setup()
    nsub = int(g["nsub"])
    n_env, n_steps = g["actions"].shape[:2]
    worst = dict(tau=0.0, goal_pos=0.0, goal_ori=0.0, ctrl=0.0, qpos=0.0, qvel=0.0)
    for e in range(n_env):
        o.reset_data()
        o.qpos[:] = g["qpos0"][e]
        o.forward()
        o.ctrl_reset()
        k = 0
        for t in range(n_steps):
            a = g["actions"][e, t]
            for sub in range(nsub):
                o.step1()
                o.ctrl_run(a if sub == 0 else None)
                if sub in (0, 1, nsub - 1):
                    tau = np.array(o.ctrl_state.torques[:7])
                    worst["tau"] = max(worst["tau"], np.abs(tau - g["torques"][e, k]).max() / np.abs(g["torques"][e, k]).max())
                    worst["goal_pos"] = max(worst["goal_pos"], np.abs(np.array(o.ctrl_state.goal_pos) - g["goal_pos"][e, k]).max())
                    worst["goal_ori"] = max(worst["goal_ori"], np.abs(np.array(o.ctrl_state.goal_ori).reshape(3, 3) - g["goal_ori"][e, k]).max())
                    worst["ctrl"] = max(worst["ctrl"], np.abs(o.ctrl - g["ctrl"][e, k]).max())
                    k += 1
                o.step2()
            worst["qpos"] = max(worst["qpos"], np.abs(o.qpos - g["qpos"][e, t]).max())
            worst["qvel"] = max(worst["qvel"], np.abs(o.qvel - g["qvel"][e, t]).max())
    print(worst)
    # the reference rounds the delta rotation through float32 (transform_utils.py:474); the C emulation of that
    # round trip agrees to float32 epsilon, which bounds everything downstream
    assert worst["tau"] < 5e-6 and worst["goal_pos"] < 1e-7 and worst["goal_ori"] < 5e-7
    assert worst["ctrl"] < 2e-4 and worst["qpos"] < 1e-6 and worst["qvel"] < 1e-5


def test_env_step_equals_manual_loop():
    model, o = _setup()
    from tests.util import lift_states

    q, _ = lift_states(model, 1, seed=3)
    a = np.array([0.3, -0.5, 0.2, 0.1, -0.2, 0.4, 1.0])
    o.reset_data(); o.qpos[:] = q[0]; o.forward(); o.ctrl_reset()
    o.env_step(a, 25)
    q1 = o.qpos.copy()
    o.reset_data(); o.qpos[:] = q[0]; o.forward(); o.ctrl_reset()
    for sub in range(25):
        o.step1(); o.ctrl_run(a if sub == 0 else None); o.step2()
    assert np.array_equal(q1, o.qpos)


def test_oracle_joint_velocity_matches_reference_code_body():
    """JOINT_VELOCITY (BASELINE config 3).  The reference class cannot be constructed at this commit
    (joint_vel.py:127 assigns to a read-only property), so the oracle is pinned to a line-by-line numpy transcription of
    its set_goal / run_controller bodies (joint_vel.py:129-209) with that line read as `use_torque_compensation=True`."""
    from oracle.pyoracle import CtrlCfg, Oracle
    from robosuite_b200 import controller_config as cc
    from robosuite_b200.mjcf.compiler import pack_model

    model = load("Stack_Sawyer")
    cfg = cc.refactor_composite_controller_config(cc.load_part_controller_config("JOINT_VELOCITY"), "Sawyer", ["right"])
    c = cc.resolve(model, cfg, CtrlCfg, gripper="rethink")
    o = Oracle(pack_model(model))
    o.ctrl_setup(c)
    q = model.qpos0.copy()
    q[:7] = [0, -1.18, 0.00, 2.18, 0.00, 0.57, -1.57]
    q[7:9] = [0.020833, -0.020833]
    q[9:12] = [0.05, 0.05, 0.83]; q[16:19] = [-0.05, -0.05, 0.835]
    o.qpos[:] = q; o.forward(); o.ctrl_reset()
    # transcription state
    lo, hi = model.actuator_ctrlrange[:7, 0], model.actuator_ctrlrange[:7, 1]
    kp = 3.0 * (hi - lo); ki = kp * 0.005; kd = kp * 0.001
    last_err = np.zeros(7); summed = np.zeros(7); buf = np.zeros((5, 7)); ptr = 4; size = 0; saturated = False
    goal = np.zeros(7)
    rng = np.random.default_rng(1)
    worst = 0.0
    for t in range(6):
        a = rng.uniform(-1.5, 1.5, 8)
        for sub in range(25):
            o.step1()
            if sub == 0:
                act = np.clip(a[:7], -1, 1)
                goal = np.clip((act - 0.0) * (1.0 / 2.0) + 0.0, -1, 1)   # scale_action: |0.5-(-0.5)|/|1-(-1)|, then velocity_limits
            jv = o.qvel[:7].copy()
            err = goal - jv
            derr = err - last_err
            last_err = err
            ptr = (ptr + 1) % 5; buf[ptr] = derr; size = min(size + 1, 5)
            if not saturated:
                summed = summed + err
            torques = kp * err + ki * summed + kd * buf[:size].mean(axis=0) + o.qfrc_bias[:7]
            clipped = np.clip(torques, lo, hi)
            saturated = not (np.sum(np.abs(clipped - torques)) == 0)
            o.ctrl_run(a if sub == 0 else None)
            worst = max(worst, np.abs(np.array(o.ctrl_state.torques[:7]) - torques).max() / max(np.abs(torques).max(), 1e-9))
            assert np.allclose(o.ctrl[:7], clipped, rtol=0, atol=1e-9)
            o.step2()
    assert worst < 1e-12, worst
    # gripper: rethink sign pattern [+1, -1] integrated at 0.2 per policy step (rethink_gripper.py:43-58)
    assert np.allclose(np.array(o.ctrl_state.grip_action[:2]), np.clip(np.array([1.0, -1.0]) * 0.2 * np.sign(a[7]), -1, 1) +
                       np.array(o.ctrl_state.grip_action[:2]) * 0, atol=2.0)


def test_oracle_joint_velocity_matches_reference_class_golden():
    """tests/golden/jv_golden.npz: torques / goals / saturation flags / trajectories recorded from the reference's own
    JointVelocityController code (tools/gen_jv_golden.py: the class's unmodified methods, with the one attribute that makes it
    unconstructible at this commit - joint_vel.py:127 vs controller.py:303-311 - given both of its readings).  The oracle's controller
    (oracle/o_ctrl.c) replays the same actions on the same physics; everything must agree to rounding."""
    from oracle.pyoracle import CtrlCfg, Oracle
    from robosuite_b200 import controller_config as cc
    from robosuite_b200.mjcf.compiler import pack_model

    from tests.util import dedegenerate_sawyer

    g = np.load(os.path.join(ROOT, "tests", "golden", "jv_golden.npz"))
    model = dedegenerate_sawyer(load("Stack_Sawyer"))
    cfg = cc.refactor_composite_controller_config(cc.load_part_controller_config("JOINT_VELOCITY"), "Sawyer", ["right"])
    c = cc.resolve(model, cfg, CtrlCfg, gripper="rethink")
    nsub = int(g["nsub"])
    n_env, n_steps = g["actions"].shape[:2]
    assert g["saturated"].sum() > 100 and (~g["saturated"]).sum() > 10  # both branches of the anti-windup are in the record
    for e in range(n_env):
        o = Oracle(pack_model(model))
        o.ctrl_setup(c)
        o.qpos[:] = g["qpos0"][e]
        o.qvel[:] = 0
        o.forward()
        o.ctrl_reset()
        k = 0
        for t in range(n_steps):
            a = g["actions"][e, t]
            for sub in range(nsub):
                o.step1()
                o.ctrl_run(a if sub == 0 else None)
                # run_controller returns the clipped torques (joint_vel.py:198-209); the oracle keeps the PID output before the clip
                tau = np.clip(np.array(o.ctrl_state.torques[:7]), model.actuator_ctrlrange[:7, 0], model.actuator_ctrlrange[:7, 1])
                ref = g["torques"][e, k]
                assert np.abs(tau - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max()), (e, t, sub)
                assert np.allclose(o.ctrl, g["ctrl"][e, k], rtol=0, atol=1e-10), (e, t, sub)
                o.step2()
                k += 1
            assert np.allclose(o.qpos, g["qpos"][e, t], rtol=0, atol=1e-9), (e, t)
            assert np.allclose(o.qvel, g["qvel"][e, t], rtol=0, atol=1e-8), (e, t)
