"""Scripted pick-and-lift episodes on the UNMODIFIED reference stack (running on oracle/mujoco_shim) to pin the staged
reward / grasp / success logic: tests/golden/reward_golden.npz holds reset state, actions and per control step the
reference's reward (shaped), _check_success() and _check_grasp() of the manipulated object.
Build container only (needs /root/reference).  Usage: python tools/gen_reward_golden.py"""
import os, sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_env_golden as g  # noqa: E402

ROOT = g.ROOT


def scripted(task, obj_body, obj_geoms_attr, steps=70, seed=0, site=None, zoff=0.0, **make_kw):  # seed 0: the composed model then equals the committed fixture (cube size is drawn at model creation)
    import robosuite as suite

    kw = dict(reward_shaping=True)
    kw.update(make_kw)
    env = suite.make(task, robots="Panda", has_renderer=False, has_offscreen_renderer=False, use_camera_obs=False,
                     hard_reset=False, control_freq=20, seed=seed, **kw)
    env.reset()
    sim = env.sim
    bid = sim.model.body_name2id(obj_body)
    sid = sim.model.site_name2id(site) if site else None
    eef = env.robots[0].eef_site_id["right"]
    rec = {"qpos0": np.array(sim.data.qpos)}
    acts, rews, succ, grasp, qs = [], [], [], [], []
    for t in range(steps):
        p_obj, p_eef = np.array(sim.data.body_xpos[bid]), np.array(sim.data.site_xpos[eef])
        if sid is not None:
            p_obj = np.array(sim.data.site_xpos[sid])
        p_obj = p_obj + np.array([0, 0, zoff])
        a = np.zeros(7)
        if t < 18:      # above the object, gripper open
            tgt = p_obj + np.array([0, 0, 0.08]); a[6] = -1
        elif t < 32:    # descend
            tgt = p_obj + np.array([0, 0, 0.0]); a[6] = -1
        elif t < 44:    # close
            tgt = p_eef; a[6] = 1
        else:           # lift
            tgt = p_eef + np.array([0, 0, 0.05]); a[6] = 1
        a[:3] = np.clip((tgt - p_eef) / 0.05 * 0.8, -1, 1)
        obs, r, done, info = env.step(a)
        acts.append(a); rews.append(r); succ.append(bool(env._check_success()))
        og = getattr(env, obj_geoms_attr) if isinstance(obj_geoms_attr, str) else obj_geoms_attr(env)
        grasp.append(bool(env._check_grasp(gripper=env.robots[0].gripper, object_geoms=og)))
        qs.append(np.array(sim.data.qpos))
    rec.update(actions=np.array(acts), reward=np.array(rews), success=np.array(succ), grasp=np.array(grasp), qpos=np.array(qs))
    return rec


def scripted_door(steps=110, seed=0):
    """press the handle down (rotates the latch to its stop), then drag it sideways: exercises the reaching and latch terms of
    the shaped Door reward over their whole range (the door itself only opens to ~0.13 rad with this open-gripper script)"""
    import robosuite as suite

    env = suite.make("Door", robots="Panda", has_renderer=False, has_offscreen_renderer=False, use_camera_obs=False,
                     hard_reset=False, reward_shaping=True, control_freq=20, seed=seed)
    env.reset()
    sim = env.sim
    eef, hs = env.robots[0].eef_site_id["right"], env.door_handle_site_id
    rec = {"qpos0": np.array(sim.data.qpos), "body_pos": np.array(sim.model.body_pos), "body_quat": np.array(sim.model.body_quat)}
    acts, rews, succ, qs = [], [], [], []
    for t in range(steps):
        ph, pe = np.array(sim.data.site_xpos[hs]), np.array(sim.data.site_xpos[eef])
        a = np.zeros(7); a[6] = -1
        if t < 25:
            tgt = ph + np.array([0.0, 0.0, 0.08])
        elif t < 50:
            tgt = ph + np.array([0.0, 0.0, -0.03])
        else:
            tgt = pe + np.array([0.04, 0.04, -0.01])
        a[:3] = np.clip((tgt - pe) / 0.05 * 0.8, -1, 1)
        obs, r, done, info = env.step(a)
        acts.append(a); rews.append(r); succ.append(bool(env._check_success())); qs.append(np.array(sim.data.qpos))
    rec.update(actions=np.array(acts), reward=np.array(rews), success=np.array(succ), grasp=np.zeros(steps, dtype=bool), qpos=np.array(qs))
    return rec


if __name__ == "__main__":
    g.install()
    out = {}
    cases = (("Lift", "cube_main", "cube", {}), ("Stack", "cubeA_main", "cubeA", {}),
             ("NutAssemblyRound", "RoundNut_main", lambda env: env.nuts[1], dict(site="RoundNut_handle_site", steps=80)),
             ("PickPlace", "Can_main", lambda env: env.objects[3], dict(zoff=0.01, steps=80)))
    for task, body, geoms, kw in cases:
        rec = scripted(task, body, geoms, **kw)
        for k, v in rec.items():
            out[f"{task}/{k}"] = np.array(v)
        print(task, "max reward %.3f" % rec["reward"].max(), "grasp steps", int(rec["grasp"].sum()), "success steps", int(rec["success"].sum()),
              "lift", float(rec["qpos"][-1][-5] if task == "Lift" else 0))
    rec = scripted("Lift", "cube_main", "cube", reward_shaping=False, reward_scale=3.0)  # sparse reward, non-default scale
    for k, v in rec.items():
        out[f"Lift_sparse/{k}"] = np.array(v)
    print("Lift sparse rewards", sorted(set(np.round(rec["reward"], 6))))
    rec = scripted_door()
    for k, v in rec.items():
        out[f"Door/{k}"] = np.array(v)
    print("Door max reward %.3f" % rec["reward"].max(), "max hinge %.3f" % rec["qpos"][:, -2].max(), "max latch %.3f" % rec["qpos"][:, -1].max())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reward_golden.npz"), **out)
