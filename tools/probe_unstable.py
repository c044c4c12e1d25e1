"""Catch the first environments that raise engine flags (or exceed a velocity bound) in a long random-action rollout and save what is needed to
replay them on the CPU oracle: the state K control steps before the event and the actions in between.
usage: python tools/probe_unstable.py Task [robot] [n_env] [steps] [precision] -> gpurun_out/unstable_<Task>.npz"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robosuite_b200 as suite  # noqa: E402

task = sys.argv[1] if len(sys.argv) > 1 else "PickPlace"
robot = sys.argv[2] if len(sys.argv) > 2 else "Panda"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 300
prec = sys.argv[5] if len(sys.argv) > 5 else "f32"
K = 6
env = suite.make(task, robots=robot, num_envs=n, seed=1, horizon=10 ** 9, precision=prec)
sim = env.sim
env.reset()
g = torch.Generator(device=env.device)
g.manual_seed(0)
hist = []  # (state before step, action)
caught, seen = [], torch.zeros(n, dtype=torch.bool, device=env.device)
for t in range(steps):
    a = torch.rand((n, env.action_dim), generator=g, device=env.device, dtype=env.dtype) * 2 - 1
    st = dict(qpos=sim.qpos.clone(), qvel=sim.qvel.clone(), ws=sim.qacc_warmstart.clone(), ctrl=sim.ctrl.clone(),
              gp=sim.ctrl_goal_pos.clone(), go=sim.ctrl_goal_ori.clone(), act=a.clone())
    hist.append(st)
    hist = hist[-K:]
    env.step(a)
    vmax = sim.qvel.abs().nan_to_num(1e9).amax(1)
    bad = ((sim.warn != 0) | (vmax > 60)) & ~seen
    if bool(bad.any()):
        for e in torch.nonzero(bad).flatten().tolist()[:8]:
            caught.append(dict(env=e, t=t, warn=int(sim.warn[e]), vmax=float(vmax[e]), ncon=int(sim.ncon[e]), nefc=int(sim.nefc[e]),
                               qpos=torch.stack([h["qpos"][e] for h in hist]).cpu().numpy(), qvel=torch.stack([h["qvel"][e] for h in hist]).cpu().numpy(),
                               ws=torch.stack([h["ws"][e] for h in hist]).cpu().numpy(), ctrl=torch.stack([h["ctrl"][e] for h in hist]).cpu().numpy(),
                               gp=torch.stack([h["gp"][e] for h in hist]).cpu().numpy(), go=torch.stack([h["go"][e] for h in hist]).cpu().numpy(),
                               act=torch.stack([h["act"][e] for h in hist]).cpu().numpy(),
                               qpos_after=sim.qpos[e].cpu().numpy(), qvel_after=sim.qvel[e].cpu().numpy()))
            print("t", t, "env", e, "warn", int(sim.warn[e]), "vmax %.1f" % float(vmax[e]), "ncon", int(sim.ncon[e]), "nefc", int(sim.nefc[e]), flush=True)
        seen |= bad
    if len(caught) >= 24:
        break
print("steps", t + 1, "envs flagged", int(seen.sum()), "of", n)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"unstable_{task}_{prec}.npz"), n=len(caught),
                    **{f"{i}/{k}": np.asarray(v) for i, c in enumerate(caught) for k, v in c.items()})
