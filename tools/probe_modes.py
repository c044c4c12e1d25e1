"""fused vs pipeline vs oracle on a contact-rich scripted rollout (gripper closing, arms pushing down)"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from robosuite_b200 import controller_config as cc
from robosuite_b200.engine import BatchedSim, CtrlCfg
from robosuite_b200.mjcf.compiler import pack_model
from oracle.pyoracle import Oracle, CtrlCfg as OCfg
from tests.util import lift_states, load
model = load("Lift_Panda"); n = 16; T = 40
q, v = lift_states(model, n, seed=21)
rng = np.random.default_rng(3)
actions = rng.uniform(-1, 1, size=(T, n, 7)); actions[:, :, 6] = 1.0; actions[8:, : n // 2, :3] = [0.0, 0.0, -1.0]
traj = {}
for mode in (0, 1):
    sim = BatchedSim(model, n, precision="f32"); sim.ctrl_config(cc.resolve(model, cc.default_composite_config(), CtrlCfg))
    sim.set_export(False); sim.set_mode(mode)
    sim.qpos.copy_(torch.as_tensor(q, dtype=torch.float32)); sim.forward(); sim.ctrl_reset()
    tr = []
    for t in range(T):
        sim.env_step(torch.as_tensor(actions[t], dtype=torch.float32, device=sim.torch_device).contiguous(), 25)
        tr.append(sim.qpos.cpu().numpy().astype(np.float64))
    traj[mode] = np.array(tr); print("mode", mode, "warn", int(sim.warn.abs().max())); sim.close()
otr = np.zeros((T, n, model.nq))
for e in range(n):
    o = Oracle(pack_model(model)); o.ctrl_setup(cc.resolve(model, cc.default_composite_config(), OCfg))
    o.qpos[:] = q[e]; o.forward(); o.ctrl_reset()
    for t in range(T):
        o.env_step(actions[t, e], 25); otr[t, e] = o.qpos
for t in (0, 4, 9, 14, 19, 29, 39):
    d01 = np.abs(traj[0][t] - traj[1][t]).max(1); d0o = np.abs(traj[0][t] - otr[t]).max(1); d1o = np.abs(traj[1][t] - otr[t]).max(1)
    print("t=%2d fused-vs-pipe max %.2e (envs>1e-4: %d) | fused-vs-oracle max %.2e median %.2e | pipe-vs-oracle max %.2e median %.2e" % (
        t, d01.max(), (d01 > 1e-4).sum(), d0o.max(), np.median(d0o), d1o.max(), np.median(d1o)))
print("per-env at t=39 fused-vs-oracle", np.round(np.abs(traj[0][39] - otr[39]).max(1), 4))
print("per-env at t=39 pipe-vs-oracle ", np.round(np.abs(traj[1][39] - otr[39]).max(1), 4))
