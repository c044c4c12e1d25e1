import sys, numpy as np, torch
sys.path.insert(0, ".")
import robosuite_b200 as suite
from oracle.pyoracle import CtrlCfg as OCfg, Oracle
from robosuite_b200 import controller_config as cc
from robosuite_b200.mjcf.compiler import pack_model
n = 6
cfg = cc.refactor_composite_controller_config(cc.load_part_controller_config("JOINT_VELOCITY"), "Sawyer", ["right"])
env = suite.make("Stack", robots="Sawyer", num_envs=n, seed=5, controller_configs=cfg, horizon=100, kernel_mode="fused")
model = env.model; sim = env.sim
q0 = sim.qpos.cpu().numpy().astype(np.float64)
sim.set_export(True)
os_ = []
for e in range(n):
    o = Oracle(pack_model(model)); o.ctrl_setup(cc.resolve(model, cfg, OCfg, gripper="rethink"))
    o.qpos[:] = q0[e]; o.forward(); o.ctrl_reset(); os_.append(o)
# forward comparison
sim.forward(); torch.cuda.synchronize()
for name, arr in (("xpos", "xpos"), ("qM", "M"), ("qfrc_bias", "qfrc_bias"), ("qfrc_passive", "qfrc_passive"), ("qacc", "qacc")):
    d = max(np.abs(getattr(sim, name)[e].cpu().numpy().reshape(-1) - getattr(os_[e], arr).reshape(-1)).max() for e in range(n))
    print("forward", name, "max abs diff %.3g" % d)
for e in range(n):
    oc = os_[e].contacts()
    print("env", e, "ncon dev", int(sim.ncon[e]), "oracle", len(oc), "nefc dev", int(sim.nefc[e]), "oracle", os_[e].nefc,
          [(model.names["geom"][c["geom1"]], model.names["geom"][c["geom2"]]) for c in oc][:6])
rng = np.random.default_rng(0)
act = rng.uniform(-1, 1, size=(n, 8))
a_t = torch.as_tensor(act, dtype=torch.float32, device="cuda").contiguous()
for sub in range(1, 26):
    # run `sub` substeps from the same start on both sides
    pass
# substep-by-substep: device env_step with nsub=1 repeatedly (policy on first only)
sim.qpos.copy_(torch.as_tensor(q0, dtype=torch.float32)); sim.qvel.zero_(); sim.qacc_warmstart.zero_(); sim.forward(); sim.ctrl_reset()
for e in range(n):
    os_[e].reset_data(); os_[e].qpos[:] = q0[e]; os_[e].forward(); os_[e].ctrl_reset()
for sub in range(25):
    # device: one substep; action consumed when nsub loop index 0 -> emulate by calling env_step(nsub=1) with action only at sub 0
    if sub == 0:
        sim.env_step(a_t, 1)
    else:
        sim.step1(); sim._check(sim._L.b2s_env_step(sim._h, a_t.data_ptr(), 1)) if False else None
        break
torch.cuda.synchronize()
for e in range(n):
    o = os_[e]; o.step1(); o.ctrl_run(act[e]); 
    tau_o = np.array(o.ctrl_state.torques[:7]); o.step2()
    print("env", e, "substep1 qpos diff %.3g qvel diff %.3g tau diff %.3g ncon dev %d nefc dev %d oracle ncon %d nefc %d niter %d" % (
        np.abs(sim.qpos[e].cpu().numpy() - o.qpos).max(), np.abs(sim.qvel[e].cpu().numpy() - o.qvel).max(),
        np.abs(sim.ctrl_torque[e].cpu().numpy()[:7] - tau_o).max(), int(sim.ncon[e]), int(sim.nefc[e]), o.ncon, o.nefc, int(sim.solver_niter[e])))
print("warn", sim.warn.tolist())
