"""Localise a small-tier discrepancy: scripted contact-rich rollout with and without the small tail tier, control step by control
step; reports the first step / environment where the states differ, for several tier settings."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robosuite_b200 import controller_config as cc
from robosuite_b200.engine import BatchedSim, CtrlCfg
from tests.util import lift_states, load

model = load("Lift_Panda")
n, steps = 16, 24
q, v = lift_states(model, n, seed=21)
rng = np.random.default_rng(3)
actions = rng.uniform(-1, 1, size=(steps, n, 7))
actions[:, :, 6] = 1.0
actions[8:, : n // 2, :3] = [0.0, 0.0, -1.0]
os.environ["B2S_NO_GJK_CACHE"] = "1"


def mk(tier):
    sim = BatchedSim(model, n, precision="f32", tier_small=tier)
    sim.ctrl_config(cc.resolve(model, cc.default_composite_config(), CtrlCfg))
    sim.set_export(False); sim.set_mode(1)
    sim.qpos.copy_(torch.as_tensor(q, dtype=torch.float32)); sim.forward(); sim.ctrl_reset()
    return sim


for tier in [tuple(int(x) for x in sys.argv[1].split(','))]:
    a, b = mk(None), mk(tier)
    first = None
    for t in range(steps):
        act = torch.as_tensor(actions[t], dtype=torch.float32, device="cuda").contiguous()
        # keep B on A's trajectory: copy A's full state into B before the step, then compare after one control step
        for name in ("qpos", "qvel", "qacc", "qacc_warmstart", "ctrl", "time", "ctrl_goal_pos", "ctrl_goal_ori", "ctrl_initial_joint", "ctrl_grip_state"):
            getattr(b, name).copy_(getattr(a, name))
        a.env_step(act, 25); b.env_step(act, 25)
        torch.cuda.synchronize()
        d = (a.qpos - b.qpos).abs().max(dim=1).values
        bad = torch.nonzero(~(d == 0)).flatten().tolist()
        if bad and first is None:
            first = (t, bad, d[bad].tolist())
            # contact / row counts of the differing environments at the START state of the next substep (fused forward + export)
            a.forward(); torch.cuda.synchronize()
            print("   tier", tier, "first difference at control step", t, "envs", bad, "|dq|", [float("%.3g" % x) for x in d[bad].tolist()],
                  "ncon", a.ncon[bad].tolist(), "nefc", a.nefc[bad].tolist(), "warn b", b.warn[bad].tolist(), flush=True)
            a.set_mode(1)
            break
    print("tier", tier, "->", "IDENTICAL over %d control steps" % steps if first is None else "differs (see above)", "| warn a %d b %d" % (int(a.warn.abs().max()), int(b.warn.abs().max())), flush=True)
    a.close(); b.close()
