"""Trace EPA iterations of one model on both sides (device lib built with -DB2S_EPA_TRACE, oracle with B2S_EPA_TRACE=1)."""
import sys, os, numpy as np, torch
sys.path.insert(0, ".")
os.environ["B2S_EPA_TRACE"] = "1"
from robosuite_b200.engine import BatchedSim
from robosuite_b200.mjcf.compiler import pack_model
from oracle.pyoracle import Oracle
from tests.util import load
name = sys.argv[1]
model = load(name)
q = model.qpos0.copy()[None]
k = 0
for j in range(model.njnt):
    if model.jnt_type[j] == 0:
        q[:, model.jnt_qposadr[j] + 1] += 0.12 * k - 0.12
        if name.startswith("PickPlace"):
            q[:, model.jnt_qposadr[j]] += 0.25 * k - 0.4; q[:, model.jnt_qposadr[j] + 2] += 0.04
        k += 1
        q[:, model.jnt_qposadr[j] + 2] += 0.02
sim = BatchedSim(model, 1, precision="f64", maxcon=48, maxefc=160)
sim.qpos.copy_(torch.as_tensor(q, dtype=sim.dtype)); sim.forward(); torch.cuda.synchronize()
sys.stdout.flush()
o = Oracle(pack_model(model)); o.qpos[:] = q[0]; o.forward()
