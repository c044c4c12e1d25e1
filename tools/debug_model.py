import sys, numpy as np, torch
sys.path.insert(0, ".")
from robosuite_b200.engine import BatchedSim
from robosuite_b200.mjcf.compiler import pack_model
from oracle.pyoracle import Oracle
from tests.util import load
name = sys.argv[1]
model = load(name); n = 1
q = np.tile(model.qpos0, (n, 1))
k = 0
for j in range(model.njnt):
    if model.jnt_type[j] == 0:
        q[:, model.jnt_qposadr[j] + 1] += 0.12 * k - 0.12
        if name.startswith("PickPlace"):
            q[:, model.jnt_qposadr[j]] += 0.25 * k - 0.4; q[:, model.jnt_qposadr[j] + 2] += 0.04
        k += 1
        q[:, model.jnt_qposadr[j] + 2] += 0.02
for prec in ("f64", "f32"):
    sim = BatchedSim(model, n, precision=prec, maxcon=48, maxefc=160)
    sim.qpos.copy_(torch.as_tensor(q, dtype=sim.dtype)); sim.forward(); torch.cuda.synchronize()
    o = Oracle(pack_model(model)); o.qpos[:] = q[0]; o.forward()
    print(prec, "warn", sim.warn.tolist(), "ncon", int(sim.ncon[0]), o.ncon, "nefc", int(sim.nefc[0]), o.nefc, "niter", int(sim.solver_niter[0]), o.geti("solver_niter"))
    for nm, arr in (("xpos", o.xpos), ("qM", o.M), ("qfrc_bias", o.qfrc_bias), ("qfrc_passive", o.qfrc_passive), ("qacc_smooth", o.qacc_smooth), ("qfrc_constraint", o.qfrc_constraint), ("qacc", o.qacc)):
        d = getattr(sim, nm)[0].cpu().numpy().reshape(-1).astype(np.float64) - arr.reshape(-1)
        i = np.abs(d).argmax()
        print("   %-16s max abs diff %.3g at %d (oracle %.4g) scale %.3g" % (nm, np.abs(d).max(), i, arr.reshape(-1)[i], np.abs(arr).max()))
    ne = o.nefc
    for nm in ("aref", "D", "force"):
        a = getattr(sim, "efc_" + nm)[0].cpu().numpy()[:ne].astype(np.float64); b = o.efc(nm)
        i = np.abs(a - b).argmax(); print("   efc_%-12s max abs diff %.3g at row %d (oracle %.4g)" % (nm, np.abs(a - b).max(), i, b[i]))
    J = sim.efc_J[0].cpu().numpy()[:ne].astype(np.float64); print("   efc_J max abs diff %.3g" % np.abs(J - o.efc("J")).max())
    sim.close()
sim = BatchedSim(model, n, precision="f64", maxcon=48, maxefc=160)
sim.qpos.copy_(torch.as_tensor(q, dtype=sim.dtype)); sim.forward(); torch.cuda.synchronize()
o = Oracle(pack_model(model)); o.qpos[:] = q[0]; o.forward()
gn = model.names["geom"]
cg = sim.contact_geom[0].cpu().numpy(); cd = sim.contact_dist[0].cpu().numpy(); cp = sim.contact_pos[0].cpu().numpy(); cf = sim.contact_frame[0].cpu().numpy(); cdim = sim.contact_dim[0].cpu().numpy()
for i, c in enumerate(o.contacts()):
    print(i, gn[c["geom1"]], gn[c["geom2"]], "| oracle dist %.6f pos %s n %s dim %d | dev (%s,%s) dist %.6f pos %s n %s dim %d" % (
        c["dist"], np.round(c["pos"], 4), np.round(c["frame"][0], 3), c["dim"], gn[cg[i, 0]][-12:], gn[cg[i, 1]][-12:], cd[i], np.round(cp[i], 4), np.round(cf[i][:3], 3), cdim[i]))
