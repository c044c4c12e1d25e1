import torch, sys, time
sys.path.insert(0, ".")
import robosuite_b200 as suite
env = suite.make("Lift", robots="Panda", num_envs=4096, seed=1, horizon=10**9)
sim = env.sim
import os
MODE = int(os.environ.get("B2S_MODE", "0"))
sim.set_mode(MODE)
g = torch.Generator(device="cuda"); g.manual_seed(0)
ts = []
for t in range(101):
    a = torch.rand((4096, 7), generator=g, device="cuda") * 2 - 1
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    if t % 25 == 0 and MODE == 0: sim.set_export(True); sim.set_profile(True)
    e0.record(); sim.env_step(a, 25); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
    if t % 25 == 0 and MODE == 1: print(t, 'ms %.2f' % ts[-1])
    if t % 25 == 0 and MODE == 0:
        nc = sim.ncon.float(); ne = sim.nefc.float(); ni = sim.solver_niter.float()
        print(t, "ms %.2f" % ts[-1], "ncon mean %.2f max %d" % (nc.mean().item(), nc.max().item()), "nefc mean %.1f max %d" % (ne.mean().item(), ne.max().item()), "niter mean %.2f max %d" % (ni.mean().item(), ni.max().item()), "warn", int(sim.warn.abs().max()))
        pr = sim.prof.float(); db = sim.dbg.float()
        names = ["kin", "vel+crb", "collide", "mkcon", "ctrl", "accel", "solve", "euler", "col:cull", "col:analytic", "col:convex", "barrier"]
        tot = pr.sum(1)
        print("   cycles/substep mean: " + " ".join(f"{n}={pr[:, i].mean().item() / 25:.0f}" for i, n in enumerate(names) if n), "| total mean %.0f max %.0f" % (tot.mean().item() / 25, tot.max().item() / 25))
        print("   per-env collide cycles: p50 %.0f p90 %.0f p99 %.0f max %.0f | candidates/substep analytic %.2f convex %.2f (max %.1f) epa-hits %.3f" % (
            *(torch.quantile(pr[:, 2], torch.tensor([0.5, 0.9, 0.99, 1.0], device="cuda")) / 25).tolist(), db[:, 0].mean().item() / 25, db[:, 1].mean().item() / 25, db[:, 1].max().item() / 25, db[:, 2].mean().item() / 25))
        sim.set_export(False); sim.set_profile(False)
print("mean ms first 20: %.2f, steps 80-100: %.2f" % (sum(ts[:20])/20, sum(ts[80:100])/20))
