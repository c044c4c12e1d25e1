import sys, torch, os
sys.path.insert(0, ".")
import robosuite_b200 as suite
mode = os.environ.get("KM", "pipeline")
env = suite.make("Lift", robots="Panda", num_envs=4096, seed=1, horizon=10**9, kernel_mode=mode)
sim = env.sim
g = torch.Generator(device="cuda"); g.manual_seed(0)
for t in range(130):
    sim.env_step(torch.rand((4096, 7), generator=g, device="cuda") * 2 - 1, 25)
    if t % 10 == 9:
        w = sim.warn
        bad = torch.nonzero(w).flatten()
        nan = torch.isnan(sim.qpos).any(1).sum().item()
        print(t, "envs with warn", bad.numel(), "bits", sorted(set(w[bad].tolist()))[:8], "nan envs", nan, "max|qvel| %.1f" % sim.qvel.abs().nan_to_num(0).max().item())
        if bad.numel():
            e = int(bad[0]); print("   env", e, "qpos", [round(x, 3) for x in sim.qpos[e].tolist()])
