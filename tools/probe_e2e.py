"""Where does the end-to-end step spend its time beyond the simulation kernels?  Ablation on config 2 (4096 Lift, horizon 500, staggered
episode phases): CUDA-event ms per control step of (0) sim.env_step only, (1) env.step without resets, (2) BatchedGymWrapper.step with
in-step resets, (3) = (2) + pinned-host action upload and obs / reward download (bench.py's e2e loop), plus the reset path alone."""
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robosuite_b200 as suite  # noqa: E402
from robosuite_b200.wrappers import BatchedGymWrapper  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
env = suite.make("Lift", robots="Panda", num_envs=n, seed=1000, horizon=500, ignore_done=True)
env.sim.set_mode(int(os.environ.get("B2S_BENCH_MODE", "1")))
dev = env.device
gen = torch.Generator(device=dev); gen.manual_seed(7)
A = torch.rand((100 + 5 * K, n, env.action_dim), generator=gen, device=dev, dtype=env.dtype) * 2 - 1
for i in range(60):
    env.sim.env_step(A[i], 25)
torch.cuda.synchronize()


def timed(fn, k=K):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for i in range(k):
        fn(i)
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, 1e3 * t_host / k


print("variant: device ms / step, host enqueue ms / step")
print("0 sim.env_step               %.2f  %.2f" % timed(lambda i: env.sim.env_step(A[60 + i], 25)))
print("1 env.step (no resets)       %.2f  %.2f" % timed(lambda i: env.step(A[60 + K + i])))
w = BatchedGymWrapper(env)
env.ignore_done = False
env.set_episode_steps(torch.randint(0, env.horizon, (n,), generator=gen, device=dev))
print("2 wrapper.step (resets)      %.2f  %.2f" % timed(lambda i: w.step(A[60 + 2 * K + i])))
h_act = torch.empty((K, n, env.action_dim), dtype=env.dtype).pin_memory(); h_act.copy_(A[60 + 3 * K:60 + 4 * K].cpu())
d_act = torch.empty((n, env.action_dim), dtype=env.dtype, device=dev)
h_obs = torch.empty((n, w.obs_dim), dtype=env.dtype).pin_memory(); h_rew = torch.empty((n,), dtype=env.dtype).pin_memory()


def full(i):
    d_act.copy_(h_act[i], non_blocking=True)
    obs, rew, term, trunc, info = w.step(d_act)
    h_obs.copy_(obs, non_blocking=True); h_rew.copy_(rew, non_blocking=True)


print("3 + host copies              %.2f  %.2f" % timed(full))
mask = torch.zeros(n, dtype=torch.bool, device=dev); mask[::512] = True
hm = mask.cpu().numpy()
print("reset of 8 environments only %.2f  %.2f" % timed(lambda i: env.reset(mask=mask, host_mask=hm)))
print("_sample_reset_state only     %.2f  %.2f" % timed(lambda i: env._sample_reset_state(n)))
q = env._sample_reset_state(n).to(env.dtype).contiguous(); m8 = mask.to(torch.uint8)
print("sim.reset_envs only          %.2f  %.2f" % timed(lambda i: env.sim.reset_envs(m8, q)))
print("reward + obs dict only       %.2f  %.2f" % timed(lambda i: (env.reward(None), env._get_observations())))
