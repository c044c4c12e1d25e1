#!/usr/bin/env python
"""bench.py - env-steps/sec of Panda-Lift OSC_POSE on the batched engine (BASELINE.json metric).

One "step" = one control step of ENVS_PER_GPU environments = 25 x {step1, OSC_POSE+GRIP controller, step2} per env
in ONE kernel launch per GPU (robosuite/environments/base.py:467-521).  Device-timed with CUDA events around each
step on the launch stream, L2 flushed between timed iterations, max over ranks.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the same path on host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = int(os.environ.get("B2S_BENCH_ENVS", "4096"))  # BASELINE.json configs[1]; the override is for scaling experiments only
N_SUBSTEPS = 25
METRIC = "env-steps/sec (device-timed) Panda-Lift OSC_POSE @4096 envs per GPU"
WORKLOAD = "4096 Panda Lift envs, OSC_POSE, fp32, random actions, 1xB200 (BASELINE.json configs[1]); weak-scaled: 4096 envs per GPU"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            for ts, line in self.rows[-3:]:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except Exception:
                    pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
class CpuArm:
    """The oracle port (fp64 C, oracle/) of the same path on `threads` host threads; each thread owns independent
    environments (the reference runs one env per process: SURVEY.md section 2.1).  Same measurement protocol as the GPU
    arm: `preroll` untimed control steps of random actions first (steady-state contact load), then timed chunks."""

    def __init__(self, n_env, threads, preroll):
        import numpy as np

        from oracle.pyoracle import CtrlCfg, Oracle
        from robosuite_b200 import controller_config as cc
        from robosuite_b200.mjcf.compiler import load_model, pack_model
        from tests.util import lift_states

        model = load_model(os.path.join(ROOT, "robosuite_b200", "assets", "models", "Lift_Panda.npz"))
        blob = pack_model(model)
        q, _ = lift_states(model, n_env, seed=0)
        self.n_env, self.threads = n_env, threads
        self.rng = np.random.default_rng(0)
        self.sims = []
        for e in range(n_env):
            o = Oracle(blob)
            o.ctrl_setup(cc.resolve(model, cc.default_composite_config(), CtrlCfg))
            o.qpos[:] = q[e]
            o.forward()
            o.ctrl_reset()
            self.sims.append(o)
        self.preroll_s = self.run(preroll)[1] if preroll > 0 else 0.0

    def run(self, n_steps):
        """n_steps more control steps on every environment -> (env-steps/s, seconds)"""
        actions = self.rng.uniform(-1, 1, size=(n_steps, self.n_env, 7))
        sims, n_env, threads = self.sims, self.n_env, self.threads

        def work(tid):
            for e in range(tid, n_env, threads):
                for t in range(n_steps):
                    sims[e].env_step(actions[t, e], N_SUBSTEPS)

        t0 = time.perf_counter()
        ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        dt = time.perf_counter() - t0
        return n_env * n_steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_env = max(cores, 8)
    per_step = 8  # control steps per env per bench "step" (bounded sample of the workload)
    arm = CpuArm(n_env, cores, args.preroll)  # same protocol as the GPU arm: untimed pre-roll into the steady-state regime
    rates = []
    for i in range(args.warmup + args.steps):
        r, dt = arm.run(per_step)
        if i >= args.warmup:
            rates.append((r, dt))
    total_steps = sum(n_env * per_step for _ in rates)
    total_t = sum(dt for _, dt in rates)
    value = total_steps / total_t
    sample = (f"{n_env} Lift envs x {per_step} control steps per bench step after {args.preroll} untimed pre-roll steps "
              f"({arm.preroll_s:.1f}s), {cores} threads, oracle port (fp64 C) incl. OSC controller")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(len(rates), 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import robosuite_b200 as suite
    from robosuite_b200.envs.base import load_task_model
    from robosuite_b200.parallel import allgather_obs, broadcast_model

    # model constants: compiled once on rank 0, broadcast to the other ranks over NCCL
    model = load_task_model("Lift", "Panda") if rank == 0 else None
    model = broadcast_model(model, src=0, device=torch.device("cuda", local)) if world > 1 else model
    env = suite.make("Lift", robots="Panda", num_envs=ENVS_PER_GPU, device=local, seed=1000 + rank, horizon=10 ** 9,
                     has_renderer=False, has_offscreen_renderer=False, use_camera_obs=False, model=model)
    env.sim.set_mode(args.mode)
    sim = env.sim
    dev = env.device
    N, K, W = ENVS_PER_GPU, args.steps, args.warmup
    gen = torch.Generator(device=dev)
    gen.manual_seed(7 + rank)
    actions = torch.rand((W + K, N, env.action_dim), generator=gen, device=dev, dtype=env.dtype) * 2 - 1
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- pre-roll (untimed): random-action rollouts settle into their steady-state contact load only after ~50
    # control steps (cube lands, arms spread out, link-link hull tests start to fire); time THAT regime
    pre = torch.rand((args.preroll, N, env.action_dim), generator=gen, device=dev, dtype=env.dtype) * 2 - 1
    for i in range(args.preroll):
        sim.env_step(pre[i], N_SUBSTEPS)
    # ---- kernel-only timing (inputs resident in HBM)
    for i in range(W):
        sim.env_step(actions[i], N_SUBSTEPS)
    barrier()
    l0 = sim.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    clocks = ClockSampler(local) if rank == 0 else None
    t_wall0 = time.time()
    for i in range(K):
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        ev[i][0].record()
        sim.env_step(actions[W + i], N_SUBSTEPS)
        ev[i][1].record()
    barrier()
    t_wall1 = time.time()
    launches = sim.launch_count - l0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None
    value = world * N * K / (ms * 1e-3)
    warn = int(sim.warn.abs().max().item())

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    h_act = torch.empty((K, N, env.action_dim), dtype=env.dtype).pin_memory()
    h_act.copy_(actions[W:W + K].cpu())
    h_obs = torch.empty((N, env.obs_dim), dtype=env.dtype).pin_memory()
    h_rew = torch.empty((N,), dtype=env.dtype).pin_memory()
    d_act = torch.empty((N, env.action_dim), dtype=env.dtype, device=dev)
    gathered = torch.empty((world * N, env.obs_dim), dtype=env.dtype, device=dev) if world > 1 else None
    h_all = torch.empty((world * N, env.obs_dim), dtype=env.dtype).pin_memory() if (world > 1 and rank == 0) else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        d_act.copy_(h_act[i], non_blocking=True)
        obs, rew, done, _ = env.step(d_act)
        flat = env.flat_obs()
        if world > 1 and args.allgather_obs:
            flat = allgather_obs(flat, gathered)  # per-step NCCL all-gather of observations
            if rank == 0:
                h_all.copy_(flat, non_blocking=True)
        h_obs.copy_(env.flat_obs(), non_blocking=True)
        h_rew.copy_(rew, non_blocking=True)
    e1.record()
    barrier()
    ms2 = e0.elapsed_time(e1)
    t2 = torch.tensor([ms2], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * N * K / (float(t2.item()) * 1e-3)
    esz = 4 if env.dtype == torch.float32 else 8
    h2d = N * env.action_dim * esz
    d2h = N * (env.obs_dim + 1) * esz

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel.  Pipeline mode: the merged tail kernel phase_kernel<R,5> (73 % of the summed
    # kernel time, profiles/r01_pipeline_summary.md); its launch duration is measured live with CUDA events on the launching
    # stream in a short eager pass (the timed region above replays a CUDA graph, which events cannot subdivide).
    # Algorithmic bytes of ONE launch = environments per launch x per-substep state round trip (SURVEY.md section 8d:
    # B_substep = 2*4*S, here counted from the arrays the kernel really reads/writes in HBM).
    m = env.model
    per_env_in = (m.nq + 3 * m.nv + m.nu + 1 + 3 + 9 + 4 + 8 + env.action_dim) * esz
    per_env_out = (m.nq + 3 * m.nv + m.nu + 1 + 3 + 9 + 4 + env.obs_dim + 4 + 8) * esz + 4
    peak, how = _peaks()
    step_bytes = N * (per_env_in + per_env_out)
    achieved_step = step_bytes / (ms / K * 1e-3) / 1e9
    kernel, launch_us, envs_per_launch = "step_kernel", ms / K * 1e3, N
    alg_bytes = step_bytes
    if args.mode == 1:
        sim.timeline(1)
        tl = []
        for i in range(2):
            sim.env_step(actions[W + i], N_SUBSTEPS)
            tl.append(sim.timeline(-1))
        sim.timeline(0)
        mean_us, cnt = tl[-1]
        groups = max(1, cnt[5] // N_SUBSTEPS)
        kernel, launch_us, envs_per_launch = "phase_kernel<float,5> (rows+controller+solve+integrate)", mean_us[5], N // groups
        sub_in = (m.nq + 2 * m.nv + m.nu + 1 + 3 + 9 + 4) * esz   # qpos qvel qacc_ws ctrl time + controller state
        sub_out = (m.nq + 3 * m.nv + m.nu + 1) * esz                # qpos qvel qacc qacc_ws ctrl time
        alg_bytes = envs_per_launch * (sub_in + sub_out)
    achieved = alg_bytes / (launch_us * 1e-6) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    # ---- CPU baseline on a bounded sample (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_env, n_steps = max(cores, 8) * 2, 50
        arm = CpuArm(n_env, cores, args.preroll)
        r, dtc = arm.run(n_steps)
        cpu = {"value": r, "unit": "env-steps/s", "cores": cores, "kind": "port",
               "sample": f"{n_env} Lift envs x {n_steps} control steps ({dtc:.1f}s) after {args.preroll} untimed pre-roll steps "
                         f"({arm.preroll_s:.1f}s), oracle port (fp64 C) incl. OSC, {cores} threads"}
    out = {
        "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if esz == 4 else "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": N, "substeps_per_step": N_SUBSTEPS, "controller": "OSC_POSE+GRIP",
                   "l2": "flushed (256 MiB memset) between timed iterations", "solver_warn_flags": warn,
                   "preroll_steps": args.preroll, "kernel_mode": "pipeline" if args.mode else "fused",
                   "multi_gpu": "env shards independent; NCCL: model broadcast at start" + (", obs all-gather per step (e2e loop)" if args.allgather_obs else "")},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": how, "kernel": kernel, "launch_us": launch_us,
                     "envs_per_launch": envs_per_launch, "alg_bytes_per_launch": alg_bytes,
                     "whole_step": {"achieved": achieved_step, "frac": achieved_step / peak, "alg_bytes": step_bytes},
                     "note": "per-environment state stays in shared memory / L2 between phases: algorithmic HBM traffic is tiny, "
                             "the kernels are latency / instruction-fetch bound (DESIGN.md section 5)"},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clk,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--preroll", type=int, default=100, help="untimed control steps before the timed region")
    ap.add_argument("--mode", type=int, default=1, help="0 fused kernel, 1 phase-kernel pipeline (default)")
    ap.add_argument("--allgather-obs", type=int, default=1, help="N>1: all-gather observations over NCCL every e2e step")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
