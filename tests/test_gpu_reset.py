"""Masked episode reset on the device (b2s_reset_envs): the environments outside the mask must not notice it
(bit-identical trajectories), the masked ones must behave exactly like a fresh episode from the same sampled state, and the
auto-resetting wrapper must never synchronise with the device to find out which episodes ended.
Reference semantics: MujocoEnv.reset = _reset_internal -> sim.forward -> controller reset -> observables reset (environments/base.py:277-347)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


def _make(task="Lift", n=64, seed=5, **kw):
    import robosuite_b200 as suite

    return suite.make(task, robots="Panda", num_envs=n, seed=seed, **kw)


def _snap(env):
    s = env.sim
    return {k: getattr(s, k).clone() for k in ("qpos", "qvel", "qacc_warmstart", "ctrl", "time", "obs", "ctrl_goal_pos", "ctrl_goal_ori")}


@pytest.mark.parametrize("task", ["Lift", "Stack"])
def test_masked_reset_is_invisible_to_the_other_environments(task):
    n = 64
    a, b = _make(task, n), _make(task, n)  # same seed: same initial states, same sampler stream
    gen = torch.Generator(device=a.device)
    gen.manual_seed(0)
    acts = torch.rand((12, n, a.action_dim), generator=gen, device=a.device, dtype=a.dtype) * 2 - 1
    a.reset()
    b.reset()
    for t in range(5):
        a.step(acts[t])
        b.step(acts[t])
    mask = torch.zeros(n, dtype=torch.bool, device=a.device)
    mask[3::7] = True
    before = _snap(a)
    a.reset(mask=mask)
    after = _snap(a)
    for k in before:  # untouched rows: bit-identical, including the observation cache and the controller goals
        assert torch.equal(before[k][~mask], after[k][~mask]), k
    assert torch.equal(after["qvel"][mask], torch.zeros_like(after["qvel"][mask]))
    assert torch.equal(after["qpos"][mask], a._reset_qpos[mask])
    assert not torch.equal(before["qpos"][mask], after["qpos"][mask])
    assert int(a.timestep[mask].max()) == 0 and int(a.timestep[~mask].min()) == 5
    # the other environments keep following the never-reset twin bit for bit
    for t in range(5, 10):
        a.step(acts[t])
        b.step(acts[t])
    assert torch.equal(a.sim.qpos[~mask], b.sim.qpos[~mask])
    assert torch.equal(a.sim.obs[~mask], b.sim.obs[~mask])
    # the masked ones replay exactly like a fresh episode started from the same sampled state with the same actions
    c = _make(task, n)
    c.reset_to(a._reset_qpos)
    for t in range(5, 10):
        c.step(acts[t])
    assert torch.equal(a.sim.qpos[mask], c.sim.qpos[mask])
    assert torch.equal(a.sim.obs[mask], c.sim.obs[mask])
    assert int(a.sim.warn.abs().max()) == 0


def test_wrapper_autoreset_uses_no_device_readback(monkeypatch):
    from robosuite_b200.wrappers import BatchedGymWrapper

    n, H = 32, 6
    env = _make("Lift", n, horizon=H)
    w = BatchedGymWrapper(env)
    w.reset()
    env.set_episode_steps(np.arange(n) % H)  # a few episodes end on every step
    gen = torch.Generator(device=env.device)
    gen.manual_seed(1)
    acts = torch.rand((3 * H, n, env.action_dim), generator=gen, device=env.device, dtype=env.dtype) * 2 - 1
    torch.cuda.synchronize()
    ends = 0
    with torch.cuda.stream(torch.cuda.current_stream()):
        torch.cuda.set_sync_debug_mode("error")  # any implicit device->host synchronisation raises
        try:
            outs = []
            for t in range(3 * H):
                obs, rew, term, trunc, info = w.step(acts[t])
                outs.append((term, "final_observation" in info))
        finally:
            torch.cuda.set_sync_debug_mode("default")
    steps = np.arange(n) % H
    for t, (term, had_final) in enumerate(outs):
        steps = steps + 1
        expect = steps >= H
        assert np.array_equal(term.cpu().numpy(), expect), t
        assert had_final == bool(expect.any())
        ends += int(expect.sum())
        steps[expect] = 0
    assert ends >= 3 * n - n
    assert np.array_equal(env.timestep.cpu().numpy(), steps)
    assert int(env.sim.warn.abs().max()) == 0


def test_reset_by_device_mask_falls_back_to_always_enqueued_reset():
    from robosuite_b200.wrappers import BatchedGymWrapper

    n, H = 16, 4
    env = _make("Lift", n, horizon=H)
    w = BatchedGymWrapper(env)
    w.reset()
    for _ in range(2):
        w.step(torch.zeros((n, env.action_dim), device=env.device, dtype=env.dtype))
    m = torch.zeros(n, dtype=torch.bool, device=env.device)
    m[:5] = True
    env.reset(mask=m)  # host mirror of the episode clocks is gone now
    assert env.host_done() is None
    seen = torch.zeros(n, dtype=torch.long, device=env.device)
    for t in range(8):
        obs, rew, term, trunc, info = w.step(torch.zeros((n, env.action_dim), device=env.device, dtype=env.dtype))
        seen += term.long()
    assert int(env.timestep.max()) < H
    assert seen.cpu().tolist() == [2] * 5 + [2] * (n - 5)
