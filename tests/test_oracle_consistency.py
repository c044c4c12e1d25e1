"""Extra known-answer / self-consistency checks of the CPU oracle's physics (the part of the path no reference vector pins).

1. Bias forces (recursive Newton-Euler) against the Lagrangian form built from the oracle's OTHER dynamics path: the mass
   matrix (composite rigid body) differentiated numerically, plus the gradient of the potential energy.
2. GJK/EPA penetration depth against a brute-force minimum of the support function over 200k sampled directions."""
import numpy as np

from tests.util import load


def _oracle(model):
    from oracle.pyoracle import Oracle
    from robosuite_b200.mjcf.compiler import pack_model

    return Oracle(pack_model(model))


def test_rne_bias_matches_lagrangian_form_from_crb_mass_matrix():
    """c(q, qd) + g(q) = sum_jk (dM_ij/dq_k - 1/2 dM_jk/dq_i) qd_j qd_k + dV/dq_i  for the arm + gripper dofs (hinges, slides)"""
    model = load("Lift_Panda")
    o = _oracle(model)
    rng = np.random.default_rng(3)
    # scalar joints of the robot only (the free cube has a quaternion: its qpos is not a generalised coordinate vector)
    sj = [j for j in range(model.njnt) if model.jnt_type[j] in (2, 3) and model.names["joint"][j].startswith(("robot0", "gripper0"))]
    qa = [int(model.jnt_qposadr[j]) for j in sj]
    da = [int(model.jnt_dofadr[j]) for j in sj]
    q0 = model.qpos0.copy()
    q0[qa[:7]] = [0.1, 0.3, -0.2, -2.0, 0.3, 2.2, 0.6]
    q0[qa[7:]] = [0.02, -0.02]
    qd = np.zeros(model.nv)
    qd[da] = rng.normal(0, 0.8, len(da))
    mass = np.asarray(model.body_mass, dtype=np.float64)
    g = 9.81

    def eval_at(q):
        o.reset_data(); o.qpos[:] = q; o.qvel[:] = 0
        o.kinematics(); o.crb()
        M = o.M.copy()
        V = float((mass * g * o.xipos[:, 2]).sum())
        return M, V

    n = len(da)
    h = 1e-6
    dM = np.zeros((n, model.nv, model.nv)); dV = np.zeros(n)
    for k in range(n):
        qp, qm = q0.copy(), q0.copy()
        qp[qa[k]] += h; qm[qa[k]] -= h
        Mp, Vp = eval_at(qp); Mm, Vm = eval_at(qm)
        dM[k] = (Mp - Mm) / (2 * h); dV[k] = (Vp - Vm) / (2 * h)
    expect = np.zeros(n)
    for i in range(n):
        acc = 0.0
        for a, j in enumerate(da):
            for b, k in enumerate(da):
                acc += (dM[b][da[i], j] - 0.5 * dM[i][j, k]) * qd[j] * qd[k]
        expect[i] = acc + dV[i]
    o.reset_data(); o.qpos[:] = q0; o.qvel[:] = qd
    # fluid forces (density / viscosity of the robosuite world) are passive forces, not part of qfrc_bias
    o.forward()
    got = o.qfrc_bias[da]
    assert np.abs(got - expect).max() < 2e-5 * max(1.0, np.abs(expect).max()), (got, expect)


def test_epa_depth_matches_brute_force_support_minimum():
    """deep mesh-mesh / box-mesh penetrations (PickPlace at qpos0: link5 in the hand, milk carton in the pedestal):
    depth = min over unit directions of the support function of the Minkowski difference"""
    model = load("PickPlace_Panda")
    q = model.qpos0.copy()
    k = 0
    for j in range(model.njnt):
        if model.jnt_type[j] == 0:
            a = model.jnt_qposadr[j]
            q[a] += 0.25 * k - 0.4; q[a + 1] += 0.12 * k - 0.12; q[a + 2] += 0.06
            k += 1
    o = _oracle(model)
    o.qpos[:] = q
    o.forward()
    rng = np.random.default_rng(0)
    D = rng.normal(size=(200000, 3)); D /= np.linalg.norm(D, axis=1)[:, None]

    def support(gid, D):
        t = int(model.geom_type[gid]); pos = o.geom_xpos[gid]; R = o.geom_xmat[gid].reshape(3, 3); sz = model.geom_size[gid]
        Dl = D @ R
        if t == 7:
            m = int(model.geom_dataid[gid]); a = int(model.mesh_vertadr[m]); n = int(model.mesh_vertnum[m])
            h = (Dl @ model.mesh_vert[a:a + n].T).max(axis=1)
        elif t == 6:
            h = (np.abs(Dl) * sz[:3]).sum(axis=1)
        else:
            return None
        return h + D @ pos

    checked = 0
    for c in o.contacts():
        g1, g2 = c["geom1"], c["geom2"]
        if 7 not in (int(model.geom_type[g1]), int(model.geom_type[g2])) or c["dist"] > -3e-3:
            continue
        s1, s2 = support(g1, D), support(g2, -D)
        if s1 is None or s2 is None:
            continue
        brute = float((s1 + s2).min())  # >= true depth, -> true depth as the sampling gets denser
        assert -c["dist"] <= brute + 1e-9, (model.names["geom"][g1], model.names["geom"][g2], c["dist"], brute)
        assert brute + c["dist"] < 0.03 * brute + 2e-4, (model.names["geom"][g1], model.names["geom"][g2], c["dist"], brute)
        checked += 1
    assert checked >= 2


def test_constraint_solution_satisfies_kkt_conditions():
    """At the solver's answer: M qacc = qfrc_smooth + J^T f (stationarity), unilateral rows f >= 0, elliptic cone rows
    inside the friction cone |f_t|_mu <= mu f_n, and f = 0 on rows whose constraint acceleration is already above its
    reference (complementarity) - for gripper-on-cube and arm-on-table contact states of Lift"""
    from tests.util import lift_states

    model = load("Lift_Panda")
    o = _oracle(model)
    q, _ = lift_states(model, 4, seed=5)
    rng = np.random.default_rng(1)
    n_cone = 0
    for e in range(4):
        o.reset_data(); o.qpos[:] = q[e]
        o.qvel[:] = rng.normal(0, 0.3, model.nv)
        o.ctrl[:] = 0
        for _ in range(40):  # let contacts develop (cube on table, closed fingers)
            o.step()
        o.forward()
        ne = o.nefc
        assert ne > 0
        J = o.efc("J")[:ne]; f = o.efc("force")[:ne]
        resid = o.M @ o.qacc - o.qfrc_smooth - J.T @ f
        scale = max(1.0, np.abs(o.qfrc_smooth).max())
        assert np.abs(resid).max() < 1e-6 * scale, np.abs(resid).max()
        assert np.allclose(J.T @ f, o.qfrc_constraint, atol=1e-9 * scale)
        for c in o.contacts():
            a = c.get("efc_address", -1)
            if a is None or a < 0:
                continue
            fn = f[a]
            assert fn >= -1e-10
            if c["dim"] >= 3:
                mu = c["friction"][0]
                ft = np.hypot(f[a + 1], f[a + 2])
                assert ft <= mu * fn * (1 + 1e-6) + 1e-9, (ft, mu * fn)
                n_cone += fn > 1e-6
    assert n_cone >= 4  # the states really exercise loaded frictional contacts
