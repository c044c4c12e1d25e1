"""CPU stand-in for robosuite_b200.engine.BatchedSim, for tests only: the same array / method surface, every environment
stepped by the fp64 oracle, the observation / task tables evaluated in numpy.  It lets the `-m "not gpu"` suite run the HOST side
of the environment layer (observation tables, reset logic, reward / success code of every task class) against the golden vectors
the reference stack produced - the device evaluation of the same tables is what the `-m gpu` tests cover."""
import numpy as np
import torch

from oracle.pyoracle import CtrlCfg as OCfg
from oracle.pyoracle import Oracle
from robosuite_b200.mjcf.compiler import pack_model

(OB_QPOS, OB_COS_QPOS, OB_SIN_QPOS, OB_QVEL, OB_QACC, OB_SITE_POS, OB_BODY_POS, OB_BODY_QUAT_XYZW, OB_SITE_QUAT_XYZW,
 OB_BODY_MINUS_SITE, OB_SITE_MINUS_SITE, OB_BODY_QUAT_REL_SITE_XYZW, OB_ZERO, OB_BODY_MINUS_BODY, OB_REL_POS_LAG,
 OB_REL_QUAT_LAG) = range(16)


def _q2m(q):  # wxyz
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def _m2q_xyzw_wpos(M):
    from scipy.spatial.transform import Rotation

    q = Rotation.from_matrix(M).as_quat()
    return -q if q[3] < 0 else q


class OracleSim:
    def __init__(self, model, n_env, device=0, precision="f64", maxcon=None, maxefc=None, tier_small=None):
        self.model, self.n_env = model, int(n_env)
        self.torch_device, self.dtype = torch.device("cpu"), torch.float64
        blob = pack_model(model)
        self.o = [Oracle(blob) for _ in range(self.n_env)]
        z = lambda *s: torch.zeros(s, dtype=torch.float64)
        self.qpos, self.qvel, self.qacc = z(n_env, model.nq), z(n_env, model.nv), z(n_env, model.nv)
        self.qacc_warmstart, self.ctrl, self.time = z(n_env, model.nv), z(n_env, model.nu), z(n_env)
        self.qpos[:] = torch.as_tensor(np.asarray(model.qpos0))
        self.warn = torch.zeros(n_env, dtype=torch.int32)
        self.obs_fresh = torch.ones(n_env, dtype=torch.int32)
        self.task_out = z(n_env, 8)
        self.obs = self.task_vec = None
        self._cfg = self._obs_tab = self._task_tab = None
        self._task = dict(body=0, site=0, left=[], right=[], obj=[], body2=-1, obj2=[], objs=[])

    # ---- configuration (same calls as BatchedSim)
    def ctrl_config(self, cfg):
        oc = OCfg()
        for name, _ in OCfg._fields_:  # engine.CtrlCfg and the oracle's struct share their layout
            setattr(oc, name, getattr(cfg, name))
        self._cfg = oc
        for o in self.o:
            o.ctrl_setup(oc)

    def obs_config(self, op, a, b):
        self._obs_tab = (np.asarray(op), np.asarray(a), np.asarray(b))
        self.obs = torch.zeros((self.n_env, len(op)), dtype=torch.float64)

    def task_table(self, rows):
        arr = np.asarray(rows, dtype=np.int64).reshape(-1, 3)
        self._task_tab = (arr[:, 0], arr[:, 1], arr[:, 2])
        self.task_vec = torch.zeros((self.n_env, len(arr)), dtype=torch.float64)

    def task_config(self, body, site, left, right, obj):
        self._task.update(body=int(body), site=int(site), left=list(left), right=list(right), obj=list(obj))

    def task_config2(self, body2, obj2):
        self._task.update(body2=int(body2), obj2=list(obj2))

    def task_objects(self, geom_lists):
        self._task["objs"] = [list(g) for g in geom_lists]

    def set_export(self, flag): pass
    def set_mode(self, mode): pass
    def close(self): pass

    # ---- state exchange with the oracles
    def body_pose_override(self, body_id):
        """same surface as BatchedSim.body_pose_override; only bodies attached to the world directly reach the oracle (its
        kinematics composes welded children from their parents, so their overrides are redundant there)"""
        m = self.model
        b = int(body_id)
        chain, k = [], b
        while k > 0:
            chain.append(k); k = int(m.body_parentid[k])
        pos, quat = np.zeros(3), np.array([1.0, 0, 0, 0])
        for k in reversed(chain):
            pos = pos + _q2m(quat) @ np.asarray(m.body_pos[k], dtype=float)
            a, c = quat, np.asarray(m.body_quat[k], dtype=float)
            quat = np.array([a[0] * c[0] - a[1:] @ c[1:], *(a[0] * c[1:] + c[0] * a[1:] + np.cross(a[1:], c[1:]))])
        P = torch.as_tensor(np.tile(pos, (self.n_env, 1))); Q = torch.as_tensor(np.tile(quat, (self.n_env, 1)))
        if not hasattr(self, "_ov"):
            self._ov = {}
        self._ov[b] = (P, Q, int(m.body_parentid[b]) == 0)
        return P, Q

    def _push(self, e):
        o = self.o[e]
        for b, (P, Q, direct) in getattr(self, "_ov", {}).items():
            if direct:
                o.set_body_pose(b, P[e].numpy(), Q[e].numpy())
        o.qpos[:] = self.qpos[e].numpy(); o.qvel[:] = self.qvel[e].numpy(); o.ctrl[:] = self.ctrl[e].numpy()
        o.qacc_warmstart[:] = self.qacc_warmstart[e].numpy(); o.time = float(self.time[e])

    def _pull(self, e):
        o = self.o[e]
        self.qpos[e] = torch.as_tensor(o.qpos.copy()); self.qvel[e] = torch.as_tensor(o.qvel.copy())
        self.qacc[e] = torch.as_tensor(o.qacc.copy()); self.qacc_warmstart[e] = torch.as_tensor(o.qacc_warmstart.copy())
        self.ctrl[e] = torch.as_tensor(o.ctrl.copy()); self.time[e] = o.time

    def _value(self, o, op, a, b, prev, fresh):
        if op == OB_QPOS: return o.qpos[a]
        if op == OB_COS_QPOS: return np.cos(o.qpos[a])
        if op == OB_SIN_QPOS: return np.sin(o.qpos[a])
        if op == OB_QVEL: return o.qvel[a]
        if op == OB_QACC: return o.qacc[a]
        if op == OB_SITE_POS: return o.site_xpos[a][b]
        if op == OB_BODY_POS: return o.xpos[a][b]
        if op == OB_BODY_QUAT_XYZW: return o.xquat[a][(b + 1) & 3]
        if op == OB_SITE_QUAT_XYZW: return _m2q_xyzw_wpos(o.site_xmat[a].reshape(3, 3))[b]
        if op == OB_BODY_MINUS_SITE: return o.xpos[a >> 8][b] - o.site_xpos[a & 255][b]
        if op == OB_SITE_MINUS_SITE: return o.site_xpos[a >> 8][b] - o.site_xpos[a & 255][b]
        if op == OB_BODY_MINUS_BODY: return o.xpos[a >> 8][b] - o.xpos[a & 255][b]
        if op in (OB_REL_POS_LAG, OB_REL_QUAT_LAG):
            if fresh or prev is None:
                return 0.0
            ps, qs, comp, site, body = a & 4095, a >> 12, b & 255, (b >> 8) & 255, (b >> 16) & 255
            Re = _q2m(o.xquat[body])
            if op == OB_REL_POS_LAG:
                return (Re.T @ (prev[ps:ps + 3] - o.site_xpos[site]))[comp]
            return _m2q_xyzw_wpos(Re.T @ _q2m(prev[qs:qs + 4][[3, 0, 1, 2]]))[comp]
        return 0.0

    def _sample_obs(self, e):
        if self._obs_tab is None:
            return
        o, prev, fresh = self.o[e], self.obs[e].numpy().copy(), bool(self.obs_fresh[e])
        op, a, b = self._obs_tab
        self.obs[e] = torch.as_tensor([self._value(o, int(op[k]), int(a[k]), int(b[k]), prev, fresh) for k in range(len(op))])
        self.obs_fresh[e] = 0

    def _sample_task(self, e):
        o, t = self.o[e], self._task
        cons = o.contacts()

        def touching(A, B):
            return any((c["geom1"] in A and c["geom2"] in B) or (c["geom2"] in A and c["geom1"] in B) for c in cons)

        bp, sp = o.xpos[t["body"]], o.site_xpos[t["site"]]
        out = np.zeros(8)
        out[0], out[1] = bp[2], np.linalg.norm(bp - sp)
        out[2] = float(touching(t["left"], t["obj"]) and touching(t["right"], t["obj"]))
        if t["body2"] >= 0:
            out[3] = np.linalg.norm(bp[:2] - o.xpos[t["body2"]][:2]); out[4] = float(touching(t["obj"], t["obj2"]))
        out[5] = sum((1 << i) for i, g in enumerate(t["objs"]) if touching(t["left"], g) and touching(t["right"], g))
        self.task_out[e] = torch.as_tensor(out)
        if self._task_tab is not None:
            op, a, b = self._task_tab
            self.task_vec[e] = torch.as_tensor([self._value(o, int(op[k]), int(a[k]), int(b[k]), None, False) for k in range(len(op))])

    # ---- stepping
    def forward(self):
        for e in range(self.n_env):
            self._push(e); self.o[e].forward()
            self.qacc[e] = torch.as_tensor(self.o[e].qacc.copy())
            if bool(self.obs_fresh[e]):
                self._sample_obs(e)
            self._sample_task(e)

    def reset_envs(self, mask=None, qpos=None):
        """b2s_reset_envs: masked environments take the sampled state, are cleared, forwarded, and get their controller rebuilt"""
        for e in range(self.n_env):
            if mask is not None and not bool(mask[e]):
                continue
            self.qpos[e] = torch.as_tensor(np.asarray(self.model.qpos0)) if qpos is None else qpos[e].to(torch.float64)
            self.qvel[e] = 0; self.qacc[e] = 0; self.qacc_warmstart[e] = 0; self.ctrl[e] = 0; self.time[e] = 0
            self.warn[e] = 0; self.obs_fresh[e] = 1
            self._push(e); self.o[e].forward()
            self.qacc[e] = torch.as_tensor(self.o[e].qacc.copy())
            self._sample_obs(e)
            self._sample_task(e)
            if self._cfg is not None:
                self.o[e].ctrl_reset()

    def ctrl_reset(self, mask=None):
        for e in range(self.n_env):
            if mask is None or bool(mask[e]):
                self.o[e].ctrl_reset()

    def env_step(self, action, n_substeps):
        act = action.numpy().astype(np.float64)
        for e in range(self.n_env):
            self._push(e)
            self.o[e].env_step(act[e], n_substeps)
            self._pull(e)
            self._sample_obs(e)   # last substep: poses of its step1, qpos / qvel after its step2
            self._sample_task(e)

    def get_state(self):
        return torch.cat([self.time[:, None], self.qpos, self.qvel], dim=1)

    # controller state arrays of BatchedSim (read-only views of the oracles' CtrlState)
    def _cs(self, field, n):
        return torch.as_tensor([[getattr(o.ctrl_state, field)[k] for k in range(n)] for o in self.o], dtype=torch.float64)

    ctrl_goal_pos = property(lambda self: self._cs("goal_pos", 3))
    ctrl_goal_ori = property(lambda self: self._cs("goal_ori", 9))
    ctrl_initial_joint = property(lambda self: self._cs("initial_joint", 8))
    ctrl_grip_state = property(lambda self: self._cs("grip_action", 4))
