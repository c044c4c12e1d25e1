"""`mujoco`-compatible shim over the CPU oracle (TEST INFRASTRUCTURE, like everything under oracle/).

Exposes just the part of the `mujoco` Python API that robosuite's hot path touches (utils/binding_utils.py: MjModel /
MjData construction, mj_step1 / mj_step2 / mj_step / mj_forward / mj_resetData, mj_jacSite / mj_jacBody / mj_jacGeom,
mj_fullM, mj_name2id / mj_id2name, the mjtObj / mjtJoint enums).  With this directory ahead of the real package on
sys.path the UNMODIFIED reference Python stack (environments, robots, controllers, observables) runs on the oracle's
physics.  Used by tools/gen_env_golden.py to produce golden observation / reward / controller vectors from the
reference's own code; it is not the Google engine and says so: __version__ carries the suffix "+b2s.oracle".
"""
import enum
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

__version__ = "3.3.0+b2s.oracle"


class mjtObj(enum.IntEnum):
    mjOBJ_UNKNOWN = 0; mjOBJ_BODY = 1; mjOBJ_XBODY = 2; mjOBJ_JOINT = 3; mjOBJ_DOF = 4; mjOBJ_GEOM = 5; mjOBJ_SITE = 6
    mjOBJ_CAMERA = 7; mjOBJ_LIGHT = 8; mjOBJ_MESH = 10; mjOBJ_TENDON = 18; mjOBJ_ACTUATOR = 19; mjOBJ_SENSOR = 20


class mjtJoint(enum.IntEnum):
    mjJNT_FREE = 0; mjJNT_BALL = 1; mjJNT_SLIDE = 2; mjJNT_HINGE = 3


class _Any:
    """placeholder for enums / classes of the rendering API, which this shim does not provide"""

    def __getattr__(self, name):
        return 0

    def __call__(self, *a, **k):
        raise NotImplementedError("rendering is not available in the oracle shim")


mjtRndFlag = mjtCamera = mjtFramebuffer = mjtFontScale = mjtCatBit = _Any()
MjrRect = MjvScene = MjvPerturb = MjvOption = MjvCamera = MjrContext = _Any()

_KIND = {mjtObj.mjOBJ_BODY: "body", mjtObj.mjOBJ_JOINT: "joint", mjtObj.mjOBJ_GEOM: "geom", mjtObj.mjOBJ_SITE: "site",
         mjtObj.mjOBJ_CAMERA: "camera", mjtObj.mjOBJ_LIGHT: "light", mjtObj.mjOBJ_MESH: "mesh", mjtObj.mjOBJ_TENDON: "tendon",
         mjtObj.mjOBJ_ACTUATOR: "actuator", mjtObj.mjOBJ_SENSOR: "sensor"}

# attributes robosuite reads through its MjModel wrapper (the wrapper's metaclass delegates every name in dir(MjModel))
_MODEL_FIELDS = [
    "nq", "nv", "nu", "na", "nbody", "njnt", "ngeom", "nsite", "ncam", "nlight", "nsensor", "ntendon", "nmesh", "nmocap",
    "names", "name_bodyadr", "name_jntadr", "name_geomadr", "name_siteadr", "name_camadr", "name_lightadr",
    "name_actuatoradr", "name_sensoradr", "name_tendonadr", "name_meshadr",
    "body_pos", "body_quat", "body_mass", "body_parentid", "body_mocapid", "body_jntadr", "body_jntnum", "body_inertia",
    "body_ipos", "body_iquat", "jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_range", "jnt_axis", "jnt_bodyid", "jnt_pos",
    "dof_damping", "dof_armature", "dof_frictionloss", "dof_bodyid", "dof_jntid",
    "geom_type", "geom_size", "geom_pos", "geom_quat", "geom_rgba", "geom_bodyid", "geom_friction", "geom_contype",
    "geom_conaffinity", "geom_group", "geom_solref", "geom_solimp", "geom_condim", "geom_margin", "geom_gap",
    "site_pos", "site_quat", "site_size", "site_rgba", "site_bodyid",
    "actuator_ctrlrange", "actuator_gear", "actuator_trnid", "actuator_forcerange", "actuator_biasprm", "actuator_gainprm",
    "actuator_biastype", "sensor_dim", "sensor_adr", "sensor_type", "opt", "qpos0", "stat", "body_geomadr", "body_geomnum",
    "body_rootid", "body_weldid", "geom_dataid", "geom_rbound", "geom_aabb", "mesh_vert", "mesh_vertadr", "mesh_vertnum",
]


class _Opt:
    pass


class MjModel:
    """compiled model (robosuite_b200.mjcf.compiler.Model) behind mujoco.MjModel's attribute names"""

    for _f in _MODEL_FIELDS:
        locals()[_f] = None
    del _f

    @classmethod
    def from_xml_string(cls, xml, assets=None):
        from robosuite_b200.mjcf.compiler import compile_mjcf

        self = cls._wrap(compile_mjcf(xml))
        self._xml = xml
        return self

    @classmethod
    def from_xml_path(cls, path, assets=None):
        with open(path) as f:
            return cls.from_xml_string(f.read())

    @classmethod
    def _wrap(cls, m):
        self = cls()
        self._m = m
        for f in _MODEL_FIELDS:
            if hasattr(m, f):
                setattr(self, f, getattr(m, f))
        names = m.names
        self._names = {k: list(v) for k, v in names.items()}
        for k, n in (("nbody", "body"), ("njnt", "joint"), ("ngeom", "geom"), ("nsite", "site"), ("nu", "actuator")):
            setattr(self, k, len(self._names.get(n, [])) if getattr(m, k, None) is None else int(getattr(m, k)))
        self.ncam = len(self._names.get("camera", [])); self.nlight = len(self._names.get("light", []))
        self.nsensor = len(self._names.get("sensor", [])); self.ntendon = 0; self.nmesh = len(self._names.get("mesh", []))
        self.na = 0
        self.nmocap = int(getattr(m, "nmocap", 0) or 0)
        if getattr(m, "body_mocapid", None) is None:
            self.body_mocapid = -np.ones(self.nbody, dtype=np.int32)
        if getattr(m, "sensor_dim", None) is None:
            self.sensor_dim = 3 * np.ones(self.nsensor, dtype=np.int32)
            self.sensor_adr = 3 * np.arange(self.nsensor, dtype=np.int32)
        for k in ("body", "jnt", "geom", "site", "cam", "light", "actuator", "sensor", "tendon", "mesh"):
            setattr(self, "name_%sadr" % k, np.zeros(1, dtype=np.int32))
        self.names = b""
        gb = np.asarray(m.geom_bodyid)
        self.body_geomnum = np.array([int((gb == b).sum()) for b in range(self.nbody)], dtype=np.int32)
        self.body_geomadr = np.array([int(np.nonzero(gb == b)[0][0]) if (gb == b).any() else -1 for b in range(self.nbody)], dtype=np.int32)
        opt = _Opt()
        opt.timestep = float(m.opt_timestep)
        opt.gravity = np.asarray(getattr(m, "opt_gravity", [0, 0, -9.81]), dtype=np.float64)
        self.opt = opt
        self._pose_key = None
        return self

    def _blob(self):
        from robosuite_b200.mjcf.compiler import pack_model

        # robosuite edits body_pos / body_quat in place at reset (door.py:417-427, pick_place.py:700-709)
        self._m.body_pos = np.asarray(self.body_pos); self._m.body_quat = np.asarray(self.body_quat)
        return pack_model(self._m)

    def _pose_signature(self):
        return np.asarray(self.body_pos).tobytes() + np.asarray(self.body_quat).tobytes()


class _Contact:
    __slots__ = ("geom1", "geom2", "dist", "pos", "frame", "dim", "geom")

    def __init__(self, c):
        self.geom1, self.geom2, self.dist, self.pos, self.frame, self.dim = c["geom1"], c["geom2"], c["dist"], c["pos"], c["frame"], c["dim"]
        self.geom = (self.geom1, self.geom2)


_DATA_FIELDS = ["qpos", "qvel", "qacc", "qacc_warmstart", "ctrl", "time", "xpos", "xquat", "xmat", "xipos", "site_xpos", "site_xmat",
                "geom_xpos", "geom_xmat", "qfrc_bias", "qfrc_applied", "qfrc_passive", "qfrc_actuator", "qfrc_constraint", "qM",
                "actuator_force", "sensordata", "ncon", "contact", "mocap_pos", "mocap_quat", "cvel", "userdata", "act"]


class MjData:
    for _f in _DATA_FIELDS:
        locals()[_f] = None
    del _f

    def __init__(self, model):
        self.__dict__["_model"] = model
        self._build()

    def _build(self, keep_state=None):
        from oracle.pyoracle import Oracle

        model = self._model
        o = Oracle(model._blob())
        model._pose_key = model._pose_signature()
        self.__dict__["_o"] = o
        for f in ("qpos", "qvel", "qacc", "qacc_warmstart", "ctrl", "xpos", "xquat", "xmat", "xipos", "site_xpos", "site_xmat",
                  "geom_xpos", "geom_xmat", "qfrc_bias", "qfrc_applied", "qfrc_passive", "qfrc_actuator", "qfrc_constraint",
                  "actuator_force", "sensordata", "mocap_pos", "mocap_quat", "cvel"):
            self.__dict__[f] = getattr(o, f)
        self.__dict__["qM"] = o.M  # dense; mj_fullM below copies it
        self.__dict__["userdata"] = np.zeros(0); self.__dict__["act"] = np.zeros(0)
        if keep_state is not None:
            t, qp, qv, ct, ws = keep_state
            o.time = t; o.qpos[:] = qp; o.qvel[:] = qv; o.ctrl[:] = ct; o.qacc_warmstart[:] = ws

    def _sync_model(self):
        """rebuild the oracle instance if robosuite changed body poses in the model (state is carried over)"""
        m = self._model
        if m._pose_key != m._pose_signature():
            o = self._o
            self._build((o.time, o.qpos.copy(), o.qvel.copy(), o.ctrl.copy(), o.qacc_warmstart.copy()))

    time = property(lambda self: self._o.time, lambda self, v: setattr(self._o, "time", v))
    ncon = property(lambda self: self._o.ncon)

    @property
    def contact(self):
        return [_Contact(c) for c in self._o.contacts()]

    def __setattr__(self, k, v):
        if k in self.__dict__ and isinstance(self.__dict__[k], np.ndarray):
            self.__dict__[k][...] = v  # mujoco semantics: assignment copies into the engine's array
        else:
            object.__setattr__(self, k, v)


def mj_forward(m, d):
    d._sync_model(); d._o.forward()


def mj_step(m, d, nstep=1):
    d._sync_model()
    for _ in range(nstep):
        d._o.step()


def mj_step1(m, d):
    d._sync_model(); d._o.step1()


def mj_step2(m, d):
    d._o.step2()


def mj_resetData(m, d):
    d._sync_model(); d._o.reset_data()


def mj_fullM(m, dst, qM):
    dst[...] = np.asarray(qM).reshape(dst.shape)


def _jac(d, jacp, jacr, point, body):
    jp, jr = d._o.jac(np.asarray(point, dtype=np.float64), int(body))
    if jacp is not None:
        jacp[...] = jp.reshape(jacp.shape)
    if jacr is not None:
        jacr[...] = jr.reshape(jacr.shape)


def mj_jacSite(m, d, jacp, jacr, site):
    _jac(d, jacp, jacr, d._o.site_xpos[site], m.site_bodyid[site])


def mj_jacBody(m, d, jacp, jacr, body):
    _jac(d, jacp, jacr, d._o.xpos[body], body)


def mj_jacGeom(m, d, jacp, jacr, geom):
    _jac(d, jacp, jacr, d._o.geom_xpos[geom], m.geom_bodyid[geom])


def mj_id2name(m, objtype, i):
    names = m._names.get(_KIND.get(objtype, ""), [])
    return names[i] if 0 <= i < len(names) else None


def mj_name2id(m, objtype, name):
    names = m._names.get(_KIND.get(objtype, ""), [])
    return names.index(name) if name in names else -1


def mju_mat2Quat(quat, mat):
    from scipy.spatial.transform import Rotation

    q = Rotation.from_matrix(np.asarray(mat).reshape(3, 3)).as_quat()
    quat[...] = [q[3], q[0], q[1], q[2]]


def mj_saveLastXML(filename, m, *a):
    """writes the MJCF the model was compiled from (the engine would re-serialise its compiled model)"""
    path = filename.decode() if isinstance(filename, bytes) else filename
    with open(path, "w") as f:
        f.write(getattr(m, "_xml", "") or "")
    return 1


def __getattr__(name):  # anything else of the rendering / visualisation API
    return _Any()
