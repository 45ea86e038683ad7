"""ORACLE (test infrastructure, not product code) — ctypes front end of oracle/engine.c.

Physics parity UNPINNED (see engine.h).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MAXB, MAXJ, MAXV, MAXQ, MAXG, MAXM, MAXU, MAXPAIR, MAXCON = 12, 12, 16, 16, 8, 8, 8, 32, 64
MAXEFC = MAXCON * 4 + 2 * MAXJ


def build(force=False):
    so = os.path.join(_HERE, "_build", "liboracle.so")
    src = [os.path.join(_HERE, "engine.c"), os.path.join(_HERE, "engine.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


class OContact(C.Structure):
    _fields_ = [("dist", C.c_double), ("pos", C.c_double * 3), ("frame", C.c_double * 9),
                ("includemargin", C.c_double), ("friction", C.c_double * 5), ("solref", C.c_double * 2),
                ("solimp", C.c_double * 5), ("mu", C.c_double),
                ("dim", C.c_int), ("geom1", C.c_int), ("geom2", C.c_int), ("efc_address", C.c_int)]


class OData(C.Structure):
    _fields_ = [("qpos", C.c_double * MAXQ), ("qvel", C.c_double * MAXV), ("ctrl", C.c_double * MAXU),
                ("qacc", C.c_double * MAXV), ("qacc_warmstart", C.c_double * MAXV),
                ("xfrc_applied", C.c_double * (MAXB * 6)),
                ("xpos", C.c_double * (MAXB * 3)), ("xquat", C.c_double * (MAXB * 4)), ("xmat", C.c_double * (MAXB * 9)),
                ("xipos", C.c_double * (MAXB * 3)), ("ximat", C.c_double * (MAXB * 9)),
                ("xanchor", C.c_double * (MAXJ * 3)), ("xaxis", C.c_double * (MAXJ * 3)),
                ("geom_xpos", C.c_double * (MAXG * 3)), ("geom_xmat", C.c_double * (MAXG * 9)),
                ("subtree_com", C.c_double * (MAXB * 3)), ("cinert", C.c_double * (MAXB * 10)),
                ("cdof", C.c_double * (MAXV * 6)), ("cdof_dot", C.c_double * (MAXV * 6)), ("cvel", C.c_double * (MAXB * 6)),
                ("qM", C.c_double * (MAXV * MAXV)), ("qLD", C.c_double * (MAXV * MAXV)),
                ("ncon", C.c_int), ("contact", OContact * MAXCON),
                ("qfrc_bias", C.c_double * MAXV), ("qfrc_passive", C.c_double * MAXV), ("qfrc_actuator", C.c_double * MAXV),
                ("qfrc_applied", C.c_double * MAXV), ("qfrc_smooth", C.c_double * MAXV), ("qacc_smooth", C.c_double * MAXV),
                ("qfrc_constraint", C.c_double * MAXV),
                ("nefc", C.c_int), ("efc_type", C.c_int * MAXEFC), ("efc_id", C.c_int * MAXEFC),
                ("efc_J", C.c_double * (MAXEFC * MAXV)), ("efc_pos", C.c_double * MAXEFC), ("efc_margin", C.c_double * MAXEFC),
                ("efc_diagApprox", C.c_double * MAXEFC), ("efc_R", C.c_double * MAXEFC), ("efc_D", C.c_double * MAXEFC),
                ("efc_KBIP", C.c_double * (MAXEFC * 4)), ("efc_vel", C.c_double * MAXEFC), ("efc_aref", C.c_double * MAXEFC),
                ("efc_force", C.c_double * MAXEFC),
                ("solver_iter", C.c_int), ("solver_improvement", C.c_double), ("solver_gradient", C.c_double),
                ("time", C.c_double)]


class OEnvCfg(C.Structure):
    _fields_ = [("max_steps", C.c_int), ("time_horizon", C.c_int), ("include_roll", C.c_int), ("her_buffer", C.c_int),
                ("direction", C.c_int), ("pos_tolerance", C.c_double), ("grasp_tolerance", C.c_double),
                ("max_translation", C.c_double), ("max_rotation", C.c_double)]


class OStepOut(C.Structure):
    _fields_ = [("reward", C.c_double), ("achieved_goal", C.c_double * 2), ("desired_goal", C.c_double * 2),
                ("done", C.c_int), ("status", C.c_int), ("grasp", C.c_int), ("pheromone", C.c_int),
                ("object_grasped", C.c_int), ("gripper_open", C.c_int),
                ("reached_target", C.c_int), ("reached_initial", C.c_int), ("fail", C.c_int),
                ("nsub_a", C.c_int), ("nsub_b", C.c_int), ("nsub_c", C.c_int),
                ("total_distance", C.c_double), ("line_distance", C.c_double), ("init_obj_pos", C.c_double * 3),
                ("final_obj_pos", C.c_double * 3), ("gripper_pos", C.c_double * 3), ("target_qpos", C.c_double * 5)]


class OEnvS(C.Structure):
    _fields_ = [("m", C.c_void_p), ("d", C.POINTER(OData)), ("cfg", OEnvCfg), ("target_dir", C.c_double * 2),
                ("body_ee", C.c_int), ("body_object", C.c_int), ("finger1_body", C.c_int * 2), ("finger2_body", C.c_int * 2),
                ("gripper_open", C.c_int), ("episode_step", C.c_int), ("status", C.c_int), ("total_substeps", C.c_long)]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_model_new.restype = C.c_void_p
        L.orc_model_free.argtypes = [C.c_void_p]
        L.orc_set_d.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
        L.orc_set_i.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
        L.orc_get_d.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
        L.orc_set_mesh.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_model_finalize.argtypes = [C.c_void_p]
        L.orc_data_new.restype = C.POINTER(OData)
        L.orc_data_new.argtypes = [C.c_void_p]
        L.orc_data_free.argtypes = [C.POINTER(OData)]
        for f in (L.orc_reset, L.orc_forward_position, L.orc_step):
            f.argtypes = [C.c_void_p, C.POINTER(OData)]
        L.orc_jac_body.argtypes = [C.c_void_p, C.POINTER(OData), C.c_int, C.c_void_p, C.c_void_p]
        L.orc_env_new.restype = C.POINTER(OEnvS)
        L.orc_env_new.argtypes = [C.c_void_p, C.POINTER(OEnvCfg), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_env_free.argtypes = [C.POINTER(OEnvS)]
        L.orc_env_reset.argtypes = [C.POINTER(OEnvS), C.POINTER(OStepOut)]
        L.orc_env_step.argtypes = [C.POINTER(OEnvS), C.c_void_p, C.POINTER(OStepOut)]
        L.orc_get_target_pose.argtypes = [C.POINTER(OEnvS), C.c_void_p, C.c_void_p]
        L.orc_check_grasp.argtypes = [C.POINTER(OEnvS)]
        L.orc_pheromone_level.argtypes = [C.POINTER(OEnvS)]
        L.orc_agent_reward.restype = C.c_double
        L.orc_agent_reward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_rollout_threads.restype = C.c_long
        L.orc_rollout_threads.argtypes = [C.c_void_p, C.POINTER(OEnvCfg), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_long)]
        L.orc_rollout_window.restype = C.c_long
        L.orc_rollout_window.argtypes = [C.c_void_p, C.POINTER(OEnvCfg), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_long), C.POINTER(C.c_double), C.POINTER(C.c_long)]
        _LIB = L
    return _LIB


_D_FIELDS = ["gravity", "body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia",
             "jnt_pos", "jnt_axis", "jnt_range", "jnt_solref", "jnt_solimp", "qpos0", "dof_armature", "dof_damping",
             "geom_pos", "geom_quat", "geom_friction", "geom_margin", "geom_gap", "geom_solref", "geom_solimp",
             "geom_rbound", "geom_size", "act_gear", "act_ctrlrange"]
_I_FIELDS = ["body_parentid", "body_weldid", "body_jntadr", "body_jntnum", "body_dofadr", "body_dofnum",
             "jnt_type", "jnt_bodyid", "jnt_qposadr", "jnt_dofadr", "jnt_limited", "dof_bodyid", "dof_jntid",
             "dof_parentid", "geom_type", "geom_bodyid", "geom_meshid", "geom_condim", "act_dofid",
             "pair_geom1", "pair_geom2"]


class Model:
    """Oracle model built from a dict of mjModel-named numpy arrays (oracle.mjcf.compile_mjcf output, or the
    product compiler's export — the engine is agnostic)."""

    def __init__(self, md, finalize=True, overrides=None):
        L = lib()
        self.md = md
        self.ptr = L.orc_model_new()
        for k in ("nbody", "njnt", "nq", "nv", "nu", "ngeom", "nmesh", "iterations", "cone_elliptic"):
            L.orc_set_i(self.ptr, k.encode(), np.array([md[k]], np.int32).ctypes.data, 1)
        L.orc_set_i(self.ptr, b"npair", np.array([len(md["pair_geom1"])], np.int32).ctypes.data, 1)
        for k in ("timestep", "impratio", "tolerance"):
            L.orc_set_d(self.ptr, k.encode(), np.array([md[k]], np.float64).ctypes.data, 1)
        for k in _D_FIELDS:
            a = np.ascontiguousarray(md[k], dtype=np.float64).reshape(-1)
            assert L.orc_set_d(self.ptr, k.encode(), a.ctypes.data, a.size) == 0, k
        for k in _I_FIELDS:
            a = np.ascontiguousarray(md[k], dtype=np.int32).reshape(-1)
            assert L.orc_set_i(self.ptr, k.encode(), a.ctypes.data, a.size) == 0, k
        for i, ms in enumerate(md["meshes"]):
            v = np.ascontiguousarray(ms["hull_verts"], dtype=np.float64)
            adr = np.zeros(len(v) + 1, np.int32)
            adr[1:] = np.cumsum([len(a) for a in ms["hull_adj"]])
            adj = np.array([x for a in ms["hull_adj"] for x in a], np.int32)
            if adj.size == 0:
                adj = np.zeros(1, np.int32)
            assert L.orc_set_mesh(self.ptr, i, len(v), v.ctypes.data, adr.ctypes.data, adj.ctypes.data) == 0
        if finalize:
            L.orc_model_finalize(self.ptr)
        for k, v in (overrides or {}).items():
            a = np.ascontiguousarray(v, dtype=np.float64).reshape(-1)
            assert L.orc_set_d(self.ptr, k.encode(), a.ctypes.data, a.size) == 0, k
        for k in ("nbody", "njnt", "nq", "nv", "nu", "ngeom"):
            setattr(self, k, int(md[k]))

    def body_id(self, name):
        return self.md["body_names"].index(name)

    def derived(self, name):
        """body_invweight0 / dof_invweight0 / meaninertia as computed by orc_model_finalize (mj_setConst)."""
        out = np.zeros(64)
        n = lib().orc_get_d(self.ptr, name.encode(), out.ctypes.data, out.size)
        assert n > 0, name
        return out[:n].copy()

    def __del__(self):
        try:
            lib().orc_model_free(self.ptr)
        except Exception:
            pass


def model_dict_from_export(cm):
    """Model dict (mjModel-named arrays) from any object with .get(name) / .names(kind) / .sizes — in practice the
    PRODUCT compiler's export (grs_model_get), so that this engine integrates exactly the model the CUDA kernels
    integrate (same principal frames, same hull-vertex numbering).  The product compiler itself is checked against
    oracle/mjcf.py in tests/test_model_compiler.py."""
    sz = cm.sizes
    md = {k: sz[k] for k in ("nbody", "njnt", "nq", "nv", "nu", "ngeom", "nmesh")}
    md["iterations"] = int(cm.get("iterations")[0])
    md["cone_elliptic"] = int(cm.get("cone_elliptic")[0])
    for k in ("timestep", "impratio", "tolerance"):
        md[k] = float(cm.get(k)[0])
    md["gravity"] = cm.get("gravity")
    for k in _D_FIELDS:
        md[k] = cm.get(k)
    for k in _I_FIELDS:
        md[k] = cm.get(k)
    md["body_names"] = cm.names("body")
    md["geom_names"] = cm.names("geom")
    meshes = []
    for n in cm.names("mesh"):
        hv = cm.get("hull_verts:" + n).reshape(-1, 3)
        adr, adj = cm.get("hull_adjadr:" + n), cm.get("hull_adj:" + n)
        meshes.append(dict(hull_verts=hv, hull_adj=[adj[adr[i]:adr[i + 1]].tolist() for i in range(len(hv))]))
    md["meshes"] = meshes
    return md


def _view(arr, n, shape=None):
    a = np.ctypeslib.as_array(arr)[:n]
    return a.reshape(shape) if shape else a


class Data:
    def __init__(self, model, _borrow=None):
        self.model = model
        self._owned = _borrow is None
        self.p = lib().orc_data_new(model.ptr) if _borrow is None else _borrow
        d, m = self.p.contents, model
        self.qpos = _view(d.qpos, m.nq)
        self.qvel = _view(d.qvel, m.nv)
        self.ctrl = _view(d.ctrl, m.nu)
        self.qacc = _view(d.qacc, m.nv)
        self.qacc_warmstart = _view(d.qacc_warmstart, m.nv)
        self.xfrc_applied = _view(d.xfrc_applied, m.nbody * 6, (m.nbody, 6))
        self.xpos = _view(d.xpos, m.nbody * 3, (m.nbody, 3))
        self.xquat = _view(d.xquat, m.nbody * 4, (m.nbody, 4))
        self.xmat = _view(d.xmat, m.nbody * 9, (m.nbody, 9))
        self.xipos = _view(d.xipos, m.nbody * 3, (m.nbody, 3))
        self.geom_xpos = _view(d.geom_xpos, m.ngeom * 3, (m.ngeom, 3))
        self.geom_xmat = _view(d.geom_xmat, m.ngeom * 9, (m.ngeom, 9))
        self.qM = _view(d.qM, m.nv * m.nv, (m.nv, m.nv))
        self.qfrc_bias = _view(d.qfrc_bias, m.nv)
        self.qfrc_smooth = _view(d.qfrc_smooth, m.nv)
        self.qacc_smooth = _view(d.qacc_smooth, m.nv)
        self.qfrc_constraint = _view(d.qfrc_constraint, m.nv)

    @property
    def ncon(self):
        return self.p.contents.ncon

    @property
    def nefc(self):
        return self.p.contents.nefc

    @property
    def solver_iter(self):
        return self.p.contents.solver_iter

    def contacts(self):
        d = self.p.contents
        return [dict(geom1=c.geom1, geom2=c.geom2, dist=c.dist, pos=np.array(c.pos[:]), frame=np.array(c.frame[:]),
                     dim=c.dim, mu=c.mu, friction=np.array(c.friction[:]))
                for c in d.contact[:d.ncon]]

    def efc(self):
        d, nv, n = self.p.contents, self.model.nv, self.p.contents.nefc
        return dict(J=np.array(d.efc_J[:n * nv]).reshape(n, nv), aref=np.array(d.efc_aref[:n]), D=np.array(d.efc_D[:n]),
                    R=np.array(d.efc_R[:n]), pos=np.array(d.efc_pos[:n]), force=np.array(d.efc_force[:n]),
                    type=np.array(d.efc_type[:n]))

    def reset(self):
        lib().orc_reset(self.model.ptr, self.p)

    def forward_position(self):
        lib().orc_forward_position(self.model.ptr, self.p)

    def step(self, n=1):
        for _ in range(n):
            lib().orc_step(self.model.ptr, self.p)

    def jac_body(self, body):
        nv = self.model.nv
        jp, jr = np.zeros((3, nv)), np.zeros((3, nv))
        lib().orc_jac_body(self.model.ptr, self.p, body, jp.ctypes.data, jr.ctypes.data)
        return jp, jr

    def __del__(self):
        try:
            if self._owned:
                lib().orc_data_free(self.p)
        except Exception:
            pass


def default_cfg(**kw):
    """config/base_config.py:13-43 defaults for the fields the env reads."""
    c = OEnvCfg(max_steps=400, time_horizon=400, include_roll=1, her_buffer=0, direction=0,
                pos_tolerance=0.002, grasp_tolerance=0.03, max_translation=0.05, max_rotation=0.15)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def _finger_ids(model):
    f1 = np.array([model.body_id("left_inner_knuckle"), model.body_id("left_inner_finger")], np.int32)
    f2 = np.array([model.body_id("right_inner_knuckle"), model.body_id("right_inner_finger")], np.int32)
    return f1, f2


class Env:
    """C restatement of RobotEnv.reset/step (robot_env.py:56-241) without the rendered image channels."""

    def __init__(self, model, **cfg):
        self.model = model
        self.cfg = default_cfg(**cfg)
        self._f1, self._f2 = _finger_ids(model)
        self.p = lib().orc_env_new(model.ptr, C.byref(self.cfg), model.body_id("ee"), model.body_id("object"),
                                   self._f1.ctypes.data, self._f2.ctypes.data)
        self.data = Data(model, _borrow=self.p.contents.d)
        self.qpos, self.qvel, self.ctrl, self.xpos = self.data.qpos, self.data.qvel, self.data.ctrl, self.data.xpos

    def reset(self):
        out = OStepOut()
        lib().orc_env_reset(self.p, C.byref(out))
        return out

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        out = OStepOut()
        lib().orc_env_step(self.p, a.ctypes.data, C.byref(out))
        return out

    def target_pose(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        t = np.zeros(5)
        lib().orc_get_target_pose(self.p, a.ctypes.data, t.ctypes.data)
        return t

    def __del__(self):
        try:
            lib().orc_env_free(self.p)
        except Exception:
            pass


def rollout_threads(model, actions, nthreads, **cfg):
    """actions: [nenv, nsteps, 6] float64.  Returns (substeps, transitions, reward_sum)."""
    c = default_cfg(**cfg)
    f1, f2 = _finger_ids(model)
    a = np.ascontiguousarray(actions, dtype=np.float64)
    rs, tr = C.c_double(0), C.c_long(0)
    sub = lib().orc_rollout_threads(model.ptr, C.byref(c), model.body_id("ee"), model.body_id("object"), f1.ctypes.data,
                                    f2.ctypes.data, a.shape[0], a.shape[1], a.ctypes.data, nthreads, C.byref(rs), C.byref(tr))
    return sub, tr.value, rs.value


def rollout_window(model, actions, nthreads, skip, **cfg):
    """actions: [nenv, nsteps, 6] float64; the first `skip` transitions of every env are rolled from reset but neither counted
    nor timed.  Returns dict(substeps, transitions, reward_sum, seconds, ncon_sum) over the counted window."""
    c = default_cfg(**cfg)
    f1, f2 = _finger_ids(model)
    a = np.ascontiguousarray(actions, dtype=np.float64)
    rs, tr, ts, nc = C.c_double(0), C.c_long(0), C.c_double(0), C.c_long(0)
    sub = lib().orc_rollout_window(model.ptr, C.byref(c), model.body_id("ee"), model.body_id("object"), f1.ctypes.data, f2.ctypes.data,
                                   a.shape[0], a.shape[1], int(skip), a.ctypes.data, nthreads, C.byref(rs), C.byref(tr), C.byref(ts), C.byref(nc))
    return dict(substeps=sub, transitions=tr.value, reward_sum=rs.value, seconds=ts.value, ncon_sum=nc.value)
