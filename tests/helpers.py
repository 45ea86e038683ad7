"""Shared test helpers.  The oracle (oracle/) is the checker; the product is only reached through the package."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
XMLS = os.path.join(ROOT, "mujoco_rl_manipulate_unknown_objects_b200", "assets", "xmls")
SCENES = ["sugar_cube", "sand_ball", "bread_crumb", "acorn"]


def scene_file(scene):
    return os.path.join(XMLS, scene + "_env.xml")


def oracle_model_from_product(cm):
    """Model dict for oracle.engine.Model built from the PRODUCT compiler's export (grs_model_get), so that the
    oracle engine and the CUDA kernels integrate exactly the same model (same principal frames, same hull
    vertex numbering).  The product compiler itself is checked against oracle/mjcf.py in test_model_compiler.py."""
    sz = cm.sizes
    md = {k: sz[k] for k in ("nbody", "njnt", "nq", "nv", "nu", "ngeom", "nmesh")}
    md["iterations"] = int(cm.get("iterations")[0])
    md["cone_elliptic"] = int(cm.get("cone_elliptic")[0])
    for k in ("timestep", "impratio", "tolerance"):
        md[k] = float(cm.get(k)[0])
    md["gravity"] = cm.get("gravity")
    for k in ("body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "jnt_pos", "jnt_axis",
              "jnt_range", "jnt_solref", "jnt_solimp", "qpos0", "dof_armature", "dof_damping", "geom_pos", "geom_quat",
              "geom_friction", "geom_margin", "geom_gap", "geom_solref", "geom_solimp", "geom_rbound", "act_gear",
              "act_ctrlrange", "body_invweight0", "dof_invweight0"):
        md[k] = cm.get(k)
    for k in ("body_parentid", "body_weldid", "body_jntadr", "body_jntnum", "body_dofadr", "body_dofnum", "jnt_type",
              "jnt_bodyid", "jnt_qposadr", "jnt_dofadr", "jnt_limited", "dof_bodyid", "dof_jntid", "dof_parentid",
              "geom_type", "geom_bodyid", "geom_meshid", "geom_condim", "act_dofid", "pair_geom1", "pair_geom2"):
        md[k] = cm.get(k)
    md["meaninertia"] = float(cm.get("meaninertia")[0])
    md["body_names"] = cm.names("body")
    md["geom_names"] = cm.names("geom")
    meshes = []
    for n in cm.names("mesh"):
        hv = cm.get("hull_verts:" + n).reshape(-1, 3)
        adr = cm.get("hull_adjadr:" + n)
        adj = cm.get("hull_adj:" + n)
        meshes.append(dict(hull_verts=hv, hull_adj=[adj[adr[i]:adr[i + 1]].tolist() for i in range(len(hv))]))
    md["meshes"] = meshes
    return md


def oracle_model(cm, finalize=True):
    from oracle import engine
    md = oracle_model_from_product(cm)
    if finalize:
        return engine.Model(md)
    # keep the product's mj_setConst outputs instead of recomputing them
    return engine.Model(md, finalize=True, overrides=dict(body_invweight0=md["body_invweight0"], dof_invweight0=md["dof_invweight0"],
                                                          meaninertia=[md["meaninertia"]]))


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
