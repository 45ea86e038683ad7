"""Shared test helpers.  The oracle (oracle/) is the checker; the product is only reached through the package."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
XMLS = os.path.join(ROOT, "mujoco_rl_manipulate_unknown_objects_b200", "assets", "xmls")
SCENES = ["sugar_cube", "sand_ball", "bread_crumb", "acorn"]


def scene_file(scene):
    return os.path.join(XMLS, scene + "_env.xml")


def oracle_model(cm):
    """Oracle engine model built from the PRODUCT compiler's export (see oracle.engine.model_dict_from_export)."""
    from oracle import engine
    return engine.Model(engine.model_dict_from_export(cm))


def oracle_model_independent(scene):
    """Oracle engine model built by the ORACLE's own compiler (oracle/mjcf.py) from the scene file — nothing of the product."""
    from oracle import engine, mjcf
    name = scene if scene.endswith(".xml") else (scene + ".xml" if scene == "gripper_two_fingers" else scene + "_env.xml")
    return engine.Model(mjcf.compile_mjcf(os.path.join(XMLS, name)))


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
