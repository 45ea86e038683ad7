"""CPU: the primitive box of gripper_two_fingers.xml (reference xmls/gripper_two_fingers.xml:130-135).

Pins the oracle's restatement of mjc_PlaneBox (engine_collision_primitive.c) on analytic facts: a box resting flat on
the plane touches it with its 4 bottom corners in corner-index order, settles at the penetration where 4 contact forces
carry its weight, and a tilted box touches with fewer corners.  The GPU path is compared with the oracle on this scene
in test_physics_parity.py."""
import os

import numpy as np

from helpers import XMLS
from oracle import engine, mjcf

BOX_XML = os.path.join(XMLS, "gripper_two_fingers.xml")


def _model():
    return engine.Model(mjcf.compile_mjcf(BOX_XML))


def test_box_compiles_with_analytic_inertia():
    md = mjcf.compile_mjcf(BOX_XML)
    g = int(np.flatnonzero(md["geom_type"] == mjcf.GEOM_BOX)[0])
    np.testing.assert_allclose(md["geom_size"][g], [0.2, 0.2, 0.2])
    b = int(md["geom_bodyid"][g])
    assert abs(md["body_mass"][b] - 1.0) < 1e-12
    np.testing.assert_allclose(md["body_inertia"][b], [1.0 / 3 * 0.08] * 3, rtol=1e-9)  # m/3 (y^2 + z^2)
    assert abs(md["geom_rbound"][g] - np.sqrt(0.12)) < 1e-9


def test_plane_box_contacts_at_rest_and_tilted():
    md = mjcf.compile_mjcf(BOX_XML)
    box = int(np.flatnonzero(md["geom_type"] == mjcf.GEOM_BOX)[0])
    om = engine.Model(md)
    d = engine.Data(om)
    d.reset()
    cons = d.contacts()
    floor_box = [c for c in cons if c["geom1"] == 0 and c["geom2"] == box]
    assert len(floor_box) == 4, cons  # the four bottom corners, exactly touching at qpos0 (z = 0.2, half size 0.2)
    pos = np.array([c["pos"] for c in floor_box])
    # corner-index order of mjc_PlaneBox: bit 0 = x sign, bit 1 = y sign, bottom face (bit 2 clear)
    np.testing.assert_allclose(pos[:, :2], [[-0.2, -0.2], [0.2, -0.2], [-0.2, 0.2], [0.2, 0.2]], atol=1e-12)
    assert all(abs(c["dist"]) < 1e-12 for c in floor_box)
    # settle: 4 contacts share the weight, the box stays put
    d.step(600)
    z = d.qpos[9]
    assert 0.2 - 2e-3 < z < 0.2 + 1e-3 and np.abs(d.qvel[7:13]).max() < 1e-6  # rests inside the 1 mm contact margin
    assert len([c for c in d.contacts() if c["geom1"] == 0 and c["geom2"] == box]) == 4
    # tilted about x by 20 degrees and lifted so that only the lowest edge is within the margin: 2 corners
    d2 = engine.Data(om)
    d2.reset()
    ang = np.deg2rad(20.0)
    d2.qpos[10:14] = [np.cos(ang / 2), np.sin(ang / 2), 0, 0]
    R = np.array([[1, 0, 0], [0, np.cos(ang), -np.sin(ang)], [0, np.sin(ang), np.cos(ang)]])
    corners = (R @ (np.array([[sx, sy, sz] for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)]) * 0.2).T).T
    d2.qpos[9] = -corners[:, 2].min()
    d2.forward_position()
    fb = [c for c in d2.contacts() if c["geom1"] == 0 and c["geom2"] == box]
    assert len(fb) == 2 and all(abs(c["dist"]) < 1e-9 for c in fb)


def test_friction_cone_obeys_coulomb_at_the_stick_slip_threshold():
    """Analytic pin of the oracle's elliptic friction cone (no MuJoCo needed): a 1 kg box on the floor, pair friction 1, pushed
    horizontally at its base (force at the centre of mass + the torque that moves its line of action to the floor).  Below mu*m*g
    it must stay put (soft-constraint creep only), just above it must accelerate with (F - mu*m*g)/m."""
    from oracle import engine, mjcf
    md = mjcf.compile_mjcf(os.path.join(XMLS, "gripper_two_fingers.xml"))
    m = engine.Model(md)
    ob = md["body_names"].index("object")
    mass, g, h = md["body_mass"][ob], 9.81, 0.2
    assert abs(mass - 1.0) < 1e-12

    def push(mult, nsteps=100):
        d = engine.Data(m)
        d.reset()
        d.xfrc_applied[m.body_id("ee"), 2] = 0.438 * g   # gripper hovers (robot_env.py:64-65), never touches the box here
        d.step(600)                                        # the box settles on its four corner contacts
        assert d.ncon == 4 and abs(d.qvel[7:13]).max() < 1e-6
        F = mult * mass * g
        d.xfrc_applied[ob, 0] = F
        d.xfrc_applied[ob, 4] = -h * F
        v = []
        for _ in range(nsteps):
            d.step()
            v.append(d.qvel[7])
        return np.array(v), d
    for mult in (0.5, 0.95):
        v, d = push(mult)
        assert d.ncon == 4 and abs(v).max() < 1e-3, (mult, abs(v).max())      # free motion would reach 0.98 / 1.86 m/s
    v, d = push(1.05)
    accel = (v[-1] - v[19]) / (80 * 0.002)
    assert d.ncon >= 2 and abs(accel - 0.05 * g) < 0.05 * 0.05 * g + 0.02, accel   # 0.4905 m/s^2 expected; measured 0.498
