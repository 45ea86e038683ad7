"""ORACLE (test infrastructure, not product code) — MJCF subset -> flat model arrays, in numpy fp64.

PARITY UNPINNED for everything that restates MuJoCo's model compiler: MuJoCo (`mujoco` 2.2.x via
dm-control==1.0.3.post1, reference setup.py:7) is not importable in the build container and the
reference holds no golden model constants.  This file restates the published
behaviour of MuJoCo's compiler for exactly the MJCF subset the reference scenes use
(reference xmls/acorn_env.xml:1-113 and siblings):

  * <compiler angle meshdir inertiafromgeom>, <option>, <default> with nested classes,
    <asset><mesh|material>, <worldbody> body/geom/joint/freejoint/inertial/camera/light, <actuator><motor>
  * mesh processing as MuJoCo 2.2.x user_mesh.cc (legacy inertia rule: pyramids from the
    area-weighted face-centroid centre, absolute volumes), recentring to the CoM / principal axes,
    convex hull (qhull, option Qt — scipy wraps the same library) and the hull vertex graph
  * inertiafromgeom="true": bodies with geoms get mass/CoM/inertia from their geoms, a body without
    geoms keeps its explicit <inertial> (the `ee` body)
  * body weld ids and the parent/weld collision filter of engine_collision_driver.c -> geom pair list
    in MuJoCo's contact order (body pair ascending, plane first inside a pair)

It is deliberately independent of the product's C++ compiler (csrc/mjcf_compiler.cpp), which has
its own XML reader, STL reader, quickhull and eigen-solver; tests/ compare the two.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

GEOM_PLANE, GEOM_BOX, GEOM_MESH = 0, 6, 7          # mjtGeom values
JNT_FREE, JNT_SLIDE, JNT_HINGE = 0, 2, 3           # mjtJoint values


# ----------------------------------------------------------------------------- small math
def quat_mul(a, b):
    return np.array([a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3],
                     a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2],
                     a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1],
                     a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0]])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([[w*w + x*x - y*y - z*z, 2*(x*y - w*z), 2*(x*z + w*y)],
                     [2*(x*y + w*z), w*w - x*x + y*y - z*z, 2*(y*z - w*x)],
                     [2*(x*z - w*y), 2*(y*z + w*x), w*w - x*x - y*y + z*z]])


def mat_to_quat(R):
    """Rotation matrix -> unit quaternion (w,x,y,z), w >= 0."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    if q[0] < 0:
        q = -q
    return q / np.linalg.norm(q)


def euler_to_quat(e):
    """MuJoCo default eulerseq 'xyz': intrinsic rotations about x, then y, then z."""
    q = np.array([1.0, 0, 0, 0])
    for ax, ang in enumerate(e):
        r = np.zeros(4)
        r[0] = np.cos(ang / 2)
        r[1 + ax] = np.sin(ang / 2)
        q = quat_mul(q, r)
    return q


def _floats(s):
    return np.array([float(x) for x in s.split()], dtype=np.float64)


# ----------------------------------------------------------------------------- meshes
def load_stl(path):
    """Binary STL -> (F,3,3) float64 triangle corner array (stored as float32 in the file)."""
    with open(path, "rb") as f:
        buf = f.read()
    n = struct.unpack_from("<I", buf, 80)[0]
    rec = np.frombuffer(buf, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]),
                        count=n, offset=84)
    return rec["v"].astype(np.float64)


def process_mesh(tri):
    """MuJoCo 2.2.x mesh processing (user_mesh.cc mjCMesh::Process, legacy inertia) [ext, from memory].

    Returns dict(pos, quat, volume, inertia_unit[3] (principal, unit density), verts (all unique
    vertices in the recentred principal frame), hull (indices into verts), faces (hull triangles as
    hull-local ids), graph adjacency lists).
    """
    a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
    nrm = np.cross(b - a, c - a)
    area2 = np.linalg.norm(nrm, axis=1)
    keep = area2 > 1e-30
    a, b, c, nrm, area2 = a[keep], b[keep], c[keep], nrm[keep], area2[keep]
    area = 0.5 * area2
    nrm = nrm / area2[:, None]
    cen = (a + b + c) / 3.0
    facecen = (cen * area[:, None]).sum(0) / area.sum()
    # pass 1: CoM from pyramids (apex = facecen), absolute volumes
    vol = np.abs(((a - facecen) * nrm).sum(1)) * area / 3.0
    pyr_cen = 0.75 * cen + 0.25 * facecen
    volume = vol.sum()
    com = (pyr_cen * vol[:, None]).sum(0) / volume
    # pass 2: second moments about the CoM from pyramids with apex = CoM
    D, E, F = a - com, b - com, c - com
    vol2 = np.abs((D * nrm).sum(1)) * area / 3.0
    P = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            P[i, j] = (vol2 / 20.0 * (2 * (D[:, i]*D[:, j] + E[:, i]*E[:, j] + F[:, i]*F[:, j])
                                      + D[:, i]*E[:, j] + D[:, j]*E[:, i]
                                      + D[:, i]*F[:, j] + D[:, j]*F[:, i]
                                      + E[:, i]*F[:, j] + E[:, j]*F[:, i])).sum()
    volume2 = vol2.sum()
    inert = np.trace(P) * np.eye(3) - P
    w, V = np.linalg.eigh(inert)
    order = np.argsort(-w)              # MuJoCo sorts principal moments in decreasing order
    w, V = w[order], V[:, order]
    if np.linalg.det(V) < 0:
        V[:, 2] = -V[:, 2]
    quat = mat_to_quat(V)
    V = quat_to_mat(quat)
    # unique vertices, recentred and rotated into the principal frame
    allv = tri.reshape(-1, 3)
    uniq = np.unique(allv, axis=0)
    verts = (uniq - com) @ V
    # convex hull with the same library MuJoCo links (qhull, triangulated output)
    from scipy.spatial import ConvexHull
    hull = ConvexHull(verts, qhull_options="Qt")
    hv = np.array(sorted(set(hull.simplices.reshape(-1).tolist())), dtype=np.int64)
    remap = {int(g): l for l, g in enumerate(hv)}
    hverts = verts[hv]
    faces = np.array([[remap[int(i)] for i in s] for s in hull.simplices], dtype=np.int32)
    # orient faces outward
    ctr = hverts.mean(0)
    for f in faces:
        n = np.cross(hverts[f[1]] - hverts[f[0]], hverts[f[2]] - hverts[f[0]])
        if np.dot(n, hverts[f[0]] - ctr) < 0:
            f[1], f[2] = f[2], f[1]
    adj = [set() for _ in range(len(hv))]
    for f in faces:
        for i in range(3):
            adj[f[i]].add(int(f[(i + 1) % 3]))
            adj[f[(i + 1) % 3]].add(int(f[i]))
    adj = [sorted(s) for s in adj]
    return dict(pos=com, quat=quat, volume=volume2, inertia_unit=w, verts=verts,
                hull_verts=hverts, hull_faces=faces, hull_adj=adj,
                tri_local=((tri - com) @ V))


def box_as_mesh(size):
    """A box geom expressed as an 8-vertex convex mesh (deviation)."""
    s = np.asarray(size, dtype=np.float64)
    hverts = np.array([[sx * s[0], sy * s[1], sz * s[2]] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])
    from scipy.spatial import ConvexHull
    hull = ConvexHull(hverts, qhull_options="Qt")
    faces = hull.simplices.astype(np.int32).copy()
    for f in faces:
        n = np.cross(hverts[f[1]] - hverts[f[0]], hverts[f[2]] - hverts[f[0]])
        if np.dot(n, hverts[f[0]]) < 0:
            f[1], f[2] = f[2], f[1]
    adj = [set() for _ in range(8)]
    for f in faces:
        for i in range(3):
            adj[f[i]].add(int(f[(i + 1) % 3]))
            adj[f[(i + 1) % 3]].add(int(f[i]))
    vol = 8 * s.prod()
    inertia = vol / 3.0 * np.array([s[1]**2 + s[2]**2, s[0]**2 + s[2]**2, s[0]**2 + s[1]**2])
    tri = np.array([[hverts[f[0]], hverts[f[1]], hverts[f[2]]] for f in faces])
    return dict(pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), volume=vol, inertia_unit=inertia,
                verts=hverts, hull_verts=hverts, hull_faces=faces, hull_adj=[sorted(a) for a in adj],
                tri_local=tri)


# ----------------------------------------------------------------------------- defaults
class _Defaults:
    def __init__(self, root):
        self.cls = {}
        top = root.find("default")
        if top is not None:
            self._walk(top, "main", {})

    def _walk(self, node, name, inherited):
        cur = {k: dict(v) for k, v in inherited.items()}
        for ch in node:
            if ch.tag != "default":
                cur.setdefault(ch.tag, {}).update(ch.attrib)
        self.cls[name] = cur
        for ch in node:
            if ch.tag == "default":
                self._walk(ch, ch.get("class"), cur)

    def get(self, tag, elem, childclass):
        cname = elem.get("class") or childclass or "main"
        out = dict(self.cls.get(cname, {}).get(tag, {}))
        out.update(elem.attrib)
        return out


# ----------------------------------------------------------------------------- compile
def compile_mjcf(xml_path):
    """Parse + compile; returns a dict of numpy arrays named after mjModel fields."""
    root = ET.parse(xml_path).getroot()
    comp = root.find("compiler").attrib if root.find("compiler") is not None else {}
    assert comp.get("angle", "degree") == "radian", "only angle=radian scenes are supported"
    meshdir = os.path.join(os.path.dirname(os.path.abspath(xml_path)), comp.get("meshdir", ""))
    inertiafromgeom = comp.get("inertiafromgeom", "auto")
    opt = root.find("option").attrib if root.find("option") is not None else {}
    dfl = _Defaults(root)

    m = dict()
    m["timestep"] = float(opt.get("timestep", 0.002))
    m["gravity"] = _floats(opt.get("gravity", "0 0 -9.81"))
    m["impratio"] = float(opt.get("impratio", 1))
    m["tolerance"] = float(opt.get("tolerance", 1e-8))
    m["iterations"] = int(opt.get("iterations", 100))
    m["cone_elliptic"] = 1 if opt.get("cone", "pyramidal") == "elliptic" else 0
    vis = root.find("visual")
    m["znear"] = 0.01
    m["zfar"] = 50.0
    if vis is not None and vis.find("map") is not None:
        m["znear"] = float(vis.find("map").get("znear", 0.01))
        m["zfar"] = float(vis.find("map").get("zfar", 50.0))

    # assets
    meshes, mesh_names, materials = [], [], {}
    asset = root.find("asset")
    for e in (asset if asset is not None else []):
        if e.tag == "mesh":
            fn = e.get("file")
            mesh_names.append(e.get("name") or os.path.splitext(os.path.basename(fn))[0])
            meshes.append(process_mesh(load_stl(os.path.join(meshdir, fn))))
        elif e.tag == "material":
            materials[e.get("name")] = dict(e.attrib)

    bodies = [dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]),
                   inertial=None, geoms=[], joints=[])]
    geoms, joints, cameras, lights = [], [], [], []

    def frame_of(attr):
        pos = _floats(attr.get("pos", "0 0 0"))
        if "quat" in attr:
            q = _floats(attr["quat"])
            q = q / np.linalg.norm(q)
        elif "euler" in attr:
            q = euler_to_quat(_floats(attr["euler"]))
        else:
            q = np.array([1.0, 0, 0, 0])
        return pos, q

    def add_geom(e, bid, childclass):
        a = dfl.get("geom", e, childclass)
        gtype = {"plane": GEOM_PLANE, "mesh": GEOM_MESH, "box": GEOM_BOX}[a.get("type", "sphere")]
        pos, quat = frame_of(a)
        g = dict(name=a.get("name", ""), body=bid, type=gtype, condim=int(a.get("condim", 3)),
                 friction=np.resize(_floats(a.get("friction", "1 0.005 0.0001")), 3),
                 margin=float(a.get("margin", 0)), gap=float(a.get("gap", 0)),
                 solref=_floats(a.get("solref", "0.02 1")), solimp=_floats(a.get("solimp", "0.9 0.95 0.001 0.5 2")),
                 contype=int(a.get("contype", 1)), conaffinity=int(a.get("conaffinity", 1)),
                 size=np.zeros(3), mesh=-1, mass=0.0, inertia=np.zeros(3))
        fr = _floats(a.get("friction", "1 0.005 0.0001"))
        base = np.array([1, 0.005, 0.0001])
        base[:len(fr)] = fr
        g["friction"] = base
        rgba = np.array([0.5, 0.5, 0.5, 1.0])
        mat = materials.get(a.get("material", ""), None)
        if mat is not None:          # effective render colour: a material replaces the geom default; <material rgba> defaults to "1 1 1 1"
            rgba = _floats(mat.get("rgba", "1 1 1 1"))
        if "rgba" in a:
            rgba = _floats(a["rgba"])
        g["rgba"] = rgba
        g["material"] = a.get("material", "")
        if gtype == GEOM_PLANE:
            g["size"] = _floats(a["size"])
            g["pos"], g["quat"] = pos, quat
        else:
            if gtype == GEOM_MESH:
                g["mesh"] = mesh_names.index(a["mesh"])
                mp = meshes[g["mesh"]]
            else:
                g["size"] = _floats(a["size"])
                meshes.append(box_as_mesh(g["size"]))
                mesh_names.append("__box_%d" % len(meshes))
                g["mesh"] = len(meshes) - 1
                mp = meshes[-1]
            # compose the user frame with the mesh's CoM / principal frame
            g["pos"] = pos + quat_to_mat(quat) @ mp["pos"]
            g["quat"] = quat_mul(quat, mp["quat"])
            density = float(a.get("density", 1000.0))
            mass = density * mp["volume"]
            inertia = density * mp["inertia_unit"]
            if "mass" in a:
                s = float(a["mass"]) / mass
                mass, inertia = mass * s, inertia * s
            g["mass"], g["inertia"] = mass, inertia
        geoms.append(g)
        bodies[bid]["geoms"].append(len(geoms) - 1)

    def walk(elem, bid, childclass):
        for e in elem:
            if e.tag == "geom":
                add_geom(e, bid, childclass)
            elif e.tag in ("joint", "freejoint"):
                if e.tag == "freejoint":
                    a = dict(type="free")
                else:
                    a = dfl.get("joint", e, childclass)
                jt = {"free": JNT_FREE, "slide": JNT_SLIDE, "hinge": JNT_HINGE}[a.get("type", "hinge")]
                joints.append(dict(name=a.get("name", ""), body=bid, type=jt,
                                   pos=_floats(a.get("pos", "0 0 0")), axis=_floats(a.get("axis", "0 0 1")),
                                   limited=a.get("limited", "false") == "true",
                                   range=_floats(a.get("range", "0 0")),
                                   armature=float(a.get("armature", 0)) if jt != JNT_FREE else 0.0,
                                   damping=float(a.get("damping", 0)) if jt != JNT_FREE else 0.0))
                bodies[bid]["joints"].append(len(joints) - 1)
            elif e.tag == "inertial":
                pos, quat = frame_of(e.attrib)
                bodies[bid]["inertial"] = dict(pos=pos, quat=quat, mass=float(e.get("mass")),
                                               inertia=_floats(e.get("diaginertia")))
            elif e.tag == "camera":
                pos, quat = frame_of(e.attrib)
                cameras.append(dict(name=e.get("name", ""), body=bid, pos=pos, quat=quat,
                                    fovy=float(e.get("fovy", 45)), mode=e.get("mode", "fixed"),
                                    target=e.get("target", "")))
            elif e.tag == "light":
                a = dfl.get("light", e, childclass)
                lights.append(dict(body=bid, pos=_floats(a.get("pos", "0 0 0")), dir=_floats(a.get("dir", "0 0 -1")),
                                   directional=a.get("directional", "false") == "true",
                                   diffuse=_floats(a.get("diffuse", "0.7 0.7 0.7")),
                                   ambient=_floats(a.get("ambient", "0 0 0")),
                                   specular=_floats(a.get("specular", "0.3 0.3 0.3"))))
            elif e.tag == "body":
                pos, quat = frame_of(e.attrib)
                bodies.append(dict(name=e.get("name", ""), parent=bid, pos=pos, quat=quat, inertial=None,
                                   geoms=[], joints=[]))
                walk(e, len(bodies) - 1, e.get("childclass") or childclass)

    walk(root.find("worldbody"), 0, None)

    nbody, njnt, ngeom = len(bodies), len(joints), len(geoms)
    m["nbody"], m["njnt"], m["ngeom"], m["nmesh"] = nbody, njnt, ngeom, len(meshes)
    m["body_names"] = [b["name"] for b in bodies]
    m["geom_names"] = [g["name"] for g in geoms]
    m["body_parentid"] = np.array([b["parent"] for b in bodies], dtype=np.int32)
    m["body_pos"] = np.array([b["pos"] for b in bodies])
    m["body_quat"] = np.array([b["quat"] for b in bodies])

    # inertial properties
    bmass, bipos, biquat, binertia = np.zeros(nbody), np.zeros((nbody, 3)), np.tile([1.0, 0, 0, 0], (nbody, 1)), np.zeros((nbody, 3))
    for i, b in enumerate(bodies):
        solid = [geoms[g] for g in b["geoms"] if geoms[g]["type"] != GEOM_PLANE]
        use_geoms = (inertiafromgeom == "true" and solid) or (inertiafromgeom == "auto" and b["inertial"] is None and solid)
        if i == 0:
            continue
        if use_geoms:
            mass = sum(g["mass"] for g in solid)
            com = sum(g["mass"] * g["pos"] for g in solid) / mass
            I = np.zeros((3, 3))
            for g in solid:
                R = quat_to_mat(g["quat"])
                d = g["pos"] - com
                I += R @ np.diag(g["inertia"]) @ R.T + g["mass"] * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
            w, V = np.linalg.eigh(I)
            order = np.argsort(-w)
            w, V = w[order], V[:, order]
            if np.linalg.det(V) < 0:
                V[:, 2] = -V[:, 2]
            bmass[i], bipos[i], biquat[i], binertia[i] = mass, com, mat_to_quat(V), w
        elif b["inertial"] is not None:
            it = b["inertial"]
            bmass[i], bipos[i], biquat[i], binertia[i] = it["mass"], it["pos"], it["quat"], it["inertia"]
    m["body_mass"], m["body_ipos"], m["body_iquat"], m["body_inertia"] = bmass, bipos, biquat, binertia

    # joints / dofs / qpos0
    jnt_qposadr, jnt_dofadr, qpos0 = [], [], []
    dof_bodyid, dof_jntid, dof_parentid, dof_armature, dof_damping = [], [], [], [], []
    body_dofadr, body_dofnum = np.full(nbody, -1, np.int32), np.zeros(nbody, np.int32)
    body_jntadr, body_jntnum = np.full(nbody, -1, np.int32), np.zeros(nbody, np.int32)
    last_dof_of_body = np.full(nbody, -1, np.int32)
    for bi, b in enumerate(bodies):
        # last dof of the nearest ancestor that has dofs
        p = b["parent"]
        prev = last_dof_of_body[p] if bi > 0 else -1
        for j in b["joints"]:
            jt = joints[j]
            if body_jntadr[bi] < 0:
                body_jntadr[bi] = j
            body_jntnum[bi] += 1
            jnt_qposadr.append(len(qpos0))
            jnt_dofadr.append(len(dof_bodyid))
            if body_dofadr[bi] < 0:
                body_dofadr[bi] = len(dof_bodyid)
            if jt["type"] == JNT_FREE:
                R0 = b["quat"]
                qpos0 += list(b["pos"]) + list(R0)
                nd = 6
            else:
                qpos0.append(0.0)
                nd = 1
            for _ in range(nd):
                dof_bodyid.append(bi)
                dof_jntid.append(j)
                dof_parentid.append(prev)
                prev = len(dof_bodyid) - 1
                dof_armature.append(jt["armature"])
                dof_damping.append(jt["damping"])
            body_dofnum[bi] += nd
        last_dof_of_body[bi] = prev
    m["nq"], m["nv"] = len(qpos0), len(dof_bodyid)
    m["qpos0"] = np.array(qpos0)
    m["jnt_type"] = np.array([j["type"] for j in joints], np.int32)
    m["jnt_bodyid"] = np.array([j["body"] for j in joints], np.int32)
    m["jnt_qposadr"] = np.array(jnt_qposadr, np.int32)
    m["jnt_dofadr"] = np.array(jnt_dofadr, np.int32)
    m["jnt_pos"] = np.array([j["pos"] for j in joints])
    m["jnt_axis"] = np.array([j["axis"] / np.linalg.norm(j["axis"]) for j in joints])
    m["jnt_limited"] = np.array([int(j["limited"]) for j in joints], np.int32)
    m["jnt_range"] = np.array([j["range"] for j in joints])
    m["jnt_names"] = [j["name"] for j in joints]
    m["body_jntadr"], m["body_jntnum"] = body_jntadr, body_jntnum
    m["body_dofadr"], m["body_dofnum"] = body_dofadr, body_dofnum
    m["dof_bodyid"] = np.array(dof_bodyid, np.int32)
    m["dof_jntid"] = np.array(dof_jntid, np.int32)
    m["dof_parentid"] = np.array(dof_parentid, np.int32)
    m["dof_armature"] = np.array(dof_armature)
    m["dof_damping"] = np.array(dof_damping)
    m["jnt_solref"] = np.array([0.02, 1.0])
    m["jnt_solimp"] = np.array([0.9, 0.95, 0.001, 0.5, 2.0])

    # weld ids (bodies without joints are welded to their parent's weld body)
    weld = np.zeros(nbody, np.int32)
    for i in range(1, nbody):
        weld[i] = i if body_jntnum[i] > 0 else weld[bodies[i]["parent"]]
    m["body_weldid"] = weld

    # geoms
    m["geom_type"] = np.array([g["type"] for g in geoms], np.int32)
    m["geom_bodyid"] = np.array([g["body"] for g in geoms], np.int32)
    m["geom_meshid"] = np.array([g["mesh"] for g in geoms], np.int32)
    m["geom_condim"] = np.array([g["condim"] for g in geoms], np.int32)
    m["geom_pos"] = np.array([g["pos"] for g in geoms])
    m["geom_quat"] = np.array([g["quat"] for g in geoms])
    m["geom_friction"] = np.array([g["friction"] for g in geoms])
    m["geom_margin"] = np.array([g["margin"] for g in geoms])
    m["geom_gap"] = np.array([g["gap"] for g in geoms])
    m["geom_solref"] = np.array([g["solref"] for g in geoms])
    m["geom_solimp"] = np.array([g["solimp"] for g in geoms])
    m["geom_rgba"] = np.array([g["rgba"] for g in geoms])
    m["geom_size"] = np.array([g["size"] for g in geoms])
    m["geom_mass"] = np.array([g["mass"] for g in geoms])
    m["geom_rbound"] = np.array([0.0 if g["type"] == GEOM_PLANE else
                                 np.linalg.norm(meshes[g["mesh"]]["hull_verts"], axis=1).max() for g in geoms])
    m["meshes"] = meshes
    m["mesh_names"] = mesh_names
    m["materials"] = materials

    # collision pairs after the parent/weld filter, in MuJoCo contact order
    pairs = []
    for g1 in range(ngeom):
        for g2 in range(g1 + 1, ngeom):
            b1, b2 = geoms[g1]["body"], geoms[g2]["body"]
            if not ((geoms[g1]["contype"] & geoms[g2]["conaffinity"]) or (geoms[g2]["contype"] & geoms[g1]["conaffinity"])):
                continue
            w1, w2 = weld[b1], weld[b2]
            if w1 == w2:
                continue
            wp1 = weld[bodies[w1]["parent"]] if w1 > 0 else 0
            wp2 = weld[bodies[w2]["parent"]] if w2 > 0 else 0
            if w1 != 0 and w2 != 0 and (w1 == wp2 or w2 == wp1):
                continue
            a, b = (g1, g2) if geoms[g1]["type"] <= geoms[g2]["type"] else (g2, g1)
            pairs.append((min(b1, b2), max(b1, b2), a, b))
    pairs.sort(key=lambda p: (p[0], p[1]))
    m["pair_geom1"] = np.array([p[2] for p in pairs], np.int32)
    m["pair_geom2"] = np.array([p[3] for p in pairs], np.int32)

    # actuators
    act = root.find("actuator")
    names = m["jnt_names"]
    adof, agear, arange = [], [], []
    for e in (act if act is not None else []):
        a = dfl.get("motor", e, None)
        j = names.index(a["joint"])
        adof.append(jnt_dofadr[j])
        agear.append(_floats(a.get("gear", "1"))[0])
        arange.append(_floats(a.get("ctrlrange", "0 0")) if a.get("ctrllimited", "false") == "true" else np.array([-1e30, 1e30]))
    m["nu"] = len(adof)
    m["act_dofid"] = np.array(adof, np.int32)
    m["act_gear"] = np.array(agear)
    m["act_ctrlrange"] = np.array(arange).reshape(-1, 2)
    m["cameras"] = cameras
    m["lights"] = lights
    return m
