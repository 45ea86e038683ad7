"""ORACLE (test infrastructure, not product code) — fp64 numpy ray caster of the RGB-D observation.

PARITY UNPINNED against MuJoCo's OpenGL renderer (physics.render, reference simulation/controller/sensor.py:64-73):
OpenGL pixels cannot be reproduced in the build container (no MuJoCo, no GL).  This file restates the rendering MODEL
the product implements — pinhole camera with MJCF `fovy` (degrees), depth = distance along the optical axis clipped to
[znear, zfar] x extent, flat-shaded triangles under OpenGL fixed-function lighting (emission + per light ambient +
diffuse + Blinn specular), MuJoCo's default head light, spot cut-off 45 deg / exponent 10, the builtin 2x2 checker
texture with texuniform repeat, the skybox gradient — with an independent algorithm (brute-force Moeller-Trumbore ray
casting in world space, fp64) so that tests/test_render.py can check the CUDA rasteriser against it.

Only tests/ may import this module.
"""
import numpy as np


def _quat2mat(q):
    w, x, y, z = q
    return np.array([[w*w + x*x - y*y - z*z, 2*(x*y - w*z), 2*(x*z + w*y)],
                     [2*(x*y + w*z), w*w - x*x + y*y - z*z, 2*(y*z - w*x)],
                     [2*(x*z - w*y), 2*(y*z + w*x), w*w - x*x - y*y + z*z]])


def _quat_mul(a, b):
    return np.array([a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3], a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2],
                     a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1], a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0]])


class Scene:
    """Static render data taken from a compiled model export (object with .get / .names / .sizes)."""

    def __init__(self, cm):
        sz = cm.sizes
        self.ngeom = sz["ngeom"]
        self.geom_type = cm.get("geom_type")
        self.geom_meshid = cm.get("geom_meshid")
        self.geom_rgba = cm.get("geom_rgba").reshape(-1, 4)
        self.geom_mat = cm.get("geom_matprop").reshape(-1, 4)  # emission, specular, shininess, textured
        self.geom_size = cm.get("geom_size").reshape(-1, 3)
        names = cm.names("mesh")
        self.tris = [cm.get("mesh_tri:" + n).reshape(-1, 3, 3) for n in names]
        self.cam_body = cm.get("cam_bodyid")
        self.cam_mode = cm.get("cam_mode")
        self.cam_target = cm.get("cam_target")
        self.cam_pos = cm.get("cam_pos").reshape(-1, 3)
        self.cam_quat = cm.get("cam_quat").reshape(-1, 4)
        self.cam_fovy = cm.get("cam_fovy")
        self.cam_names = cm.names("camera")
        self.light_directional = cm.get("light_directional")
        self.light_pos = cm.get("light_pos").reshape(-1, 3)
        self.light_dir = cm.get("light_dir").reshape(-1, 3)
        self.light_diffuse = cm.get("light_diffuse").reshape(-1, 3)
        self.light_ambient = cm.get("light_ambient").reshape(-1, 3)
        self.light_specular = cm.get("light_specular").reshape(-1, 3)
        self.tex1, self.tex2, self.texrepeat = cm.get("tex_rgb1"), cm.get("tex_rgb2"), cm.get("texrepeat")
        self.sky1, self.sky2 = cm.get("sky_rgb1"), cm.get("sky_rgb2")
        ext = float(cm.get("extent")[0])
        self.znear, self.zfar = float(cm.get("znear")[0]) * ext, float(cm.get("zfar")[0]) * ext


def camera_pose(scene, cam, xpos, xquat, subtree_com):
    """engine_core_smooth.c : mj_camlight — camera position and 3x3 frame (columns x right, y up, z backwards)."""
    b = scene.cam_body[cam]
    R = _quat2mat(xquat[b])
    pos = xpos[b] + R @ scene.cam_pos[cam]
    if scene.cam_mode[cam] == 1 and scene.cam_target[cam] >= 0:
        z = pos - subtree_com[scene.cam_target[cam]]
        z /= np.linalg.norm(z)
        x = np.cross([0, 0, 1.0], z)
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        y /= np.linalg.norm(y)
        mat = np.stack([x, y, z], 1)
    else:
        q = _quat_mul(xquat[b], scene.cam_quat[cam])
        mat = _quat2mat(q / np.linalg.norm(q))
    return pos, mat


def render(scene, geom_xpos, geom_xmat, cam_pos, cam_mat, fovy_deg, width, height):
    """Returns (rgb uint8 [h,w,3], depth float64 [h,w] metres, hit geom id [h,w] (-1 = sky))."""
    H, W = height, width
    ify = np.tan(0.5 * np.deg2rad(fovy_deg))
    ifx = ify * W / H
    jj, ii = np.meshgrid(np.arange(W), np.arange(H))
    dx = ((2 * (jj + 0.5)) / W - 1) * ifx
    dy = (1 - (2 * (ii + 0.5)) / H) * ify
    dcam = np.stack([dx, dy, -np.ones_like(dx)], -1).reshape(-1, 3)
    dw = dcam @ cam_mat.T                       # world ray directions, optical-axis component 1
    P = dw.shape[0]
    best_t = np.full(P, scene.zfar)
    best_g = np.full(P, -1)
    best_n = np.zeros((P, 3))
    for g in range(scene.ngeom):
        if scene.geom_type[g] == 0:
            continue
        tri = scene.tris[scene.geom_meshid[g]] @ geom_xmat[g].T + geom_xpos[g]   # world-space triangles
        a, e1, e2 = tri[:, 0], tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
        nrm = np.cross(e1, e2)
        for lo in range(0, len(tri), 1024):
            A, E1, E2, N = a[lo:lo + 1024], e1[lo:lo + 1024], e2[lo:lo + 1024], nrm[lo:lo + 1024]
            # Moeller-Trumbore, rays x triangles
            pv = np.cross(dw[:, None, :], E2[None])                 # [P,T,3]
            det = (pv * E1[None]).sum(-1)
            with np.errstate(divide="ignore", invalid="ignore"):
                inv = 1.0 / det
                tv = cam_pos[None, None, :] - A[None]
                u = (tv * pv).sum(-1) * inv
                qv = np.cross(tv, E1[None])
                v = (qv * dw[:, None, :]).sum(-1) * inv
                t = (qv * E2[None]).sum(-1) * inv
            ok = (det != 0) & (u >= 0) & (v >= 0) & (u + v <= 1) & (t >= scene.znear) & (t <= scene.zfar)
            t = np.where(ok, t, np.inf)
            k = t.argmin(1)
            tm = t[np.arange(P), k]
            upd = tm < best_t
            best_t[upd], best_g[upd], best_n[upd] = tm[upd], g, N[k[upd]]
    # floor plane
    pg = int(np.where(scene.geom_type == 0)[0][0]) if (scene.geom_type == 0).any() else -1
    if pg >= 0:
        pn = geom_xmat[pg][:, 2]
        denom = dw @ pn
        with np.errstate(divide="ignore", invalid="ignore"):
            t = ((geom_xpos[pg] - cam_pos) @ pn) / denom
        hp = cam_pos + t[:, None] * dw - geom_xpos[pg]
        u, v = hp @ geom_xmat[pg][:, 0], hp @ geom_xmat[pg][:, 1]
        sx, sy = scene.geom_size[pg][0], scene.geom_size[pg][1]
        ok = (denom != 0) & (t >= scene.znear) & (t < best_t) & ((sx <= 0) | (np.abs(u) <= sx)) & ((sy <= 0) | (np.abs(v) <= sy))
        best_t[ok], best_g[ok] = t[ok], pg
        best_n[ok] = pn
    rgb = np.zeros((P, 3))
    sky = best_g < 0
    mx = np.abs(dw).max(1)
    f = 0.5 * (dw[:, 2] / mx + 1)
    rgb[sky] = scene.sky2 + f[sky, None] * (scene.sky1 - scene.sky2)
    fwd = -cam_mat[:, 2]
    hit = ~sky
    p = cam_pos + best_t[:, None] * dw
    for g in np.unique(best_g[hit]):
        sel = best_g == g
        n = best_n[sel] / np.linalg.norm(best_n[sel], axis=1, keepdims=True)
        vdir = cam_pos - p[sel]
        vdir /= np.linalg.norm(vdir, axis=1, keepdims=True)
        n = np.where(((n * vdir).sum(1) < 0)[:, None], -n, n)
        rgba, (em, spec, shin, textured) = scene.geom_rgba[g][:3], scene.geom_mat[g]
        shin = max(shin * 128, 1e-3)
        col = np.tile(em * rgba, (sel.sum(), 1))
        nl = len(scene.light_directional)
        for l in range(nl + 1):
            att = np.ones(sel.sum())
            if l == nl:
                L = np.tile(-fwd, (sel.sum(), 1))
                amb, dif, spc = np.full(3, 0.1), np.full(3, 0.4), np.full(3, 0.5)
            else:
                amb, dif, spc = scene.light_ambient[l], scene.light_diffuse[l], scene.light_specular[l]
                ld = scene.light_dir[l] / np.linalg.norm(scene.light_dir[l])
                if scene.light_directional[l]:
                    L = np.tile(-ld, (sel.sum(), 1))
                else:
                    L = scene.light_pos[l] - p[sel]
                    L /= np.linalg.norm(L, axis=1, keepdims=True)
                    cs = -(L @ ld)
                    att = np.where(cs < 0.70710678, 0.0, np.abs(cs) ** 10)
            ndl = np.maximum((n * L).sum(1), 0)
            Hh = L + vdir
            Hh /= np.linalg.norm(Hh, axis=1, keepdims=True)
            ndh = np.maximum((n * Hh).sum(1), 0)
            sp = np.where(ndl > 0, spec * ndh ** shin, 0.0)
            col += amb * rgba + att[:, None] * (dif * rgba * ndl[:, None] + spc * sp[:, None])
        tex = np.ones((sel.sum(), 3))
        if textured and g == pg:
            hp = p[sel] - geom_xpos[pg]
            u, v = hp @ geom_xmat[pg][:, 0], hp @ geom_xmat[pg][:, 1]
            odd = ((np.floor(u * scene.texrepeat[0] * 2) + np.floor(v * scene.texrepeat[1] * 2)).astype(np.int64) & 1).astype(bool)
            tex = np.where(odd[:, None], scene.tex2, scene.tex1)
        rgb[sel] = np.minimum(col, 1.0) * tex
    depth = np.where(sky, scene.zfar, best_t)
    rgb8 = (np.clip(rgb, 0, 1) * 255 + 0.5).astype(np.uint8)
    return rgb8.reshape(H, W, 3), depth.reshape(H, W), best_g.reshape(H, W)


def observation(rgb8, depth, grasp, pheromone, full_observation=True):
    """RobotEnv.get_observation (robot_env.py:275-293): transform_depth, pad channel, dstack, uint8, CHW."""
    from .envmath import transform_depth
    pad = np.zeros(depth.shape)
    pad[0, 0], pad[0, 1] = grasp, pheromone
    chans = [rgb8.astype(np.float64)]
    if full_observation:
        chans.append(transform_depth(depth.astype(np.float32))[..., None].astype(np.float64))
    chans.append(pad[..., None])
    return np.dstack(chans).astype(np.uint8).transpose(2, 0, 1)
