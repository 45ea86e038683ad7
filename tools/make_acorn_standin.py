#!/usr/bin/env python
"""Write assets/xmls/meshes/acorn_standin.stl — a SYNTHETIC stand-in for the reference's acorn.stl.

The reference tree does not contain xmls/meshes/acorn.stl (.MISSING_LARGE_BLOBS:30), yet
acorn_env.xml (xmls/acorn_env.xml:29) is the headline scene.  This script generates a closed,
outward-oriented surface of revolution shaped like an acorn (nut + wider cap + stalk), 0.62 m tall
and 0.5 m across — the same scale as the three object meshes that are present (extents 0.5–0.73 m)
— standing on z = 0 and centred on the z axis, which suits the acorn scene's small geom offset
(-0.025, 0.05, 0).  Every number measured on "acorn_env" in this repository is measured on this
stand-in and is labelled so.
"""
import os
import struct
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                   "mujoco_rl_manipulate_unknown_objects_b200", "assets", "xmls", "meshes",
                   "acorn_standin.stl")

# (z, radius) profile from the tip (bottom) to the stalk (top)
PROFILE = [(0.00, 0.000), (0.02, 0.070), (0.07, 0.135), (0.14, 0.185), (0.22, 0.215),
           (0.30, 0.225), (0.36, 0.220), (0.38, 0.250), (0.44, 0.245), (0.50, 0.200),
           (0.54, 0.120), (0.56, 0.040), (0.60, 0.030), (0.62, 0.000)]
NSEG = 28


def main():
    ang = np.linspace(0.0, 2 * np.pi, NSEG, endpoint=False)
    rings = [np.stack([r * np.cos(ang), r * np.sin(ang), np.full(NSEG, z)], 1) for z, r in PROFILE[1:-1]]
    bottom = np.array([0.0, 0.0, PROFILE[0][0]])
    top = np.array([0.0, 0.0, PROFILE[-1][0]])
    tris = []
    for i in range(NSEG):
        j = (i + 1) % NSEG
        tris.append((bottom, rings[0][j], rings[0][i]))
        for a, b in zip(rings[:-1], rings[1:]):
            tris.append((a[i], a[j], b[j]))
            tris.append((a[i], b[j], b[i]))
        tris.append((top, rings[-1][i], rings[-1][j]))
    with open(OUT, "wb") as f:
        f.write(b"acorn_standin: synthetic, see tools/make_acorn_standin.py".ljust(80, b" "))
        f.write(struct.pack("<I", len(tris)))
        for a, b, c in tris:
            n = np.cross(b - a, c - a)
            n = n / (np.linalg.norm(n) + 1e-30)
            f.write(struct.pack("<12fH", *n, *a, *b, *c, 0))
    print("wrote", OUT, len(tris), "triangles")


if __name__ == "__main__":
    main()
