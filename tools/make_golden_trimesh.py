#!/usr/bin/env python
"""Golden vector for terrain.mesh_type "trimesh": the triangles the reference's own convert_heightfield_to_trimesh
(/root/reference/legged_gym/utils/terrain_utils.py:835-902, imported unchanged; its `trimesh` / `scipy` imports are satisfied by
an empty stand-in module) builds from a small random heightfield, and -- by brute force over exactly those triangles -- the
surface height and unit normal under random query points.  tests/test_oracle_physics.py pins the oracle's (and through it the
kernel's) collision surface against it.

    python tools/make_golden_trimesh.py [--out tests/golden/trimesh_surface.npz]
"""
import argparse
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "trimesh_surface.npz"))
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    for name in ("trimesh",):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["trimesh"].Trimesh = lambda *a, **k: None
    sys.path.insert(0, os.path.join(args.reference, "legged_gym", "utils"))
    import terrain_utils                                     # the reference module, by file (no legged_gym package import needed)
    rng = np.random.default_rng(7)
    rows, cols, hs, vs = 9, 11, 0.1, 0.005
    hf = rng.integers(-40, 40, (rows, cols)).astype(np.int16)
    vertices, triangles = terrain_utils.convert_heightfield_to_trimesh(hf, hs, vs)            # slope_threshold None, like terrain_utils.py:92
    xy = np.stack([rng.uniform(0.0, (rows - 1) * hs, 400), rng.uniform(0.0, (cols - 1) * hs, 400)], 1)
    h, nrm = np.full(len(xy), np.nan), np.zeros((len(xy), 3))
    margin = np.full(len(xy), -np.inf)
    for tri in triangles.astype(np.int64):
        a, b, c = vertices[tri].astype(np.float64)
        T = np.array([[b[0] - a[0], c[0] - a[0]], [b[1] - a[1], c[1] - a[1]]])
        lam = np.linalg.solve(T, (xy - a[:2]).T).T
        l0 = 1.0 - lam.sum(1)
        m = np.minimum(np.minimum(lam[:, 0], lam[:, 1]), l0)          # how far inside this triangle the point lies
        take = m > margin
        z = a[2] * l0 + b[2] * lam[:, 0] + c[2] * lam[:, 1]
        n = np.cross(b - a, c - a)
        n = n / np.linalg.norm(n) * np.sign(n[2])
        h[take], margin[take] = z[take], m[take]
        nrm[take] = n
    keep = margin > 1e-3                                       # drop points on an edge: either neighbour is a valid answer there
    np.savez_compressed(args.out, hf=hf, hscale=hs, vscale=vs, xy=xy[keep], height=h[keep], normal=nrm[keep],
                        vertices=vertices, triangles=triangles.astype(np.int32))
    print(f"wrote {args.out}: {int(keep.sum())} query points over {len(triangles)} reference triangles")


if __name__ == "__main__":
    main()
