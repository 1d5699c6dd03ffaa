#!/usr/bin/env python3
"""Seeded synthetic triangle meshes for the _trianglegrid configs (BASELINE.json configs 4 and 5).

The reference imposes constraints a mesh must meet or it renders as nothing:
  * all coordinates > 0          (the host's running bbox maximum starts at FLT_MIN),
  * 2*area >= ~0.04 per triangle (TriangleIntersect culls |det| < 0.01 absolutely),
  * no cell with more than 62 references at the clamped 128^3 resolution,
  * written in triangles.txt format with "%f" and no trailing newline.
Coordinates are generated as integer micro-units / 1e6, so float32(atof("%f" % x)) == float32(x): arrays
handed straight to pt_set_scene are bit-identical to what the parsers read back from the text file.

    python scenes/gen_mesh.py <ntriangles> <out_dir> [--seed N] [--box L]    # writes a full grid scene dir
"""
import os
import sys

import numpy as np


def soup(n, seed=20261018, box_lo=0.5, box_size=60.0, edge=(0.25, 0.40)):
    """n randomly oriented near-equilateral triangles, centres uniform in [box_lo, box_lo+box_size]^3.
    Returns float32 (n, 12): v0.xyzw v1.xyzw v2.xyzw with w = 0."""
    rng = np.random.default_rng(seed)
    margin = edge[1]
    c = rng.uniform(box_lo + margin, box_lo + box_size - margin, (n, 3))
    a = rng.uniform(edge[0], edge[1], n)
    # random orthonormal pair (u, v)
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    w = rng.normal(size=(n, 3))
    v = np.cross(u, w)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    r = (a / np.sqrt(3.0))[:, None]
    out = np.zeros((n, 12), np.float64)
    for k in range(3):
        ang = 2.0 * np.pi * k / 3.0
        p = c + r * (np.cos(ang) * u + np.sin(ang) * v)
        out[:, 4 * k:4 * k + 3] = np.rint(p * 1e6) / 1e6          # integer micro-units
    assert out[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]].min() > 0
    return out.astype(np.float32)


def bbox_like_reference(tris):
    """Running bounds exactly as the reference host keeps them (max starts at FLT_MIN)."""
    xyz = tris.reshape(-1, 4)[:, :3]
    lo = xyz.min(axis=0)
    hi = np.maximum(xyz.max(axis=0), np.float32(np.finfo(np.float32).tiny))
    return np.append(lo, 0).astype(np.float32), np.append(hi, 0).astype(np.float32)


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import write_scenes
    n = int(sys.argv[1])
    out = sys.argv[2]
    seed = int(sys.argv[sys.argv.index("--seed") + 1]) if "--seed" in sys.argv else 20261018
    box = float(sys.argv[sys.argv.index("--box") + 1]) if "--box" in sys.argv else 60.0
    write_scenes.write_variant("grid", out)
    t = soup(n, seed, box_size=box)
    write_scenes.write_triangles(os.path.join(out, "triangles.txt"), t[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]])
    print("wrote %d triangles to %s/triangles.txt" % (n, out))
