#!/usr/bin/env python3
"""Regenerate the reference-format scene files from scenes/scene_data.py.

    python scenes/write_scenes.py <variant> <dir> [--mesh NAME]

Formats (readers: CLSuperPathTracer.c:62-139):
  spheres.txt / squares.txt / planes.txt : 9 decimal ints, one per line, no trailing newline
  triangles.txt : per triangle  x\\ny\\nz\\n\\n (x3) then one more \\n ; the last triangle ends right
                  after its last z (no trailing newline) — a trailing newline would make the
                  reference's feof() loop read a spurious extra triangle
  lights.txt    : x\\ny\\nz\\nintensity per light, no trailing newline
Also provides write_triangles() for synthetic meshes (configs 4/5).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import scene_data  # noqa: E402

VARIANT_DIRS = {
    "base": "CLSuperPathTracer",
    "lmem": "CLSuperPathTracer_lmem",
    "nodof": "CLSuperPathTracer_lmem_NoDoF",
    "grid": "CLSuperPathTracer_trianglegrid",
    "bidir": "CLSuperBidirectionalPathTracer",
}


def mesh_text(name):
    m = scene_data.MESHES[name]
    parts = []
    nf = len(m["faces"])
    for fi, face in enumerate(m["faces"]):
        for vi, v in enumerate(face):
            x, y, z = m["verts"][v]
            last = fi == nf - 1 and vi == 2
            parts.append(x + "\n" + y + "\n" + z + ("" if last else "\n\n"))
        if fi != nf - 1:
            parts.append("\n")
    return "".join(parts) + m["tail"]


def write_triangles(path, tris):
    """tris: iterable of 9-float rows (v0 v1 v2), written with %f like the reference meshes."""
    with open(path, "w") as f:
        first = True
        for t in tris:
            if not first:
                f.write("\n\n\n")
            first = False
            f.write("\n\n".join("\n".join("%f" % c for c in t[3 * v:3 * v + 3]) for v in range(3)))


def write_variant(variant, out_dir, mesh=None):
    v = scene_data.VARIANTS[variant]
    os.makedirs(out_dir, exist_ok=True)
    open(os.path.join(out_dir, "spheres.txt"), "w").write("\n".join(str(x) for x in v["spheres"]))
    sq = "\n".join(str(x) for x in v["squares"])
    open(os.path.join(out_dir, "squares.txt"), "w").write(sq)
    if variant == "nodof":
        # the NoDoF host opens planes.txt (CLSuperPathTracer_lmem_NoDoF/CLSuperPathTracer.c:303), which
        # the reference forgot to ship: same content as its squares.txt
        open(os.path.join(out_dir, "planes.txt"), "w").write(sq)
    open(os.path.join(out_dir, "lights.txt"), "w").write("\n".join(v["lights"]))
    open(os.path.join(out_dir, "triangles.txt"), "w").write(mesh_text(mesh or v["mesh"]))
    if variant == "base":
        open(os.path.join(out_dir, "torus.txt"), "w").write(mesh_text("torus"))


def verify_against_reference(ref="/root/reference"):
    import filecmp
    import tempfile
    ok = True
    for variant, d in VARIANT_DIRS.items():
        with tempfile.TemporaryDirectory() as tmp:
            write_variant(variant, tmp)
            for fn in os.listdir(tmp):
                if fn == "planes.txt":
                    continue
                same = filecmp.cmp(os.path.join(tmp, fn), os.path.join(ref, d, fn), shallow=False)
                print("%-6s %-14s %s" % (variant, fn, "identical" if same else "DIFFERENT"))
                ok &= same
    return ok


if __name__ == "__main__":
    if sys.argv[1] == "--verify":
        sys.exit(0 if verify_against_reference() else 1)
    mesh = None
    if "--mesh" in sys.argv:
        mesh = sys.argv[sys.argv.index("--mesh") + 1]
    write_variant(sys.argv[1], sys.argv[2], mesh)
