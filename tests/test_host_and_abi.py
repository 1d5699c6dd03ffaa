"""CPU-only checks of the product's host code and of the C-ABI library surface (no compute calls)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import opencl_montecarlo_path_tracing_b200 as pt
from opencl_montecarlo_path_tracing_b200 import _lib, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pth?_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    cuda, host = _lib.cuda_lib(), _lib.host_lib()
    names = declared_functions("ptcuda.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(cuda, n), "libptcuda.so does not export %s" % n
        assert n in _lib.PTCUDA_SYMBOLS, "ctypes table misses %s" % n
    for n in declared_functions("pthost.h"):
        assert hasattr(host, n), "libpthost.so does not export %s" % n
    assert cuda.pt_abi_version() == 5


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of every ABI struct, as gcc sees include/ptcuda.h, against the ctypes mirrors."""
    import subprocess
    structs = {"pt_scene": _lib.pt_scene, "pt_camera": _lib.pt_camera, "pt_grid": _lib.pt_grid,
               "pt_render_params": _lib.pt_render_params, "pt_counters": _lib.pt_counters}
    lines = []
    for name, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ptcuda.h"\nint main(void){%s return 0;}\n' % "".join(lines))
    exe = str(tmp_path / "layout")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, str(src)])
    got = dict(l.split() for l in subprocess.check_output([exe], text=True).splitlines())
    for name, cls in structs.items():
        assert int(got[name]) == C.sizeof(cls), name
        for f, _ in cls._fields_:
            assert int(got["%s.%s" % (name, f)]) == getattr(cls, f).offset, (name, f)
    assert C.sizeof(_lib.pt_camera) == 64 and C.sizeof(_lib.pt_grid) == 68


@pytest.mark.skipif(_lib.cuda_lib().pt_device_count() > 0, reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run: it must fail loudly, never compute on the CPU."""
    with pytest.raises(pt.PtError):
        pt.Renderer(device=0)
    lib = _lib.cuda_lib()
    assert not lib.pt_create(0)
    assert b"error" in lib.pt_last_error()


def test_host_parsers_match_oracle(scene_dirs, oracle_sep):
    for v in ("base", "lmem", "nodof", "grid", "torus"):
        variant = "base" if v == "torus" else v
        mine = pt.load_scene_dir(scene_dirs[v], variant)
        ref = oracle_sep.load_scene_dir(scene_dirs[v], variant)
        assert np.array_equal(mine.spheres, ref["spheres"]) and np.array_equal(mine.squares, ref["squares"])
        assert np.array_equal(mine.triangles.view(np.uint32), ref["triangles"].view(np.uint32))
        assert np.array_equal(mine.lights.view(np.uint32), ref["lights"].view(np.uint32))
        assert np.array_equal(mine.box_min.view(np.uint32), ref["box_min"].view(np.uint32))
        assert np.array_equal(mine.box_max.view(np.uint32), ref["box_max"].view(np.uint32))


def test_parser_quirks(tmp_path, oracle_sep):
    """feof()/fgets() behaviour of the reference readers."""
    import write_scenes
    d = str(tmp_path)
    # trailing newline after the last light replays the last record
    open(os.path.join(d, "lights.txt"), "w").write("10\n4\n10\n200\n15\n2\n7\n150\n")
    open(os.path.join(d, "spheres.txt"), "w").write("1\n2\n3")          # fewer than 9 rows
    open(os.path.join(d, "squares.txt"), "w").write("\n".join(str(i) for i in range(12)))  # more than 9 rows
    write_scenes.write_triangles(os.path.join(d, "triangles.txt"), [[1, 2, 3, 4, 5, 6, 7, 8, 9], [2, 2, 3, 4, 5, 6, 7, 8, 10]])
    # a complete 13-line tail (z newline + both separator lines) hides EOF from feof() => spurious 3rd triangle
    open(os.path.join(d, "triangles.txt"), "a").write("\n\n\n")
    h = _lib.host_lib()
    lights = ((C.c_float * 4) * 5)()
    assert h.pth_parse_lights(os.path.join(d, "lights.txt").encode(), C.byref(lights), 0) == 3
    assert [lights[2][k] for k in range(4)] == [15.0, 2.0, 7.0, 150.0]
    assert oracle_sep.parse_lights(os.path.join(d, "lights.txt")).shape[0] == 3
    arr = (C.c_int32 * 9)(*([-7] * 9))
    assert h.pth_parse_bitmap(os.path.join(d, "spheres.txt").encode(), arr) == 3 and arr[:3] == [1, 2, 3] and arr[3] == -7
    assert h.pth_parse_bitmap(os.path.join(d, "squares.txt").encode(), arr) == 9 and arr[:] == list(range(9))
    sc = pt.load_scene_dir(d, "base")
    ref = oracle_sep.load_scene_dir(d, "base")
    assert sc.ntriangles == 3 == ref["triangles"].shape[0]
    assert np.array_equal(sc.triangles, ref["triangles"])
    # cap: MAX_TRIANGLES semantics
    assert pt.load_scene_dir(d, "base", max_triangles=1).ntriangles == 1
    with pytest.raises(FileNotFoundError):
        pt.load_scene_dir(os.path.join(d, "nope"), "base")


def test_threaded_triangle_parser_equals_the_fgets_loop(tmp_path, oracle_sep):
    """pth_parse_triangles converts the complete 13-line records of a big file on several threads and emulates fgets()/feof()
    on the rest; the oracle's reader is the reference's sequential FILE loop (CLSuperPathTracer.c:62-107).  Same floats, same
    count, same bounding box on files with every kind of ending, stray blank lines, odd tokens and over-long lines."""
    import gen_mesh
    import write_scenes
    d = str(tmp_path)
    write_scenes.write_variant("base", d)
    tris = gen_mesh.soup(20000, seed=9, box_size=20.0)
    path = os.path.join(d, "triangles.txt")
    write_scenes.write_triangles(path, tris[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]])
    clean = open(path).read()
    lines = clean.split("\n")
    variants = {
        "as written (no trailing newline)": clean,
        "trailing newline": clean + "\n",
        "complete 13-line tail": clean + "\n\n\n",
        "cut in the middle of a record": "\n".join(lines[: 13 * 15000 + 5]),
        "cut in the middle of a number": clean[: len(clean) // 2 + 3],
        "a stray blank line shifts every later record": "\n".join(lines[:1000] + [""] + lines[1000:]),
        "tokens atof reads differently": clean.replace(lines[7], " +1.5e0xyz", 1).replace(lines[20], "nan", 1).replace(lines[33], "0x1p3", 1),
        "one over-long line": "\n".join(lines[:50] + ["9" * 700] + lines[51:]),
        "empty file": "",
        "a single newline": "\n",
        "carriage returns": clean[:5000].replace("\n", "\r\n"),
    }
    for name, text in variants.items():
        open(path, "w").write(text)
        for cap in (65536, 15001, 1):
            mine = pt.load_scene_dir(d, "grid", max_triangles=cap)
            ref = oracle_sep.load_scene_dir(d, "grid", max_triangles=cap)
            assert mine.ntriangles == ref["triangles"].shape[0], (name, cap, mine.ntriangles, ref["triangles"].shape[0])
            assert np.array_equal(mine.triangles.view(np.uint32), ref["triangles"].view(np.uint32)), (name, cap)
            assert np.array_equal(mine.box_min.view(np.uint32), ref["box_min"].view(np.uint32)), (name, cap)
            assert np.array_equal(mine.box_max.view(np.uint32), ref["box_max"].view(np.uint32)), (name, cap)


def test_camera_and_grid_dims_match_oracle_and_golden(scene_dirs, oracle_sep):
    g = json.load(open(os.path.join(G, "golden_host.json")))
    cam = pt.camera()
    oc = oracle_sep.camera()
    for k in ("cam_forward", "cam_up", "cam_right", "eye_offset"):
        assert np.array_equal(np.array(getattr(cam, k)[:], np.float32).view(np.uint32), oc[k].view(np.uint32))
    line = "Cam_forward %f %f %f\nCam_up %f %f %f\nCam_right %f %f %f\n eye_offset %f %f %f" % (
        *cam.cam_forward[:3], *cam.cam_up[:3], *cam.cam_right[:3], *cam.eye_offset[:3])
    assert line == g["base"]["camera_print"]
    for v, mod in (("grid", 3.0), ("grid", 6.5), ("torus", 3.0), ("torus", 40.0)):
        sc = pt.load_scene_dir(scene_dirs[v], "grid")
        gd = pt.grid_dims(sc, mod)
        res, cell = oracle_sep.grid_dims(sc.box_min, sc.box_max, sc.ntriangles, mod)
        assert list(gd.res[:3]) == list(res[:3])
        assert np.array_equal(np.array(gd.cell_size[:3], np.float32).view(np.uint32), cell[:3].view(np.uint32))
    assert list(pt.grid_dims(pt.load_scene_dir(scene_dirs["grid"], "grid")).res[:3]) == g["grid"]["grid_size"]


def test_pam_writer(tmp_path):
    g = json.load(open(os.path.join(G, "golden_host.json")))
    img = (np.arange(512 * 512 * 4) % 251).astype(np.uint8).reshape(512, 512, 4)
    p = str(tmp_path / "x.ppm")
    pt.save_pam(p, img)
    raw = open(p, "rb").read()
    assert raw.startswith(g["base"]["pam_header"].encode()) and len(g["base"]["pam_header"]) == 69
    assert raw[69:] == img.tobytes()


def test_seeds_env(monkeypatch):
    h = _lib.host_lib()
    s = (C.c_uint32 * 4)()
    monkeypatch.setenv("PT_SEEDS", "123456789,42,7,99999")
    h.pth_seeds(s)
    assert s[:] == [123456789, 42, 7, 99999]
    monkeypatch.delenv("PT_SEEDS")
    h.pth_seeds(s)
    assert all(v < 2 ** 27 for v in s[:])      # the reference masks its wall-clock seeds to 27 bits


def test_stripe_sharding_partitions_rows():
    for H, stripe, n in ((512, 8, 2), (360, 8, 4), (1080, 16, 8), (50, 8, 3), (7, 8, 2)):
        owned = [sharding.stripe_rows(H, stripe, r, n) for r in range(n)]
        allrows = np.sort(np.concatenate(owned))
        assert np.array_equal(allrows, np.arange(H))
        for r in range(n):
            vr = sharding.virtual_rows(H, stripe, r, n)
            mapped = [sharding.map_row(v, stripe, r, n) for v in range(vr)]
            assert [m for m in mapped if m < H] == list(owned[r])
