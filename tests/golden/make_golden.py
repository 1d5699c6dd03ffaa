#!/usr/bin/env python3
"""Generates tests/golden/*.json|*.npz from the REFERENCE ITSELF: the unmodified reference host programs
and kernels compiled by `make -C oracle ref` (oracle/refrt).  Only runnable in the build container
(needs /root/reference); the outputs are committed so the pin travels.

  golden_rng.json      MWC64XVEC2 / randomizeId known answers from the reference's own inline functions
  golden_trace.npz     TraceRay(base) and TraceRay(lmem) on random rays: material, t, normal (bit patterns)
  golden_host.json     what the reference hosts print: camera, counts, bbox, grid size; PAM header bytes
  golden_images.npz    result.ppm of every variant for two seed sets at 512x512: SHA-256 + selected rows
  golden_grid.npz      cell contents (sorted ids) written by the reference's initTrianglesGrid kernel
  golden_vlpgrid.npz   CLSuperMetropolisPathTracer_vlpgrid: VLP bounding box, VLP grid cells and frames of its pathTracer kernel
                       (injected VPL buffers + grid), all from the reference's own kernels
  golden_metropolis.npz  seed paths and VPL buffers of lightTracer + MetropolisLightTracer in FIX mode (one-line patch, see
                       oracle/Makefile) for several path counts / mutation rounds / seeds
  golden_bidir.npz     CLSuperBidirectionalPathTracer: the VPL buffer its lightTracer kernel writes (bit patterns,
                       several N_VLP), result.ppm SHA-256 + rows, and what the host prints
"""
import ctypes as C
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "scenes"))
import write_scenes  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
SEED_SETS = [(1, 2, 3, 4), (123456789, 42, 7, 99999)]
ROWS = [100, 120, 250, 300, 350, 400, 511]
METRO_EXTRA_SEEDS = [(1, 0xC0000001, 3, 4), (5, 0xF0000000, 6, 7), (0xE0000000, 2, 3, 4), (0x40000000, 0x50000000, 9, 11),
                     (0x26666666, 0xD9999999, 0, 0), (0x26666661, 0x59999999, 7, 5), (0x2666666F, 0x19999999, 1, 2)]   # first pair accepted, pointing down
METRO_LIGHTS_BELOW = np.array([[9, 0.2, 1.0, 400], [3, -0.3, 2.0, 300], [12, 0.1, 0.5, 200]], np.float32)
FRAME_VLPGRID = (256, 192)                      # frames of the vlpgrid program's pathTracer kernel
FRAME_VLPGRID_ROWS = [40, 60, 80, 100, 130, 160, 191]


def lib(variant):
    return C.CDLL(os.path.join(REF, "libref_%s.so" % variant))


def rng_golden():
    L = lib("base")
    L.ref_probe_randomize_id.restype = C.c_uint32
    out = {"randomize_id": {str(i): int(L.ref_probe_randomize_id(C.c_uint32(i))) for i in (0, 1, 61, 12345, 262143, 2 ** 31 - 1, 2 ** 32 - 1)},
           "streams": []}
    for seeds in SEED_SETS + [(0, 0, 0, 0), (134217727,) * 4]:
        for gid in (0, 1, 12345, 262143, 16777215):
            n = 64
            f = (C.c_float * (2 * n))()
            st = (C.c_uint32 * 4)()
            L.ref_probe_rng((C.c_uint32 * 4)(*seeds), C.c_uint32(gid), n, f, st)
            bits = np.frombuffer(bytes(f), np.uint32)
            out["streams"].append({"seeds": list(seeds), "gid": gid, "float_bits": [int(b) for b in bits], "state": [int(s) for s in st]})
    json.dump(out, open(os.path.join(HERE, "golden_rng.json"), "w"))


def trace_golden(tmp):
    rng = np.random.default_rng(20261018)
    d = os.path.join(tmp, "lmem")
    write_scenes.write_variant("lmem", d)
    from oracle.pyoracle import OracleLib
    o = OracleLib(0)
    sc = o.load_scene_dir(d, "lmem")
    n = 4000
    # rays from the camera region towards the scene plus fully random ones
    orig = np.empty((n, 3), np.float32)
    dirs = np.empty((n, 3), np.float32)
    orig[: n // 2] = np.array([17, 16, 8], np.float32) + rng.normal(0, 0.1, (n // 2, 3)).astype(np.float32)
    tgt = np.stack([rng.uniform(0, 18, n // 2), rng.uniform(-2, 8, n // 2), rng.uniform(0, 13, n // 2)], 1).astype(np.float32)
    dd = tgt - orig[: n // 2]
    dirs[: n // 2] = dd / np.linalg.norm(dd, axis=1, keepdims=True)
    orig[n // 2:] = np.stack([rng.uniform(0, 18, n - n // 2), rng.uniform(-3, 8, n - n // 2), rng.uniform(0.01, 13, n - n // 2)], 1)
    dd = rng.normal(0, 1, (n - n // 2, 3))
    dirs[n // 2:] = dd / np.linalg.norm(dd, axis=1, keepdims=True)
    tin = np.where(rng.uniform(size=n) < 0.5, 1e9, rng.uniform(0.5, 30, n)).astype(np.float32)
    res = {}
    for variant in ("base", "lmem"):
        L = lib(variant)
        L.ref_probe_trace_ray.restype = C.c_int
        m = np.zeros(n, np.int32); t = np.zeros(n, np.float32); nn = np.zeros((n, 3), np.float32)
        tris = np.ascontiguousarray(sc["triangles"], np.float32)
        for k in range(n):
            tt = C.c_float(float(tin[k]))
            no = (C.c_float * 3)()
            m[k] = L.ref_probe_trace_ray((C.c_float * 3)(*orig[k]), (C.c_float * 3)(*dirs[k]), C.byref(tt), no,
                                         (C.c_int32 * 9)(*[int(x) for x in sc["spheres"]]), (C.c_int32 * 9)(*[int(x) for x in sc["squares"]]),
                                         tris.ctypes.data_as(C.c_void_p), tris.shape[0])
            t[k] = tt.value
            nn[k] = no[:]
        res["m_" + variant] = m; res["t_" + variant] = t.view(np.uint32); res["n_" + variant] = nn.view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "golden_trace.npz"), origins=orig, dirs=dirs, t_in=tin, **res)


def run_ref(variant, d, w, h, seeds, extra=()):
    env = dict(os.environ, PT_SEEDS=",".join(str(s) for s in seeds))
    exe = "CLSuperBidirectionalPathTracer" if variant == "bidir" else "CLSuperPathTracer"
    out = subprocess.run([os.path.join(REF, "bin", variant, exe), str(w), str(h), *extra], cwd=d, env=env,
                         capture_output=True, text=True, check=True).stdout
    raw = open(os.path.join(d, "result.ppm"), "rb").read()
    k = raw.index(b"ENDHDR\n") + 7
    return out, raw[:k], np.frombuffer(raw[k:], np.uint8).reshape(h, w, 4)


def images_and_host(tmp):
    host = {}
    imgs = {}
    for variant in ("base", "lmem", "nodof", "grid"):
        d = os.path.join(tmp, "img_" + variant)
        write_scenes.write_variant(variant, d)
        for si, seeds in enumerate(SEED_SETS):
            out, hdr, img = run_ref(variant, d, 512, 512, seeds)
            imgs["%s_s%d_sha256" % (variant, si)] = np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8)
            imgs["%s_s%d_rows" % (variant, si)] = img[ROWS].copy()
        cam = re.search(r"Cam values:\n(.*\n.*\n.*\n.*)\n", out).group(1)
        host[variant] = {"camera_print": cam, "ntriangles": int(re.search(r"Number of triangles: (\d+)", out).group(1)),
                         "nlights": int(re.search(r"Number of lights: (\d+)", out).group(1)), "pam_header": hdr.decode(),
                         "lights_print": re.findall(r"Light \d+: .*", out)}
        if variant == "grid":
            host[variant]["bbox_print"] = re.search(r"vmax: .*", out).group(0)
            host[variant]["grid_size"] = [int(x) for x in re.search(r"Triangles grid size: (\d+) x (\d+) x (\d+)", out).groups()]
    # non-square image, torus mesh, another grid modifier
    d = os.path.join(tmp, "torus")
    write_scenes.write_variant("base", d, mesh="torus")
    out, hdr, img = run_ref("base", d, 640, 360, SEED_SETS[0])
    imgs["torus_640x360_sha256"] = np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8)
    imgs["torus_640x360_rows"] = img[[150, 200, 300, 359]].copy()
    host["torus"] = {"ntriangles": int(re.search(r"Number of triangles: (\d+)", out).group(1))}
    d = os.path.join(tmp, "gridtorus")
    write_scenes.write_variant("grid", d, mesh="torus")
    out, hdr, img = run_ref("grid", d, 512, 512, SEED_SETS[0], extra=("6.5",))
    imgs["gridtorus_m6.5_sha256"] = np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8)
    imgs["gridtorus_m6.5_rows"] = img[ROWS].copy()
    host["gridtorus_m6.5"] = {"bbox_print": re.search(r"vmax: .*", out).group(0),
                              "grid_size": [int(x) for x in re.search(r"Triangles grid size: (\d+) x (\d+) x (\d+)", out).groups()]}
    imgs["rows"] = np.array(ROWS)
    json.dump(host, open(os.path.join(HERE, "golden_host.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "golden_images.npz"), **imgs)


def grid_golden(tmp):
    """Run the reference's initTrianglesGrid kernel through refrt's CL entry points and dump the cells."""
    from oracle.pyoracle import OracleLib
    o = OracleLib(0)
    out = {}
    for name, mesh, modifier in (("default", None, 3.0), ("torus", "torus", 3.0), ("torus_fine", "torus", 40.0)):
        d = os.path.join(tmp, "g_" + name)
        write_scenes.write_variant("grid", d, mesh=mesh)
        sc = o.load_scene_dir(d, "grid")
        res, cell = o.grid_dims(sc["box_min"], sc["box_max"], sc["triangles"].shape[0], modifier)
        L = lib("grid")
        for fn in ("clCreateKernel", "clCreateBuffer", "clCreateProgramWithSource"):
            getattr(L, fn).restype = C.c_void_p
        err = C.c_int()
        k = C.c_void_p(L.clCreateKernel(None, b"initTrianglesGrid", C.byref(err)))
        ncells = int(res[0] * res[1] * res[2])
        cells = np.zeros(ncells * 128, np.uint8)
        tris = np.ascontiguousarray(sc["triangles"], np.float32)
        bc = C.c_void_p(L.clCreateBuffer(None, C.c_uint64(1 << 5), C.c_size_t(cells.nbytes), cells.ctypes.data_as(C.c_void_p), C.byref(err)))
        bt = C.c_void_p(L.clCreateBuffer(None, C.c_uint64(1 << 5), C.c_size_t(tris.nbytes), tris.ctypes.data_as(C.c_void_p), C.byref(err)))
        vmin = (C.c_float * 4)(*sc["box_min"]); r4 = (C.c_int32 * 4)(*[int(x) for x in res]); c4 = (C.c_float * 4)(*cell)
        L.clSetKernelArg(k, 0, C.c_size_t(8), C.byref(bc)); L.clSetKernelArg(k, 1, C.c_size_t(8), C.byref(bt))
        L.clSetKernelArg(k, 2, C.c_size_t(16), vmin); L.clSetKernelArg(k, 3, C.c_size_t(16), r4); L.clSetKernelArg(k, 4, C.c_size_t(16), c4)
        gws = (C.c_size_t * 1)(tris.shape[0])
        assert L.clEnqueueNDRangeKernel(None, k, 1, None, gws, None, 0, None, None) == 0
        L.clEnqueueMapBuffer.restype = C.c_void_p
        ptr = L.clEnqueueMapBuffer(None, bc, 1, 1, C.c_size_t(0), C.c_size_t(cells.nbytes), 0, None, None, C.byref(err))
        got = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(ncells, 128)).copy()
        nels = got[:, :4].copy().view(np.uint32).reshape(-1)
        ids = got[:, 4:].copy().view(np.uint16).reshape(ncells, 62)
        # atomic order is arbitrary: store each cell's ids sorted; nels may exceed 62 in the reference (overflow bug)
        srt = np.full((ncells, 62), 65535, np.uint16)
        for c in range(ncells):
            n = min(int(nels[c]), 62)
            srt[c, :n] = np.sort(ids[c, :n])
        out[name + "_res"] = res; out[name + "_cell"] = cell.view(np.uint32); out[name + "_nels"] = nels; out[name + "_ids"] = srt
        # frames of the reference's own pathTracer kernel on this buffer and on the grid in its DEFINED form (ascending ids:
        # the atomic_inc arrival order of initVLPsGrid is not part of the program's meaning): full SHA-256 + selected rows
        if name in ("bidir", "synthetic"):
            cb = np.zeros((ncells, 128), np.uint8)
            cb[:, :4] = nels.astype(np.uint32).reshape(-1, 1).view(np.uint8)
            ids_clean = np.where(srt == 65535, 0, srt).astype(np.uint16)
            cb[:, 4:] = ids_clean.view(np.uint8).reshape(ncells, 124)
            cam = o.camera()
            for si, seeds in enumerate(SEED_SETS):
                W, H = FRAME_VLPGRID
                img = ref_vlpgrid_pathtracer(L, sc, cam, seeds, W, H, vpl, np.ascontiguousarray(cb.reshape(-1)), vmin, res, cell)
                key = "%s_frame_s%d" % (name, si)
                out[key + "_sha256"] = np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8)
                out[key + "_rows"] = img[FRAME_VLPGRID_ROWS]
        out[name + "_modifier"] = np.float32(modifier)
    np.savez_compressed(os.path.join(HERE, "golden_grid.npz"), **out)


def ref_light_tracer(L, sc, seeds, n_vlp):
    """The reference's lightTracer kernel through refrt's CL entry points, argument order of
    CLSuperBidirectionalPathTracer.c:154-177 -> (n_vlp*nlights, 4) float32."""
    for fn in ("clCreateKernel", "clCreateBuffer", "clEnqueueMapBuffer"):
        getattr(L, fn).restype = C.c_void_p
    err = C.c_int()
    k = C.c_void_p(L.clCreateKernel(None, b"lightTracer", C.byref(err)))
    COPY = C.c_uint64(1 << 5)

    def buf(a):
        return C.c_void_p(L.clCreateBuffer(None, COPY, C.c_size_t(a.nbytes), a.ctypes.data_as(C.c_void_p), C.byref(err)))
    sph = np.ascontiguousarray(sc["spheres"], np.int32); sq = np.ascontiguousarray(sc["squares"], np.int32)
    tris = np.ascontiguousarray(sc["triangles"], np.float32); lights = np.ascontiguousarray(sc["lights"], np.float32)
    nl = lights.shape[0]
    vpl = np.full((n_vlp * nl, 4), np.nan, np.float32)
    bs, bq, bt, bl, bv = buf(sph), buf(sq), buf(tris), buf(lights), buf(vpl)
    ntri = C.c_int32(tris.shape[0]); nlc = C.c_int32(nl); sd = (C.c_uint32 * 4)(*seeds)
    args = [(8, C.byref(bs)), (8, C.byref(bq)), (8, C.byref(bt)), (4, C.byref(ntri)), (8, C.byref(bl)), (4, C.byref(nlc)),
            (8, C.byref(bv)), (16, sd), (36, None), (36, None), (48 * tris.shape[0], None), (16 * nl, None)]
    for i, (size, ptr) in enumerate(args):
        assert L.clSetKernelArg(k, i, C.c_size_t(size), ptr) == 0
    gws = (C.c_size_t * 1)(n_vlp)
    assert L.clEnqueueNDRangeKernel(None, k, 1, None, gws, None, 0, None, None) == 0
    ptr = L.clEnqueueMapBuffer(None, bv, 1, 1, C.c_size_t(0), C.c_size_t(vpl.nbytes), 0, None, None, C.byref(err))
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n_vlp * nl, 4)).copy()


def bidir_golden(tmp):
    from oracle.pyoracle import OracleLib
    o = OracleLib(0)
    d = os.path.join(tmp, "bidir")
    write_scenes.write_variant("bidir", d)
    sc = o.load_scene_dir(d, "bidir")
    L = lib("bidir")
    out = {}
    env_seeds = os.environ.pop("PT_SEEDS", None)     # the probe passes its seeds as the kernel argument
    for si, seeds in enumerate(SEED_SETS):
        # 512: the default; 700: total/512 truncates to 2; 96: total/512 == 0 -> inf / NaN intensities
        for n_vlp in (512, 700, 96):
            out["vpl_s%d_n%d" % (si, n_vlp)] = ref_light_tracer(L, sc, seeds, n_vlp).view(np.uint32)
    if env_seeds is not None:
        os.environ["PT_SEEDS"] = env_seeds
    host = {}
    for si, seeds in enumerate(SEED_SETS):
        log, hdr, img = run_ref("bidir", d, 512, 512, seeds)
        out["img_s%d_sha256" % si] = np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8)
        out["img_s%d_rows" % si] = img[ROWS].copy()
    cam = re.search(r"Cam values:\n(.*\n.*\n.*\n.*)\n", log).group(1)
    host["bidir"] = {"camera_print": cam, "ntriangles": int(re.search(r"Number of triangles: (\d+)", log).group(1)),
                     "nlights": int(re.search(r"Number of lights: (\d+)", log).group(1)), "pam_header": hdr.decode(),
                     "vpl_print": re.search(r"virtual light sampling : (\d+) virtual lights", log).group(1)}
    for n_vlp, size in ((700, (320, 256)), (96, (256, 256))):
        log, hdr, img = run_ref("bidir", d, size[0], size[1], SEED_SETS[0], extra=(str(n_vlp),))
        out["img_n%d_sha256" % n_vlp] = np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8)
        out["img_n%d" % n_vlp] = img.copy() if n_vlp == 96 else img[[100, 180, 255]].copy()
        out["img_n%d_size" % n_vlp] = np.array(size)
    out["rows"] = np.array(ROWS)
    out["host_json"] = np.frombuffer(json.dumps(host).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "golden_bidir.npz"), **out)


def canonical_paths(paths):
    """The reference's Path records carry stack garbage in everything that is not defined — vertices at and beyond `length`, the
    w components, the padding after `length`: zero it, so that the golden file is a function of the program alone."""
    p = np.array(paths, np.uint32).reshape(-1, 20).copy()
    for i in range(p.shape[0]):
        n = min(int(p[i, 16]), 4)
        p[i, 4 * n:16] = 0
        p[i, 3:16:4] = 0
        p[i, 17:] = 0
    return p


def ref_metropolis_vpls(L, sc, seeds, n_paths, rounds):
    """FIX build of the reference (oracle/Makefile: `float t = 1e9;` in VerifyIntersection, nothing else): kernel lightTracer
    writes the seed paths into ITS OWN buffer (the reference host hands it d_virtual_lights by mistake,
    CLSuperMetropolisPathTracer.c:579), kernel MetropolisLightTracer mutates them `rounds` times and deposits 4 VPLs per path.
    -> (seed paths as (n*nl, 20) uint32 words [4 x float4 + length + 3 pad], VPLs as (4*n*nl, 4) float32)."""
    for fn in ("clCreateKernel", "clCreateBuffer", "clEnqueueMapBuffer"):
        getattr(L, fn).restype = C.c_void_p
    err = C.c_int()
    COPY = C.c_uint64(1 << 5)

    def buf(a):
        return C.c_void_p(L.clCreateBuffer(None, COPY, C.c_size_t(a.nbytes), a.ctypes.data_as(C.c_void_p), C.byref(err)))
    sph = np.ascontiguousarray(sc["spheres"], np.int32); sq = np.ascontiguousarray(sc["squares"], np.int32)
    tris = np.ascontiguousarray(sc["triangles"], np.float32); lights = np.ascontiguousarray(sc["lights"], np.float32)
    nl = lights.shape[0]
    paths = np.zeros((n_paths * nl, 20), np.uint32)                 # sizeof(Path) = 80
    vpl = np.full((n_paths * nl * 4, 4), np.nan, np.float32)
    bs, bq, bt, bl, bp, bv = buf(sph), buf(sq), buf(tris), buf(lights), buf(paths), buf(vpl)
    ntri = C.c_int32(tris.shape[0]); nlc = C.c_int32(nl); sd = (C.c_uint32 * 4)(*seeds); rd = C.c_int32(rounds)
    k = C.c_void_p(L.clCreateKernel(None, b"lightTracer", C.byref(err)))
    args = [(8, C.byref(bs)), (8, C.byref(bq)), (8, C.byref(bt)), (4, C.byref(ntri)), (8, C.byref(bl)), (4, C.byref(nlc)),
            (8, C.byref(bp)), (16, sd), (36, None), (36, None), (16 * nl, None)]
    for i, (size, ptr) in enumerate(args):
        assert L.clSetKernelArg(k, i, C.c_size_t(size), ptr) == 0, i
    gws = (C.c_size_t * 1)(n_paths)
    assert L.clEnqueueNDRangeKernel(None, k, 1, None, gws, None, 0, None, None) == 0
    ptr = L.clEnqueueMapBuffer(None, bp, 1, 1, C.c_size_t(0), C.c_size_t(paths.nbytes), 0, None, None, C.byref(err))
    paths_out = canonical_paths(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=paths.shape))
    k2 = C.c_void_p(L.clCreateKernel(None, b"MetropolisLightTracer", C.byref(err)))
    args = [(8, C.byref(bs)), (8, C.byref(bq)), (8, C.byref(bt)), (4, C.byref(ntri)), (8, C.byref(bl)), (4, C.byref(nlc)),
            (8, C.byref(bp)), (8, C.byref(bv)), (16, sd), (4, C.byref(rd)), (36, None), (36, None), (16 * nl, None)]
    for i, (size, ptr) in enumerate(args):
        assert L.clSetKernelArg(k2, i, C.c_size_t(size), ptr) == 0, i
    assert L.clEnqueueNDRangeKernel(None, k2, 1, None, gws, None, 0, None, None) == 0
    ptr = L.clEnqueueMapBuffer(None, bv, 1, 1, C.c_size_t(0), C.c_size_t(vpl.nbytes), 0, None, None, C.byref(err))
    return paths_out, np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=vpl.shape).copy()


def metropolis_golden(tmp):
    """golden_metropolis.npz: seed paths and VPL buffers of the FIX-mode Metropolis kernels for several (paths, rounds, seeds)."""
    from oracle.pyoracle import OracleLib
    o = OracleLib(0)
    d = os.path.join(tmp, "metro")
    write_scenes.write_variant("bidir", d)
    sc = o.load_scene_dir(d, "bidir")
    L = lib("vlpgrid_fix")
    env_seeds = os.environ.pop("PT_SEEDS", None)
    out = {}
    for si, seeds in enumerate(SEED_SETS):
        for n_paths, rounds in ((512, 8), (512, 0), (300, 1), (128, 40)):
            paths, vpl = ref_metropolis_vpls(L, sc, seeds, n_paths, rounds)
            key = "s%d_n%d_r%d" % (si, n_paths, rounds)
            out[key + "_paths"] = paths
            out[key + "_vpl"] = vpl.view(np.uint32)
    # a light list below the squares: their undersides face the lights, so many VPLs are non-zero and the mutation rounds
    # (vertex additions along the second direction) change the buffer
    sc2 = dict(sc, lights=METRO_LIGHTS_BELOW)
    for n_paths, rounds in ((256, 0), (256, 8), (256, 3)):
        paths, vpl = ref_metropolis_vpls(L, sc2, SEED_SETS[0], n_paths, rounds)
        key = "below_n%d_r%d" % (n_paths, rounds)
        out[key + "_paths"] = paths
        out[key + "_vpl"] = vpl.view(np.uint32)
    out["below_lights"] = METRO_LIGHTS_BELOW
    # Mutate itself (the kernel keeps the mutated path private): the reference's function through ref_probe_mutate on the seed
    # paths of 96 work-items x 2 lights, after 1, 2 and 8 rounds
    paths0 = out["s0_n512_r0_paths"]
    sph = np.ascontiguousarray(sc["spheres"], np.int32); sq = np.ascontiguousarray(sc["squares"], np.int32)
    tris = np.ascontiguousarray(sc["triangles"], np.float32); lights = np.ascontiguousarray(sc["lights"], np.float32)
    mut = {r: [] for r in (1, 2, 8)}
    which = []
    for gi in range(0, 512, 16):
        for l in range(lights.shape[0]):
            which.append((gi, l))
            for r in mut:
                p = paths0[gi + l * 512].copy()
                L.ref_probe_mutate((C.c_uint32 * 4)(*SEED_SETS[0]), C.c_uint32(gi), sph.ctypes.data_as(C.c_void_p), sq.ctypes.data_as(C.c_void_p),
                                   tris.ctypes.data_as(C.c_void_p), C.c_int(tris.shape[0]), (C.c_float * 3)(*lights[l, :3]),
                                   p.ctypes.data_as(C.c_void_p), C.c_int(r))
                mut[r].append(p)
    out["mutate_which"] = np.array(which, np.int32)
    for r in mut:
        out["mutate_r%d" % r] = canonical_paths(mut[r])
    # The first RNG pair of EVERY work-item is (seeds.x ^ seeds.z, seeds.y ^ seeds.w) * 2^-32 (the seeding XORs one hash into all
    # four words), and Mutate branches on it: with the host's 27-bit seeds it is always < 0.03125, so `y > 0.3 / 0.7 / 0.9` and
    # `probability < x` never fire.  Seeds outside that range reach the other branches (the kernels take any uint4).
    for ei, seeds in enumerate(METRO_EXTRA_SEEDS):
        paths, vpl = ref_metropolis_vpls(L, sc, seeds, 256, 8)
        out["extra%d_paths" % ei] = paths
        out["extra%d_vpl" % ei] = vpl.view(np.uint32)
        res = []
        for gi in range(256):
            for l in range(lights.shape[0]):
                p = paths[gi + l * 256].copy()
                L.ref_probe_mutate((C.c_uint32 * 4)(*seeds), C.c_uint32(gi), sph.ctypes.data_as(C.c_void_p), sq.ctypes.data_as(C.c_void_p),
                                   tris.ctypes.data_as(C.c_void_p), C.c_int(tris.shape[0]), (C.c_float * 3)(*lights[l, :3]),
                                   p.ctypes.data_as(C.c_void_p), C.c_int(8))
                res.append(p)
        out["extra%d_mutated" % ei] = canonical_paths(res)
    out["extra_seeds"] = np.array(METRO_EXTRA_SEEDS, np.uint32)
    if env_seeds is not None:
        os.environ["PT_SEEDS"] = env_seeds
    np.savez_compressed(os.path.join(HERE, "golden_metropolis.npz"), **out)


def ref_vlpgrid_pathtracer(L, sc, cam, seeds, W, H, vpl, cells_bytes, vmin, res, cell):
    """Kernel pathTracer of CLSuperMetropolisPathTracer_vlpgrid (metropolispathtracer.ocl:649-684) through refrt's CL entry
    points, argument order of CLSuperMetropolisPathTracer.c:324-392, on an injected VPL buffer and VLP grid -> (H, W, 4) uint8."""
    for fn in ("clCreateKernel", "clCreateBuffer", "clEnqueueMapBuffer"):
        getattr(L, fn).restype = C.c_void_p
    err = C.c_int()
    k = C.c_void_p(L.clCreateKernel(None, b"pathTracer", C.byref(err)))
    COPY = C.c_uint64(1 << 5)

    def buf(a):
        return C.c_void_p(L.clCreateBuffer(None, COPY, C.c_size_t(a.nbytes), a.ctypes.data_as(C.c_void_p), C.byref(err)))
    sph = np.ascontiguousarray(sc["spheres"], np.int32); sq = np.ascontiguousarray(sc["squares"], np.int32)
    tris = np.ascontiguousarray(sc["triangles"], np.float32); lights = np.ascontiguousarray(sc["lights"], np.float32)
    nl = lights.shape[0]
    img = np.zeros((H, W, 4), np.uint8)
    vpl = np.ascontiguousarray(vpl, np.float32).reshape(-1, 4)
    bi, bs, bq, bt, bv, bc, bl = buf(img), buf(sph), buf(sq), buf(tris), buf(vpl), buf(cells_bytes), buf(lights)
    ntri = C.c_int32(tris.shape[0]); nv = C.c_int32(vpl.shape[0]); nlc = C.c_int32(nl); sd = (C.c_uint32 * 4)(*seeds)
    f4 = lambda v: (C.c_float * 4)(*[float(x) for x in v])
    r4 = (C.c_int32 * 4)(*[int(x) for x in res])
    args = [(8, C.byref(bi)), (8, C.byref(bs)), (8, C.byref(bq)), (8, C.byref(bt)), (4, C.byref(ntri)), (8, C.byref(bv)), (4, C.byref(nv)),
            (8, C.byref(bc)), (16, f4(vmin)), (16, f4(cell)), (16, r4), (8, C.byref(bl)), (4, C.byref(nlc)),
            (16, f4(cam["cam_forward"])), (16, f4(cam["cam_up"])), (16, f4(cam["cam_right"])), (16, f4(cam["eye_offset"])), (16, sd),
            (36, None), (36, None), (16 * nl, None)]
    for i, (size, ptr) in enumerate(args):
        assert L.clSetKernelArg(k, i, C.c_size_t(size), ptr) == 0, i
    gws = (C.c_size_t * 2)(W, H)
    assert L.clEnqueueNDRangeKernel(None, k, 2, None, gws, None, 0, None, None) == 0
    ptr = L.clEnqueueMapBuffer(None, bi, 1, 1, C.c_size_t(0), C.c_size_t(img.nbytes), 0, None, None, C.byref(err))
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(H, W, 4)).copy()


def vlpgrid_golden(tmp):
    """VLP bounding box (reduceMinAndMax_lmem + _nwg) and VLP grid (initVLPsGrid) of CLSuperMetropolisPathTracer_vlpgrid, run
    by the reference's own kernels (libref_vlpgrid.so through refrt's CL entry points, launched as its host does:
    CLSuperMetropolisPathTracer.c:262-296, 298-321, 586-647) on injected VLP buffers."""
    from oracle.pyoracle import OracleLib
    o = OracleLib(0)
    d = os.path.join(tmp, "vlp_bidir")
    write_scenes.write_variant("bidir", d)
    sc = o.load_scene_dir(d, "bidir")
    rng = np.random.default_rng(20261018)
    synth = np.zeros((3000, 4), np.float32)
    synth[:, :3] = rng.uniform(-5, 30, (3000, 3))
    synth[:, 3] = rng.uniform(0, 0.02, 3000) ** 2 * 50
    synth[rng.random(3000) < 0.3, 3] = 0.0                         # dummy lights
    few = np.array([[1, 2, 3, 0.25], [0, 0, 0, 0], [4, 1, 0.5, 0.01]], np.float32)
    buffers = {"bidir": ref_light_tracer(lib("bidir"), sc, SEED_SETS[0], 512), "synthetic": synth, "few": few,
               "all_dummy": np.zeros((300, 4), np.float32)}
    L = lib("vlpgrid")
    for fn in ("clCreateKernel", "clCreateBuffer", "clEnqueueMapBuffer"):
        getattr(L, fn).restype = C.c_void_p
    err = C.c_int()
    out = {}
    for name, vpl in buffers.items():
        vpl = np.ascontiguousarray(vpl, np.float32).reshape(-1, 4)
        n = vpl.shape[0]
        lws = 256
        nwg = (n + lws - 1) // lws
        box = np.zeros(max(nwg, 1) * 8, np.float32)
        b1 = C.c_void_p(L.clCreateBuffer(None, C.c_uint64(1 << 5), C.c_size_t(vpl.nbytes), vpl.ctypes.data_as(C.c_void_p), C.byref(err)))
        b2 = C.c_void_p(L.clCreateBuffer(None, C.c_uint64(1 << 5), C.c_size_t(box.nbytes), box.ctypes.data_as(C.c_void_p), C.byref(err)))
        nn = C.c_int(n)
        k = C.c_void_p(L.clCreateKernel(None, b"reduceMinAndMax_lmem", C.byref(err)))
        L.clSetKernelArg(k, 0, C.c_size_t(8), C.byref(b1)); L.clSetKernelArg(k, 1, C.c_size_t(8), C.byref(b2))
        L.clSetKernelArg(k, 2, C.c_size_t(32 * lws), None); L.clSetKernelArg(k, 3, C.c_size_t(4), C.byref(nn))
        assert L.clEnqueueNDRangeKernel(None, k, 1, None, (C.c_size_t * 1)(nwg * lws), (C.c_size_t * 1)(lws), 0, None, None) == 0
        if nwg > 1:
            k2 = C.c_void_p(L.clCreateKernel(None, b"reduceMinAndMax_lmem_nwg", C.byref(err)))
            nw = C.c_int(nwg)
            L.clSetKernelArg(k2, 0, C.c_size_t(8), C.byref(b2)); L.clSetKernelArg(k2, 1, C.c_size_t(8), C.byref(b2))
            L.clSetKernelArg(k2, 2, C.c_size_t(32 * lws), None); L.clSetKernelArg(k2, 3, C.c_size_t(4), C.byref(nw))
            assert L.clEnqueueNDRangeKernel(None, k2, 1, None, (C.c_size_t * 1)(lws), (C.c_size_t * 1)(lws), 0, None, None) == 0
        ptr = L.clEnqueueMapBuffer(None, b2, 1, 1, C.c_size_t(0), C.c_size_t(32), 0, None, None, C.byref(err))
        bb = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(8,)).copy()
        vmin, vmax = bb[:4].copy(), bb[4:].copy()
        out[name + "_vpl"] = vpl.view(np.uint32)
        out[name + "_vmin"] = vmin.view(np.uint32); out[name + "_vmax"] = vmax.view(np.uint32)
        if name == "all_dummy":
            continue                                               # FLT_MAX / FLT_MIN box: the host's grid formula is meaningless
        res, cell = o.grid_dims(vmin, vmax, n, 3.0)
        ncells = int(res[0] * res[1] * res[2])
        cells = np.zeros(ncells * 128, np.uint8)
        bc = C.c_void_p(L.clCreateBuffer(None, C.c_uint64(1 << 5), C.c_size_t(cells.nbytes), cells.ctypes.data_as(C.c_void_p), C.byref(err)))
        kg = C.c_void_p(L.clCreateKernel(None, b"initVLPsGrid", C.byref(err)))
        v4 = (C.c_float * 4)(*vmin); r4 = (C.c_int32 * 4)(*[int(x) for x in res]); c4 = (C.c_float * 4)(*cell)
        L.clSetKernelArg(kg, 0, C.c_size_t(8), C.byref(bc)); L.clSetKernelArg(kg, 1, C.c_size_t(8), C.byref(b1))
        L.clSetKernelArg(kg, 2, C.c_size_t(16), v4); L.clSetKernelArg(kg, 3, C.c_size_t(16), r4); L.clSetKernelArg(kg, 4, C.c_size_t(16), c4)
        # a cell that overflows stores whichever 62 lights ARRIVE first (atomic_inc): a race between work-items.  Run this one
        # kernel with the work-items in sequential order (one OpenMP thread), so that the stored set — the 62 lowest indices —
        # is a function of the program and the file regenerates bit for bit.
        gomp = C.CDLL("libgomp.so.1")
        gomp.omp_get_max_threads.restype = C.c_int
        nthreads = gomp.omp_get_max_threads()
        gomp.omp_set_num_threads(1)
        assert L.clEnqueueNDRangeKernel(None, kg, 1, None, (C.c_size_t * 1)(n), None, 0, None, None) == 0
        gomp.omp_set_num_threads(nthreads)
        ptr = L.clEnqueueMapBuffer(None, bc, 1, 1, C.c_size_t(0), C.c_size_t(cells.nbytes), 0, None, None, C.byref(err))
        got = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(ncells, 128)).copy()
        nels = got[:, :4].copy().view(np.uint32).reshape(-1)
        ids = got[:, 4:].copy().view(np.uint16).reshape(ncells, 62)
        srt = np.full((ncells, 62), 65535, np.uint16)              # atomic_inc order is arbitrary: ids stored sorted
        for c in range(ncells):
            m = min(int(nels[c]), 62)
            srt[c, :m] = np.sort(ids[c, :m])
        out[name + "_res"] = res; out[name + "_cell"] = cell.view(np.uint32); out[name + "_nels"] = nels; out[name + "_ids"] = srt
        # frames of the reference's own pathTracer kernel on this buffer and on the grid in its DEFINED form (ascending ids:
        # the atomic_inc arrival order of initVLPsGrid is not part of the program's meaning): full SHA-256 + selected rows
        if name in ("bidir", "synthetic"):
            cb = np.zeros((ncells, 128), np.uint8)
            cb[:, :4] = nels.astype(np.uint32).reshape(-1, 1).view(np.uint8)
            ids_clean = np.where(srt == 65535, 0, srt).astype(np.uint16)
            cb[:, 4:] = ids_clean.view(np.uint8).reshape(ncells, 124)
            cam = o.camera()
            for si, seeds in enumerate(SEED_SETS):
                W, H = FRAME_VLPGRID
                img = ref_vlpgrid_pathtracer(L, sc, cam, seeds, W, H, vpl, np.ascontiguousarray(cb.reshape(-1)), vmin, res, cell)
                key = "%s_frame_s%d" % (name, si)
                out[key + "_sha256"] = np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8)
                out[key + "_rows"] = img[FRAME_VLPGRID_ROWS]
    np.savez_compressed(os.path.join(HERE, "golden_vlpgrid.npz"), **out)


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref", "oracle"])
    with tempfile.TemporaryDirectory() as tmp:
        if sys.argv[1:] == ["bidir"]:
            bidir_golden(tmp)
            sys.exit(0)
        if sys.argv[1:] == ["vlpgrid"]:
            vlpgrid_golden(tmp)
            sys.exit(0)
        if sys.argv[1:] == ["metropolis"]:
            metropolis_golden(tmp)
            sys.exit(0)
        rng_golden()
        trace_golden(tmp)
        images_and_host(tmp)
        grid_golden(tmp)
        bidir_golden(tmp)
        vlpgrid_golden(tmp)
        metropolis_golden(tmp)
    print("golden vectors written to", HERE)
