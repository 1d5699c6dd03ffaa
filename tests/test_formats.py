"""Data formats on either side of the hot path: the PAM reader matching the reference's writer,
a real P6 PPM, a self-contained PNG writer and the OBJ -> triangles.txt importer.  CPU only."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import opencl_montecarlo_path_tracing_b200 as pt
from conftest import REFERENCE


def _image(h, w, seed=3):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    img[..., 3] = 255
    return img


def test_pam_round_trip(tmp_path):
    img = _image(37, 53)
    p = str(tmp_path / "a.ppm")
    pt.save_pam(p, img)
    back, maxval = pt.load_pam(p)
    assert maxval == 255 and back.dtype == np.uint8 and np.array_equal(back, img)


def test_pam_reader_variants(tmp_path):
    """3-channel files are padded to 4 values per pixel, 16-bit samples are big-endian, unknown header lines and
    TUPLTYPE are skipped (pamalign.h:52-131, 166-210)."""
    rng = np.random.default_rng(5)
    rgb = rng.integers(0, 256, (5, 7, 3), dtype=np.uint8)
    p = str(tmp_path / "rgb.pam")
    open(p, "wb").write(b"P7\n# a comment line\nWIDTH 7\nHEIGHT 5\nDEPTH 3\nMAXVAL 255\nTUPLTYPE RGB\nENDHDR\n" + rgb.tobytes())
    back, _ = pt.load_pam(p)
    assert back.shape == (5, 7, 4) and np.array_equal(back[..., :3], rgb) and (back[..., 3] == 0).all()
    g16 = rng.integers(0, 65536, (4, 6, 1), dtype=np.uint16)
    p = str(tmp_path / "g16.pam")
    open(p, "wb").write(b"P7\nWIDTH 6\nHEIGHT 4\nDEPTH 1\nMAXVAL 65535\nTUPLTYPE GRAYSCALE\nENDHDR\n" + g16.astype(">u2").tobytes())
    back, maxval = pt.load_pam(p)
    assert maxval == 65535 and back.dtype == np.uint16 and np.array_equal(back, g16)


def test_pam_reader_errors(tmp_path):
    with pytest.raises(pt.PtError):
        pt.load_pam(str(tmp_path / "missing.pam"))
    p = str(tmp_path / "p6.ppm")
    pt.save_ppm(p, _image(4, 4))
    with pytest.raises(pt.PtError):                         # "not a PAM file"
        pt.load_pam(p)
    p = str(tmp_path / "short.pam")
    open(p, "wb").write(b"P7\nWIDTH 4\nHEIGHT 4\nDEPTH 4\nMAXVAL 255\nTUPLTYPE RGB_ALPHA\nENDHDR\n" + b"\x01" * 10)
    with pytest.raises(pt.PtError):                         # truncated payload
        pt.load_pam(p)
    p = str(tmp_path / "hdr.pam")
    open(p, "wb").write(b"P7\nWIDTH 4\nENDHDR\n")
    with pytest.raises(pt.PtError):                         # "incomplete header"
        pt.load_pam(p)
    p = str(tmp_path / "five.pam")
    open(p, "wb").write(b"P7\nWIDTH 1\nHEIGHT 1\nDEPTH 5\nMAXVAL 255\nENDHDR\n" + b"\x00" * 5)
    with pytest.raises(pt.PtError):                         # "can't process PAM file with 5 channels"
        pt.load_pam(p)


def test_p6_ppm(tmp_path):
    img = _image(19, 23)
    p = str(tmp_path / "a.ppm")
    pt.save_ppm(p, img)
    raw = open(p, "rb").read()
    assert raw.startswith(b"P6\n23 19\n255\n")
    body = raw[len(b"P6\n23 19\n255\n"):]
    assert np.array_equal(np.frombuffer(body, np.uint8).reshape(19, 23, 3), img[..., :3])
    try:
        from PIL import Image
    except ImportError:
        return
    assert np.array_equal(np.asarray(Image.open(p)), img[..., :3])


def _decode_png(raw):
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, ihdr, types = 8, b"", None, []
    while pos < len(raw):
        n, = struct.unpack(">I", raw[pos:pos + 4])
        typ, data = raw[pos + 4:pos + 8], raw[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + data) & 0xFFFFFFFF == crc, typ
        types.append(typ)
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", data)
        elif typ == b"IDAT":
            idat += data
        pos += 12 + n
    assert types[0] == b"IHDR" and types[-1] == b"IEND"
    w, h, depth, ctype, comp, flt, inter = ihdr
    assert (depth, ctype, comp, flt, inter) == (8, 6, 0, 0, 0)
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 4 * w + 1)   # checks the Adler-32 too
    assert (rows[:, 0] == 0).all()
    return rows[:, 1:].reshape(h, w, 4)


@pytest.mark.parametrize("size", [(1, 1), (19, 23), (200, 333)])    # 200x333: several 65535-byte stored blocks
def test_png(tmp_path, size):
    img = _image(*size)
    img[..., 3] = np.random.default_rng(1).integers(0, 256, size, dtype=np.uint8)
    p = str(tmp_path / "a.png")
    pt.save_png(p, img)
    assert np.array_equal(_decode_png(open(p, "rb").read()), img)
    try:
        from PIL import Image
    except ImportError:
        return
    assert np.array_equal(np.asarray(Image.open(p).convert("RGBA")), img)


def test_obj_import_feeds_the_triangle_parser(tmp_path):
    obj = tmp_path / "cube.obj"
    obj.write_text("# unit cube, quads + one triangle with texture/normal indices and a negative index\n"
                   "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 1\nv 1 0 1\nv 1 1 1\nv 0 1 1\n"
                   "vn 0 0 1\nvt 0 0\n"
                   "f 1 2 3 4\nf 5/1/1 6/1/1 7/1/1 8/1/1\nf 1//1 2//1 6//1\nf -1 -2 -3\n")
    out = str(tmp_path / "triangles.txt")
    n = pt.import_obj(str(obj), out, scale=2.0, translate=(3.0, 1.0, 5.0))
    assert n == 2 + 2 + 1 + 1
    text = open(out).read()
    assert not text.endswith("\n")                          # a trailing newline would add a spurious triangle
    for name in ("spheres.txt", "squares.txt"):
        (tmp_path / name).write_text("\n".join(["0"] * 9))
    (tmp_path / "lights.txt").write_text("10\n4\n10\n200")
    sc = pt.load_scene_dir(str(tmp_path), "grid")
    assert sc.ntriangles == n
    t = sc.triangles.reshape(n, 3, 4)
    assert (t[..., 3] == 0).all()
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], np.float32) * 2 + np.array([3, 1, 5], np.float32)
    expect = [(0, 1, 2), (0, 2, 3), (4, 5, 6), (4, 6, 7), (0, 1, 5), (7, 6, 5)]
    for k, tri in enumerate(expect):
        assert np.array_equal(t[k, :, :3], v[list(tri)]), k
    assert np.allclose(sc.box_min[:3], [3, 1, 5]) and np.allclose(sc.box_max[:3], [5, 3, 7])
    with pytest.raises(pt.PtError):
        pt.import_obj(str(tmp_path / "nope.obj"), out)
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nf 1 2 3\n")                    # index out of range
    with pytest.raises(pt.PtError):
        pt.import_obj(str(bad), out)


@pytest.mark.needs_reference
@pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, "pamalign.h")), reason="needs /root/reference")
def test_against_the_reference_pamalign(tmp_path):
    """The reference's own load_pam / save_pam (pamalign.h, compiled from where it lies): it reads what we write
    and we read what it writes, byte for byte."""
    src = tmp_path / "harness.c"
    src.write_text('#include "pamalign.h"\n'
                   'int main(int argc, char **argv) {\n'
                   '    imgInfo img;\n'
                   '    if (load_pam(argv[1], &img)) return 1;\n'
                   '    printf("%u %u %u %u %u %zu\\n", img.width, img.height, img.channels, img.maxval, img.depth, img.data_size);\n'
                   '    FILE *f = fopen(argv[2], "wb"); fwrite(img.data, 1, img.data_size, f); fclose(f);\n'
                   '    return save_pam(argv[3], &img);\n'
                   '}\n')
    exe = str(tmp_path / "harness")
    subprocess.check_call(["gcc", "-w", "-I", REFERENCE, "-o", exe, str(src)])
    img = _image(41, 29)
    ours = str(tmp_path / "ours.ppm")
    pt.save_pam(ours, img)
    out = subprocess.check_output([exe, ours, str(tmp_path / "raw.bin"), str(tmp_path / "theirs.ppm")], text=True)
    assert out.split() == ["29", "41", "4", "255", "8", str(41 * 29 * 4)]
    assert open(tmp_path / "raw.bin", "rb").read() == img.tobytes()              # the reference reads our file
    assert open(tmp_path / "theirs.ppm", "rb").read() == open(ours, "rb").read()  # and rewrites it identically
    back, _ = pt.load_pam(str(tmp_path / "theirs.ppm"))                          # we read the reference's file
    assert np.array_equal(back, img)
