#!/usr/bin/env python3
"""The drop-in trianglegrid executable on the 1 M-triangle soup, from scene FILES (13 M text lines parsed by the C host):
PT_GPUS=1 vs PT_GPUS=n must write byte-identical result.ppm.  usage: cli_soup_check.py [ngpus] [spp] [width] [height]"""
import hashlib, os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scenes"))
import gen_mesh, write_scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
spp = sys.argv[2] if len(sys.argv) > 2 else "64"
Wd = sys.argv[3] if len(sys.argv) > 3 else "1920"
Ht = sys.argv[4] if len(sys.argv) > 4 else "1080"
exe = os.path.join(ROOT, "opencl_montecarlo_path_tracing_b200", "bin", "CLSuperPathTracer_trianglegrid", "CLSuperPathTracer")
d = tempfile.mkdtemp()
write_scenes.write_variant("grid", d)
t0 = time.time()
tris = gen_mesh.soup(1 << 20)
write_scenes.write_triangles(os.path.join(d, "triangles.txt"), tris[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]])
print("scene written in %.1f s (%.0f MB)" % (time.time() - t0, os.path.getsize(os.path.join(d, "triangles.txt")) / 1e6), flush=True)
digests = {}
for g in ((1, n) if n > 1 else (1,)):
    env = dict(os.environ, PT_SEEDS="1,2,3,4", PT_GPUS=str(g), PT_SPP=spp, PT_MAX_TRIANGLES=str(1 << 20), PT_STATS="1", PT_TIMING="1")
    t0 = time.time()
    p = subprocess.run([exe, Wd, Ht], cwd=d, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    print("".join(l + "\n" for l in p.stderr.splitlines() if l.startswith("PT_TIMING")), end="")
    raw = open(os.path.join(d, "result.ppm"), "rb").read()
    digests[g] = hashlib.sha256(raw).hexdigest()
    print("PT_GPUS=%d wall %.1f s |" % (g, time.time() - t0), re.search(r"Number of triangles: \d+", p.stdout).group(0), "|",
          re.search(r"Triangles grid size: .*", p.stdout).group(0), "|", re.search(r"init triangles grid : .*", p.stdout).group(0), "|",
          re.search(r"rendering : .*", p.stdout).group(0), flush=True)
print("byte-identical:", len(set(digests.values())) == 1, digests)
