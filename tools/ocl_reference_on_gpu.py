#!/usr/bin/env python3
"""Runs the UNMODIFIED reference programs against NVIDIA's real OpenCL runtime on the B200 (the binaries built by
`make -C oracle ref` into oracle/_ref/ocl/, kernel text embedded) and compares, for the same fixed seeds,
  * their result.ppm with this repo's CUDA output and with the CPU oracle (both arithmetic policies),
  * their kernel time (OpenCL event profiling, as printed by the reference) with the CUDA kernel time.
TEST/BENCH INFRASTRUCTURE.  Writes gpurun_out/ocl_vs_cuda.json.
"""
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scenes"))
import numpy as np  # noqa: E402
import opencl_montecarlo_path_tracing_b200 as pt  # noqa: E402
import write_scenes  # noqa: E402
from oracle.pyoracle import OracleLib  # noqa: E402

SEEDS = (1, 2, 3, 4)
W = H = int(os.environ.get("OCL_SIZE", "512"))


def read_pam(path):
    raw = open(path, "rb").read()
    k = raw.index(b"ENDHDR\n") + 7
    return np.frombuffer(raw[k:], np.uint8).reshape(H, W, 4)


def stats(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))[..., :3]
    return {"identical_channels": float((d == 0).mean()), "within_1_lsb": float((d <= 1).mean()), "max_abs": int(d.max()),
            "rmse_lsb": float(np.sqrt((d.astype(np.float64) ** 2).mean())), "identical_pixels": float((d.max(axis=2) == 0).mean())}


def main():
    out = {"size": [W, H], "seeds": list(SEEDS), "variants": {}}
    env = dict(os.environ, OCL_ICD_FILENAMES="libnvidia-opencl.so.1", PT_SEEDS=",".join(map(str, SEEDS)))
    o0, o1 = OracleLib(0), OracleLib(1)
    with pt.Renderer(0) as r, tempfile.TemporaryDirectory() as tmp:
        for v in sys.argv[1:] or ["base", "lmem", "nodof", "grid"]:
            d = os.path.join(tmp, v)
            write_scenes.write_variant(v, d)
            exe = os.path.join(ROOT, "oracle", "_ref", "ocl", v, "CLSuperBidirectionalPathTracer" if v == "bidir" else "CLSuperPathTracer")
            best = None
            p = None
            for it in range(3):
                p = subprocess.run([exe, str(W), str(H)], cwd=d, env=env, capture_output=True, text=True, timeout=600)
                if p.returncode != 0:
                    out["variants"][v] = {"error": (p.stdout + p.stderr)[-2000:]}
                    print(v, "FAILED", (p.stdout + p.stderr)[-1500:], flush=True)
                    best = None
                    break
                ms = sum(float(x) for x in re.findall(r"(?:rendering|reduce img samples|virtual light sampling) : .*? in ([0-9.eE+-]+)ms", p.stdout))
                best = ms if best is None else min(best, ms)
            if best is None:
                continue
            log = p.stdout
            ocl_img = read_pam(os.path.join(d, "result.ppm")).copy()
            scene = pt.load_scene_dir(d, v)
            r.set_scene(scene)
            if v == "grid":
                r.build_grid(pt.grid_dims(scene))
            cuda_ms = 1e9
            for it in range(5):
                light_ms = r.light_tracer(SEEDS, 512) if v == "bidir" else 0.0    # same sum as the OpenCL figure above
                res = r.render(v, W, H, SEEDS)
                cuda_ms = min(cuda_ms, res.ms + light_ms)
            sc = o0.load_scene_dir(d, v)
            or0 = o0.render(v, W, H, SEEDS, sc, want_accum=False, want_rng=False)["image"]
            or1 = o1.render(v, W, H, SEEDS, sc, want_accum=False, want_rng=False)["image"]
            rays = res.counters["rays"]
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ocl_img_%s.npz" % v), opencl=ocl_img, cuda=res.image)
            blog = re.search(r"=== BUILD LOG ===\n(.*?)\n=========", log, re.S)
            dev = re.search(r"selected device \d+: (.*)", log)
            out["variants"][v] = {
                "opencl_kernel_ms": best, "cuda_kernel_ms": cuda_ms, "speedup": best / cuda_ms, "rays": rays,
                "opencl_mrays_per_s": rays / 1e3 / best, "cuda_mrays_per_s": rays / 1e3 / cuda_ms,
                "cuda_vs_opencl": stats(res.image, ocl_img), "oracle_fma_vs_opencl": stats(or1, ocl_img),
                "oracle_separate_vs_opencl": stats(or0, ocl_img), "cuda_equals_oracle_fma": bool(np.array_equal(res.image, or1)),
                "opencl_build_log": blog.group(1)[:300] if blog else None, "opencl_device": dev.group(1) if dev else None,
            }
            print(v, json.dumps(out["variants"][v]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ocl_vs_cuda.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
