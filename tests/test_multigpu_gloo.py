"""N > 1 path on CPU: world_size-2 gloo.  Each rank renders its row stripes (here with the oracle standing
in for the GPU — this test is about the sharding + reduce logic, not the kernels), a SUM reduce assembles
the accumulation buffer on rank 0, and the result must be bit-identical to the single-process frame."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, scene_dir, out_path, variant):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from opencl_montecarlo_path_tracing_b200 import sharding
    from oracle.pyoracle import OracleLib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = OracleLib(0)
    sc = o.load_scene_dir(scene_dir, variant)
    # bidir: every rank traces the light pass itself (same seeds -> the same VPL buffer on every rank, no exchange)
    extra = {"vpls": o.light_tracer((1, 2, 3, 4), sc, 512)} if variant == "bidir" else {}
    W, H, stripe = 128, 64, 8
    acc = np.zeros((H, W, 4), np.float32)
    rows = sharding.stripe_rows(H, stripe, rank, world)
    # contiguous runs of owned rows -> one oracle call each
    runs = np.split(rows, np.where(np.diff(rows) != 1)[0] + 1)
    for run in runs:
        if len(run):
            part = o.render(variant, W, H, (1, 2, 3, 4), sc, rows=(int(run[0]), int(run[-1]) + 1), want_rng=False, nthreads=2, **extra)
            acc[run[0]:run[-1] + 1] = part["accum"][run[0]:run[-1] + 1]
    t = torch.from_numpy(acc)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("variant", ["lmem", "bidir"])
def test_two_rank_stripes_reduce_to_the_full_frame(scene_dirs, tmp_path, oracle_sep, variant):
    import torch.multiprocessing as mp
    out = str(tmp_path / "acc.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, scene_dirs[variant], out, variant), nprocs=2, join=True)
    got = np.load(out)
    full = oracle_sep.render(variant, 128, 64, (1, 2, 3, 4), oracle_sep.load_scene_dir(scene_dirs[variant], variant), want_rng=False)
    assert np.array_equal(got.view(np.uint32), full["accum"].view(np.uint32))
