"""Row-stripe sharding of an image over GPUs: the host-side statement of the mapping the
kernels implement in pt::map_row (csrc/pt_host.h).

Stripes of `stripe` rows are dealt round-robin to the ranks; every work-item still seeds its RNG from its
GLOBAL pixel id, so the union of the ranks' stripes is bit-identical to a single-GPU render.  Each rank
renders into a zeroed float accumulation buffer; one sum-reduce (NCCL on GPUs) assembles the frame.
"""
import numpy as np


def stripe_rows(height, stripe, rank, nranks, row_begin=0, row_end=None):
    """Image rows owned by `rank` (ascending)."""
    row_end = height if row_end is None else min(row_end, height)
    rows = np.arange(row_begin, row_end)
    if nranks <= 1 or stripe <= 0:
        return rows
    s = (rows - row_begin) // stripe
    return rows[s % nranks == rank]


def virtual_rows(height, stripe, rank, nranks, row_begin=0, row_end=None):
    """Number of virtual rows the kernel walks for this rank (LaunchArgs::nrows in csrc/ptcuda.cu)."""
    row_end = height if row_end is None else min(row_end, height)
    r = row_end - row_begin
    if nranks <= 1 or stripe <= 0:
        return r
    nstripes = (r + stripe - 1) // stripe
    mine = max(0, (nstripes - rank + nranks - 1) // nranks)
    return mine * stripe


def map_row(vr, stripe, rank, nranks, row_begin=0):
    """virtual row -> image row, as pt::map_row."""
    if nranks <= 1 or stripe <= 0:
        return row_begin + vr
    s, o = divmod(vr, stripe)
    return row_begin + (s * nranks + rank) * stripe + o
