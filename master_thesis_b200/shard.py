"""Partitioning of the hot path across ranks (one process per GPU).

SURVEY.md 8(e): warp / v_map / mask_out / correlation / CHN pack + composite act
on each (b, f) pair independently -> contiguous blocks of the flattened B x F
range; CM_Module couples the references of one sample (softmax over refs,
model_cpn.py:237) -> shard by b only.  No data-path collective exists; the only
collective of the reference is DDP's gradient all-reduce (NCCL), which the
kernels do not touch.
"""


def block_range(n, rank, world):
    """Contiguous block [lo, hi) of range(n) owned by ``rank`` (sizes differ by <= 1)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world: %r/%r" % (rank, world))
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def frame_shard(b, f, rank, world):
    """Frames (b_i, f_i) of the flattened B x F range owned by ``rank``."""
    lo, hi = block_range(b * f, rank, world)
    return [(n // f, n % f) for n in range(lo, hi)]


def frame_shard_groups(b, f, rank, world):
    """Same shard grouped per sample: {b_i: [f_lo, f_hi)} - each group is one kernel
    launch over a (1, C, f_hi - f_lo, H, W) view, the target of sample b_i replicated."""
    groups = {}
    for bi, fi in frame_shard(b, f, rank, world):
        lo, hi = groups.get(bi, (fi, fi))
        groups[bi] = (min(lo, fi), max(hi, fi + 1))
    return groups


def batch_shard(b, rank, world):
    """Samples owned by ``rank`` for CM_Module (shard by b only)."""
    return block_range(b, rank, world)
