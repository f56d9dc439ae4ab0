"""Multi-GPU host logic: one process per GPU (torch.distributed), shard the work, exchange only tiny results.

The path shards naturally: strings of a batch are independent, and the start positions of a buffer
search are independent attempts.  No text ever crosses GPUs.

  batches   contiguous string ranges balanced by bytes (shard_strings); results stay sharded; one all-reduce
            of the per-rank match counts (8 bytes per rank).
  buffer    contiguous byte slabs plus a read-only halo behind each slab (attempts started in the slab may run
            into it) and 3 bytes in front (character-boundary look-back); each rank finds the smallest winning
            start of its slab; one all-reduce(MIN) of that 8-byte key; the owner of the winner computes the span;
            one broadcast of 16 bytes.

Everything here is host logic over a `scan` / `finish` pair, so it runs unchanged on CPU with the gloo backend
(tests/test_dist_gloo.py plugs a Python model of the kernels in) and on GPUs with NCCL (plugs the C ABI in).
"""
import numpy as np

NO_START = (1 << 64) - 1


def shard_strings(offsets, world, rank):
    """contiguous range [first, last) of strings for `rank`, balanced by bytes (offsets: n+1 ascending int64)"""
    offsets = np.asarray(offsets)
    n = len(offsets) - 1
    total = int(offsets[-1] - offsets[0])
    cuts = [0]
    for r in range(1, world):
        target = offsets[0] + (total * r) // world
        cuts.append(int(np.searchsorted(offsets[:n], target, side="left")))
    cuts.append(n)
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts[rank], cuts[rank + 1]


def slab_bounds(length, world, rank, align=16):
    """contiguous byte slab [lo, hi) of a buffer of `length` bytes for `rank` (cut at multiples of `align`)"""
    per = -(-length // world)
    per = -(-per // align) * align
    lo = min(length, rank * per)
    hi = min(length, (rank + 1) * per)
    return lo, hi


def window_for_slab(length, lo, hi, halo):
    """the bytes a rank must hold to scan the starts [lo, hi): 3 bytes of look-back, `halo` bytes of look-ahead"""
    w_lo = max(0, lo - 3)
    w_hi = min(length, hi + halo)
    return w_lo, w_hi


def all_reduce_int(value, op, group=None, device=None):
    """all-reduce of one integer (uint64 keys are shifted into int64 order for MIN)"""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=op, group=group)
    return int(t.item())


def count_matches(local_count, group=None, device=None):
    import torch.distributed as dist
    return all_reduce_int(int(local_count), dist.ReduceOp.SUM, group, device)


def buffer_search(scan, finish, length, rank, world, slab, window, group=None, device=None, scan_all=None):
    """Leftmost-longest search of one pattern in one buffer that is split across ranks.

    scan(start_lo, start_hi) -> (key, undecided[, occurrences]): smallest winning start of this rank's slab as a
        1-based position in NUL||text||NUL of the whole text (NO_START if none), the number of attempts that ran off
        the window and -- for a pattern with a prefix literal -- how many occurrences of the literal the slab holds.
    scan_all(start_lo, start_hi) -> (key, undecided): the same over EVERY character boundary.  Given for a pattern
        with a prefix literal: Forgex takes its candidate starts from the literal's occurrences, and from every
        boundary when the literal occurs nowhere in the text (api_internal_m.F90:76-104) -- "nowhere" is a property
        of the whole text, hence one more 8-byte all-reduce.
    finish(key) -> (from, to) computed by the rank whose window holds the winner (or (-1, -1) if its window is
        too short).
    Returns (from, to, undecided_total); every rank gets the same answer.
    """
    import torch
    import torch.distributed as dist
    lo, hi = slab
    active = hi > lo or (rank == 0 and length == 0)
    res = scan(lo, hi) if active else (NO_START, 0, 0)
    key, undecided = res[0], res[1]
    if scan_all is not None:
        occurrences = all_reduce_int(res[2] if len(res) > 2 else 0, dist.ReduceOp.SUM, group, device)
        if occurrences == 0:
            key, undecided = scan_all(lo, hi) if active else (NO_START, 0)
    # uint64 MIN through int64: keys are < 2**63 except NO_START, which maps to the largest int64
    k64 = key if key != NO_START else (1 << 63) - 1
    best = all_reduce_int(k64, dist.ReduceOp.MIN, group, device)
    undecided_total = all_reduce_int(undecided, dist.ReduceOp.SUM, group, device)
    span = torch.zeros(2, dtype=torch.int64, device=device)
    owner = torch.tensor([-1], dtype=torch.int64, device=device)
    if best != (1 << 63) - 1:
        pos = best - 2                       # text index of the winning start (-1: the leading NUL)
        mine = (lo <= pos < hi) or (pos < 0 and rank == 0)
        if mine:
            f, t = finish(best)
            span[0], span[1] = f, t
            owner[0] = rank
        if dist.is_initialized() and world > 1:
            dist.all_reduce(owner, op=dist.ReduceOp.MAX, group=group)
            dist.broadcast(span, src=int(owner.item()), group=group)
    return int(span[0].item()), int(span[1].item()), undecided_total


def gpu_buffer_search(pattern_obj, d_window, window_origin, length, rank, world, slab, group=None):
    """buffer_search over the C ABI: d_window is this rank's CUDA tensor holding text[window_origin : ...]"""
    import torch
    dev = d_window.device
    wlen = d_window.numel()
    is_first = window_origin == 0
    is_last = window_origin + wlen == length
    best = torch.empty(3, dtype=torch.int64, device=dev)
    ft = torch.zeros(2, dtype=torch.int64, device=dev)
    prefixed = bool(pattern_obj.info()["prefix_scan"])

    def run(lo, hi, every_boundary):
        best[0] = -1          # all ones
        best[1] = 0
        best[2] = 0
        call = pattern_obj.buffer_scan_all_dev if every_boundary else pattern_obj.buffer_scan_dev
        call(d_window, wlen, lo - window_origin, hi - window_origin, window_origin, is_first, is_last, best)
        b = best.cpu().numpy().view(np.uint64)
        return int(b[0]), int(b[1]), int(b[2])

    def scan(lo, hi):
        return run(lo, hi, False)

    def scan_all(lo, hi):
        return run(lo, hi, True)[:2]

    def finish(key):
        k = torch.tensor([key], dtype=torch.int64, device=dev)
        pattern_obj.buffer_finish_dev(d_window, wlen, window_origin, is_last, k, ft)
        r = ft.cpu().numpy()
        return int(r[0]), int(r[1])

    return buffer_search(scan, finish, length, rank, world, slab, None, group, dev, scan_all if prefixed else None)
