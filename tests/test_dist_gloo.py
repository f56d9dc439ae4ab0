"""world_size-2 (and 3) CPU tests of the multi-GPU host logic over the gloo backend.

The kernels are replaced by tests/table_model.py (a Python model of what they compute on the product's own tables),
so what is under test is the sharding, the halo handling and the tiny collectives of forgex_b200/dist.py.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import forgex_b200 as fx
from forgex_b200 import dist as fxd
from tests import oracle_lib as O
from tests.table_model import Anchored, Model
from tools import synth


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def model_scan(anch, text, window_lo, window_hi, lo, hi, length):
    """what fx_buffer_scan_dev computes: smallest winning start in [lo, hi), attempts confined to the window"""
    wtext = text[window_lo:window_hi]
    is_first, is_last = window_lo == 0, window_hi == length
    best, undecided = fxd.NO_START, 0
    t = anch.t
    if is_first and lo == 0 and t["start_nul"] != 0 and is_last:
        if anch.attempt_at(wtext, 1) >= 0:
            return 1, 0
    for pos in range(lo, hi):
        rel = pos - window_lo
        b = wtext[rel]
        if (anch.nxt(t["q0"], b) & 0x3FFF) == 0:
            continue
        if (b & 0xC0) == 0x80:
            # character boundary under the sequential strict decoder: decode from a known boundary a few bytes back
            q = max(0, rel - 3)
            while q > 0 and (wtext[q] & 0xC0) == 0x80:
                q -= 1
            p = q
            while p < rel:
                p += anch.char_len(wtext, p)
            if p != rel:
                continue
        if is_last:
            ok = anch.attempt(wtext, t["q0"], rel, -1) >= 0
        else:
            # open end: walk only the bytes of the window; alive at its end without an accept = undecided
            st, last, alive = t["q0"], -1, True
            for j in range(rel, len(wtext)):
                st = anch.nxt(st & 0x3FFF, wtext[j])
                if st & 0x8000:
                    last = j + 1
                if (st & 0x3FFF) == 0:
                    alive = False
                    break
            ok = last >= 0
            if alive and not ok:
                undecided += 1
        if ok:
            best = pos + 2
            break
    return best, undecided


def model_scan_prefix(anch, pre, text, window_lo, window_hi, lo, hi, length):
    """fx_buffer_scan_dev for a pattern with a prefix literal: candidates = the literal's occurrences that start in
    [lo, hi); returns (key, undecided, occurrences)"""
    wtext = text[window_lo:window_hi]
    is_last = window_hi == length
    best, undecided, occ = fxd.NO_START, 0, 0
    t = anch.t
    for pos in range(lo, hi):
        rel = pos - window_lo
        if rel + len(pre) > len(wtext):
            if not is_last and wtext[rel:] == pre[:len(wtext) - rel]:
                undecided += 1
            continue
        if wtext[rel:rel + len(pre)] != pre:
            continue
        occ += 1
        if pos == 0 and t["start_nul"] != 0 and is_last and anch.attempt_at(wtext, 1) >= 0:
            best = min(best, 1)
        if best != fxd.NO_START:
            continue
        assert is_last or rel + 64 < len(wtext)          # (test texts keep matches short: no open-end bookkeeping here)
        if anch.attempt(wtext, t["q0"], rel, -1) >= 0:
            best = pos + 2
    return best, undecided, occ


def worker(rank, world, port, cfg, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if cfg["kind"] == "batch":
            buf, off = synth.gen_c2(cfg["n"])
            first, last = fxd.shard_strings(off, world, rank)
            p = fx.Pattern(synth.PATTERNS["c2"], "in")
            m = Model(p)
            local = [m.boolean(bytes(buf[off[i]:off[i + 1]])) for i in range(first, last)]
            total = fxd.count_matches(sum(local))
            ranges = [None] * world
            dist.all_gather_object(ranges, (first, last))
            ret[rank] = (total, ranges, local)
        else:
            text = cfg["text"]
            length = len(text)
            p = fx.Pattern(cfg["pattern"], "regex")
            anch = Anchored(p, True)
            lo, hi = fxd.slab_bounds(length, world, rank, align=cfg.get("align", 16))
            w_lo, w_hi = fxd.window_for_slab(length, lo, hi, cfg["halo"])

            win = None
            if cfg.get("widen"):      # the rank keeps its window as a tensor; a too-short halo grows by P2P reads from the successors
                win = fxd._GpuWindow(torch.from_numpy(np.frombuffer(text, dtype=np.uint8)[w_lo:w_hi].copy()), w_lo, length,
                                     rank, world, (lo, hi), None)

            def scan(a, b):
                if win is not None:
                    held = bytes(win.t.numpy())
                    assert held == text[w_lo:win.end()]           # what arrived over P2P is the text itself
                    return model_scan(anch, text, w_lo, win.end(), a, b, length)
                return model_scan(anch, text, w_lo, w_hi, a, b, length)

            def finish(key):
                f, t = anch.regex(text)   # the owner holds the match in its window; same attempt, whole-text coordinates
                return f, t
            if cfg.get("prefixed"):
                assert p.info()["prefix_scan"] == 1
                pre = p.literals()[1]

                def scan_pre(a, b):
                    return model_scan_prefix(anch, pre, text, w_lo, w_hi, a, b, length)
                ret[rank] = fxd.buffer_search(scan_pre, finish, length, rank, world, (lo, hi), (w_lo, w_hi), scan_all=scan)
            else:
                stats = {}
                r = fxd.buffer_search(scan, finish, length, rank, world, (lo, hi), (w_lo, w_hi),
                                      widen=win.widen if win is not None else None, stats=stats)
                ret[rank] = r + (stats.get("widenings", 0),) if win is not None else r
    finally:
        dist.destroy_process_group()


def run(world, cfg):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(worker, args=(world, free_port(), cfg, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_batch_counts(world):
    n = 600
    res = run(world, {"kind": "batch", "n": n})
    buf, off = synth.gen_c2(n)
    exp = O.Compiled(synth.PATTERNS["c2"], 0).bool_batch(0, buf, off)
    total, ranges, _ = res[0]
    assert total == int(exp.sum())
    assert all(r[0] == total for r in res)
    # the shards tile [0, n) without gaps or overlaps and are balanced by bytes
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    sizes = [off[b] - off[a] for a, b in ranges]
    assert max(sizes) - min(sizes) <= 2 * 256
    got = np.concatenate([np.array(r[2], dtype=np.uint8) for r in res])
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("world,match_at", [(2, 0.1), (2, 0.9), (3, 0.5), (2, None)])
def test_sharded_buffer_search(world, match_at):
    text = bytes(synth.gen_c4(6000, match_at))
    res = run(world, {"kind": "buffer", "text": text, "pattern": synth.PATTERNS["c4"], "halo": 512})
    exp = O.Compiled(synth.PATTERNS["c4"], 0).regex_buffer(np.frombuffer(text, dtype=np.uint8))
    for r in res:
        assert (r[0], r[1]) == exp
        assert r[2] == 0


def test_sharded_buffer_match_across_the_cut():
    """a match that starts in rank 0's slab and ends in rank 1's is found by rank 0 through its halo"""
    line = synth.C4_MATCH_LINE
    text = b"INFO x\n" * 30 + line + b"\n" + b"INFO y\n" * 30
    cut = len(text) // 2
    assert text.find(line) < cut < text.find(line) + len(line)
    res = run(2, {"kind": "buffer", "text": text, "pattern": synth.PATTERNS["c4"], "halo": 256, "align": 1})
    exp = O.Compiled(synth.PATTERNS["c4"], 0).regex_buffer(np.frombuffer(text, dtype=np.uint8))
    assert exp[0] > 0
    for r in res:
        assert (r[0], r[1]) == exp


def test_short_halo_is_reported():
    line = synth.C4_MATCH_LINE
    text = b"INFO x\n" * 30 + line + b"\n" + b"INFO y\n" * 30
    res = run(2, {"kind": "buffer", "text": text, "pattern": synth.PATTERNS["c4"], "halo": 8, "align": 1})
    assert res[0][2] >= 1   # an attempt ran off rank 0's window: the caller must widen the halo


@pytest.mark.parametrize("world", [2, 3])
def test_short_halo_is_widened_over_p2p(world):
    """with a widen callback the undecided attempt is resolved: the halo doubles (bytes sent by the ranks that own them)
    until the attempt that ran off the window can be decided"""
    line = synth.C4_MATCH_LINE
    text = b"INFO x\n" * (30 * (world - 1)) + line + b"\n" + b"INFO y\n" * 30
    res = run(world, {"kind": "buffer", "text": text, "pattern": synth.PATTERNS["c4"], "halo": 8, "align": 1, "widen": True})
    exp = O.Compiled(synth.PATTERNS["c4"], 0).regex_buffer(np.frombuffer(text, dtype=np.uint8))
    assert exp[0] > 0
    for r in res:
        assert (r[0], r[1]) == exp and r[2] == 0 and r[3] >= 1


@pytest.mark.parametrize("world,case", [(2, 0), (2, 1), (3, 2), (2, 3), (2, 4)])
def test_sharded_buffer_search_with_prefix_literal(world, case):
    """candidates are the prefix literal's occurrences; "occurs nowhere" is decided over ALL ranks (one more 8-byte
    all-reduce) before the search falls back to every boundary"""
    import random
    rng = random.Random(5)
    filler = bytes(rng.choice(b"abcdeghijklmnpqrstuvwxyz .,;:-_0123456789") for _ in range(3000))
    text = [filler + b"foobaz" + filler,                           # one occurrence, in the middle
            filler + filler + b"foobar",                            # only the last rank sees the literal
            filler + b"\xc1\xa6oobar" + filler,                     # literal nowhere: the overlong start wins by brute force
            b"\xc1\xa6oobar" + filler + b"fooba " + filler,         # literal somewhere: the overlong start is never tried
            filler + filler][case]
    res = run(world, {"kind": "buffer", "text": text, "pattern": b"foo(bar|baz)", "halo": 512, "prefixed": True})
    exp = O.Compiled(b"foo(bar|baz)", 0).regex_buffer(np.frombuffer(text, dtype=np.uint8))
    assert (exp[0] > 0) == (case in (0, 1, 2))
    for r in res:
        assert (r[0], r[1]) == exp
        assert r[2] == 0
