"""One buffer split across the GPUs of a box: forgex_b200.dist.gpu_buffer_search over NCCL, checked against the
single-GPU fx_regex_buffer answer.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/dist_buffer_demo.py

Every rank builds the same seeded text, keeps only its slab (+ look-back / halo) on its GPU, scans its slab's starts
(fx_buffer_scan_dev), and the ranks agree on the winner with 8-byte all-reduces (MIN of the start key; SUM of the prefix
occurrences for a pattern with a prefix literal); the owner of the winning start computes the span.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import forgex_b200 as fx  # noqa: E402
from forgex_b200 import dist as fxd  # noqa: E402
from tools import synth  # noqa: E402


def ascii_text(nbytes, plant, at, seed):
    r = np.random.default_rng(seed)
    t = r.integers(0x20, 0x7F, size=nbytes, dtype=np.uint8)
    t[t == ord("f")] = ord("g")                       # no accidental `foo`
    if plant is not None:
        k = int(nbytes * at)
        t[k:k + len(plant)] = np.frombuffer(plant, dtype=np.uint8)
    return t


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [("c4 match at 70%", synth.PATTERNS["c4"], lambda: synth.gen_c4(16 << 20, 0.7)),
             ("c4 no match", synth.PATTERNS["c4"], lambda: synth.gen_c4(16 << 20, None)),
             ("prefix literal, match at 40%", b"foo(bar|baz)", lambda: ascii_text(48 << 20, b"xx foobaz yy", 0.4, 1)),
             ("prefix literal, literal in rank 0 only", b"foo(bar|baz)", lambda: ascii_text(48 << 20, b"fooba!", 0.1, 2)),
             ("prefix literal nowhere, overlong start", b"foo(bar|baz)", lambda: ascii_text(48 << 20, b"\xc1\xa6oobar", 0.8, 3)),
             ("prefix literal nowhere, no match", b"foo(bar|baz)", lambda: ascii_text(48 << 20, None, 0, 4))]
    out = []
    for name, pat, make in cases:
        text = make()
        nbytes = len(text)
        lo, hi = fxd.slab_bounds(nbytes, world, rank)
        w_lo, w_hi = fxd.window_for_slab(nbytes, lo, hi, 4096)
        d_win = torch.from_numpy(text[w_lo:w_hi].copy()).cuda()
        p = fx.Pattern(pat, "regex")
        fxd.gpu_buffer_search(p, d_win, w_lo, nbytes, rank, world, (lo, hi))     # warm-up
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        f, t, undecided = fxd.gpu_buffer_search(p, d_win, w_lo, nbytes, rank, world, (lo, hi))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            exp = fx.Pattern(pat, "regex").regex_buffer(text)          # the whole text on one GPU
            out.append({"case": name, "bytes": nbytes, "span": [f, t], "single_gpu": list(exp), "equal": (f, t) == tuple(exp),
                        "undecided": undecided, "ms": round(dt * 1000, 3)})
    if rank == 0:
        print(json.dumps({"n_gpus": world, "cases": out, "all_equal": all(c["equal"] and c["undecided"] == 0 for c in out)}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
