"""One-off fuzz of the split-buffer search (not part of the suites): random patterns over texts of ~1 MB cut into one
slab per GPU with a SHORT halo (so that attempts, literal occurrences and winners run off a rank's window and the halo
has to be widened over P2P), forgex_b200.dist.gpu_buffer_search against the single-GPU fx_regex_buffer answer.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_fuzz.py SECONDS
"""
import json
import os
import random
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import forgex_b200 as fx  # noqa: E402
from forgex_b200 import dist as fxd  # noqa: E402
from forgex_b200 import _lib as L  # noqa: E402
from tests.test_host_tables import gen_pattern, gen_text  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    seed, tried, declined, widened, bad = 0, 0, 0, 0, []
    t0 = time.time()
    go = torch.ones(1, dtype=torch.int64, device="cuda")
    while True:
        if rank == 0:
            go[0] = 1 if time.time() - t0 < budget else 0
        dist.broadcast(go, 0)
        if int(go.item()) == 0:
            break
        rng = random.Random(55000 + seed)                     # every rank builds the same text and patterns
        sep = (b"\n", b" ", b"")[seed % 3]
        text = np.frombuffer(sep.join(gen_text(rng) for _ in range(rng.choice((20000, 120000)))), dtype=np.uint8)
        nbytes = len(text)
        lo, hi = fxd.slab_bounds(nbytes, world, rank)
        halo = rng.choice((16, 64, 4096))
        w_lo, w_hi = fxd.window_for_slab(nbytes, lo, hi, halo)
        d_win = torch.from_numpy(text[w_lo:w_hi].copy()).cuda()
        for _ in range(10):
            pat = gen_pattern(rng).encode()
            p = fx.Pattern(pat, "regex")
            if p.status != 0:
                continue
            stats = {}
            try:
                f, t, undecided = fxd.gpu_buffer_search(p, d_win, w_lo, nbytes, rank, world, (lo, hi), stats=stats)
            except fx.ForgexError as e:
                assert e.status in (L.FX_ERR_PREFILTER_UNSUPPORTED, L.FX_ERR_DFA_STATE_CAP), (pat, e.status)
                declined += 1                                 # sequential candidate list / NFA engine: stated limits of the window forms
                continue
            tried += 1
            widened += 1 if stats.get("widenings", 0) > 0 else 0
            if rank == 0:
                exp = p.regex_buffer(text)
                if (f, t) != tuple(exp) or undecided != 0:
                    bad.append((seed, pat.decode("utf-8", "replace"), [f, t], list(exp), undecided, nbytes, halo))
        seed += 1
    if rank == 0:
        print(json.dumps({"n_gpus": world, "seeds": seed, "searches": tried, "declined_by_the_window_forms": declined,
                          "searches_that_widened_their_halo": widened, "mismatches": bad[:5], "all_equal": not bad}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
