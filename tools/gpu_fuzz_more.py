"""One-off extended GPU fuzz (not part of the -m gpu suite): the suite's generated-pattern test with more seeds, once in the
default flow and once with the state-map scan forced for every buffer search (FX_STATEMAP=2), for a bounded time."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import test_gpu_parity as T  # noqa: E402


class Env:
    def setenv(self, k, v):
        os.environ[k] = v


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 150.0
    t0 = time.time()
    done = []
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    while time.time() - t0 < budget:
        os.environ["FX_STATEMAP"] = "2" if seed % 2 else "1"
        print("seed", seed, "FX_STATEMAP", os.environ["FX_STATEMAP"], flush=True)
        T.test_generated_patterns_on_gpu(seed, Env())
        done.append((seed, os.environ["FX_STATEMAP"]))
        seed += 1
    print("extended fuzz ok:", len(done), "seeds", done[0], "..", done[-1], "in %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
