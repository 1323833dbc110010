"""Writes tests/golden/identity_<model>.npz: the fp32 oracle's outputs for the 2 x 2048 crops of
tests/test_gpu_parseq_identity.py (ids, max-probabilities, top-2 margins of the final logits, smallest AR-decision margin
per row), so that the GPU test compares against them instead of spending minutes of host time per run.

    python tests/golden/make_golden_identity.py            # both models (about 10 minutes on 8 cores)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_gpu_parseq_identity as T  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    for case in T.CASES:
        name = case[0]
        if len(sys.argv) > 1 and sys.argv[1] not in name:
            continue
        ids, prob, margin, armin = T.oracle_all(*case)
        path = os.path.join(T.GOLDEN, "identity_%s.npz" % name)
        np.savez_compressed(path, case=np.asarray(case[1:], dtype=np.int64), ids=ids.astype(np.int16),
                            prob=prob.astype(np.float32), margin=margin.astype(np.float32),
                            ar_margin_min=armin.astype(np.float32))
        print(name, ids.shape, "->", path, os.path.getsize(path), "bytes", flush=True)


if __name__ == "__main__":
    main()
