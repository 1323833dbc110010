"""Runs the PARSeq engine a few times on synthetic crops (for ncu launch lists / captures).
usage: run_parseq_once.py [n_crops] [width] [reps] [peaked(0/1)]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import parseq as ops  # noqa: E402
from oracle import weights  # noqa: E402
from yomitoku_b200 import TextRecognizer  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
W = int(sys.argv[2]) if len(sys.argv) > 2 else 184
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
peaked = int(sys.argv[4]) if len(sys.argv) > 4 else 1
name = "parseq-large-v4_1"
rec = TextRecognizer(model_name=name, from_pretrained=False, device="cuda", dynamic_width=True, batch_bucketing=True)
if peaked:
    rec.model.load_state_dict(weights.make_parseq_state_dict(ops.SPECS[name], seed=4, peaked=True))
rng = np.random.default_rng(0)
canv = [rng.integers(0, 256, size=(32, W, 3), dtype=np.uint8) for _ in range(n)]
groups = [i // 128 for i in range(n)]
for _ in range(reps):
    ids, probs, glen = rec.model.recognize_crops(canv, [W] * n, groups, groups[-1] + 1)
    torch.cuda.synchronize()
    print("steps", glen.tolist()[:4], rec.model.last_phase_ms(), "gflop", rec.model.last_flops() / 1e9)
