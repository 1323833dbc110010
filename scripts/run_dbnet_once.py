"""Runs the DBNet engine a few times on one synthetic 1200x1600 page (for ncu launch lists / captures)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import weights  # noqa: E402
from yomitoku_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L = _lib.lib()
sd = weights.make_dbnet_state_dict(seed=1)
tab, keep = _lib.tensor_table(sd)
h = ctypes.c_void_p()
assert L.ytk_dbnet_create(tab, len(tab), 1280, 1600, ctypes.byref(h)) == 0, L.ytk_last_error()
rng = np.random.default_rng(0)
pages = torch.from_numpy(rng.integers(0, 256, size=(n, 1200, 1600, 3), dtype=np.uint8)).cuda()
out = torch.empty(n, 1184, 1600, device="cuda")
for _ in range(reps):
    assert L.ytk_dbnet_forward_u8(h, pages.data_ptr(), 1, n, 1200, 1600, out.data_ptr(), 1, None) == 0
    torch.cuda.synchronize()
print("launches", L.ytk_launch_count())
