"""One launch sequence of the device-side DBNet post-processing front half on 8 bench-like maps (for ncu launch lists
and CUDA-event timing): python scripts/run_dbpost_once.py [n_pages] [iters]"""
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yomitoku_b200 import _lib  # noqa: E402
from yomitoku_b200.synth import synthetic_page, synthetic_prob_map  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
maps = [synthetic_prob_map(synthetic_page(i)[1], (1184, 1600), (1200, 1600)) for i in range(n)]
dev = torch.from_numpy(np.stack(maps)).cuda()
L = _lib.lib()
labels = torch.empty((n, 1184, 1600), dtype=torch.int32, device="cuda")
runs = torch.empty((n, 32768, 24), dtype=torch.uint8, device="cuda")
meta = torch.empty((n, 4), dtype=torch.int32, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(iters):
    ev[0].record()
    _lib.check(L.ytk_dbnet_post_front(dev.data_ptr(), n, 1184, 1600, 0.3, labels.data_ptr(), labels.numel() * 4,
                                      runs.data_ptr(), 32768, meta.data_ptr(), None))
    ev[1].record()
    torch.cuda.synchronize()
    print("post_front %d pages: %.3f ms" % (n, ev[0].elapsed_time(ev[1])), meta.cpu().numpy()[0].tolist())
