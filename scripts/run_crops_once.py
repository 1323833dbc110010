"""Runs the device-side crop extraction a few times on the bench's workload (16 synthetic 1200x1600 pages, 200 text
lines each -> 3200 crops per call), for ncu launch lists / captures, and prints CUDA-event times + algorithmic bytes.

    python scripts/run_crops_once.py [pages] [reps]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yomitoku_b200 import _lib  # noqa: E402
from yomitoku_b200.data import crop_geometry  # noqa: E402
from yomitoku_b200.models import extract_crops_device  # noqa: E402
from yomitoku_b200.synth import synthetic_page  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
pages, geoms = [], []
for i in range(P):
    pg, q = synthetic_page(i)
    pages.append(pg)
    geoms.append(crop_geometry(pg.shape, q, [32, 800], True, page=i)[0])
geoms = np.concatenate(geoms)
dev = torch.from_numpy(np.stack(pages)).cuda()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
times = []
for r in range(reps):
    ev[0].record()
    canv, total = extract_crops_device(dev, geoms)
    ev[1].record()
    torch.cuda.synchronize()
    times.append(ev[0].elapsed_time(ev[1]))
# algorithmic bytes: every ROI pixel reads <= 4 page pixels (counted once: the bounding box) and is written once;
# every ROI pixel is read once by the resize; every canvas byte is written once
roi = int((geoms["w"].astype(np.int64) * geoms["h"] * 3).sum())
box = int((geoms["rw"].astype(np.int64) * geoms["rh"] * 3).sum())
print(json.dumps({"crops": int(len(geoms)), "pages": P, "ms_per_call": times, "roi_bytes": roi, "box_bytes": box,
                  "canvas_bytes": int(total), "algorithmic_bytes": box + 2 * roi + int(total),
                  "launches": int(_lib.lib().ytk_launch_count())}))
