"""Diagnostic: host-side time of every part of the device-crops recognizer step in the bench's `value` loop
(detector on the default stream, then crop extraction, then the PARSeq forward), against the host-crops step."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yomitoku_b200 import TextDetector, TextRecognizer, _lib  # noqa: E402
from yomitoku_b200.data import ParseqDataset, crop_geometry  # noqa: E402
from yomitoku_b200.models import extract_crops_device  # noqa: E402
from yomitoku_b200.synth import synthetic_page  # noqa: E402
from yomitoku_b200.text_recognizer import plan_mini_batches  # noqa: E402

P = 16
det = TextDetector(from_pretrained=False, device="cuda")
rec = TextRecognizer(model_name=sys.argv[1] if len(sys.argv) > 1 else "parseq-large-v4_1", from_pretrained=False,
                     device="cuda", dynamic_width=True, batch_bucketing=True)
L = _lib.lib()
pages, quads = zip(*[synthetic_page(i) for i in range(P)])
pages_dev = torch.from_numpy(np.stack(pages)).cuda()
prob_dev = torch.empty((P, 1184, 1600), dtype=torch.float32, device="cuda")
flat_c, flat_p, flat_g, flat_geoms, g0 = [], [], [], [], 0
for pi, (pg, q) in enumerate(zip(pages, quads)):
    ds = ParseqDataset(rec._cfg, pg, q, dynamic_width=True)
    g, _ = crop_geometry(pg.shape, q, rec._cfg.data.img_size, True, page=pi)
    order = np.argsort(ds.content_widths).tolist()
    plan = plan_mini_batches([c.shape[1] for c in ds.data], order, True, rec._cfg.data.batch_size, None, None)
    padded, group = rec._collate_widths(ds.data, plan)
    for b in plan:
        for i in b:
            flat_c.append(ds.data[i]); flat_p.append(padded[i]); flat_g.append(g0 + group[i])
    flat_geoms.append(g[np.asarray([i for b in plan for i in b], np.int64)])
    g0 += len(plan)
sel = np.concatenate(flat_geoms)
buf, total, descs, _ = rec.model.pack_crops(flat_c, flat_p, flat_g)
buf_dev = buf.cuda()
n = len(flat_c)


def det_step():
    for s in range(0, P, 8):
        _lib.check(L.ytk_dbnet_forward_u8(det.model._ensure(), pages_dev[s:s + 8].data_ptr(), 1, 8, 1200, 1600,
                                          prob_dev[s:s + 8].data_ptr(), 1, None))


def step(mode, pre_sync):
    t = [time.perf_counter()]
    det_step()
    t.append(time.perf_counter())
    if pre_sync:
        torch.cuda.synchronize()
    t.append(time.perf_counter())
    if mode == "host":
        canv = buf_dev
    else:
        canv, tot = extract_crops_device(pages_dev, sel)
        assert tot == total
    t.append(time.perf_counter())
    rec.model.run_packed(canv, total, descs, n, g0)
    t.append(time.perf_counter())
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    return [round((b - a) * 1e3, 2) for a, b in zip(t[:-1], t[1:])]


for mode in ("host", "dev", "host", "dev"):
    for pre_sync in (False, True):
        for it in range(4):
            r = step(mode, pre_sync)
            print(mode, "pre_sync" if pre_sync else "no_sync", it,
                  dict(zip(("det_launch", "sync", "extract_call", "forward", "final_sync"), r)),
                  rec.model.last_phase_ms(), flush=True)
