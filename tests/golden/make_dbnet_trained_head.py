"""Trains the `decoder.binarize` head of the seeded random DBNet (backbone / FPN frozen at their seeded values) on
synthetic pages so that the detector's OWN probability map contains the page's text lines, and stores the head in
tests/golden/dbnet_head_trained.npz.  With it tests/test_gpu_dbnet.py thresholds the device's map itself (instead of a
synthetic map carrying the device's error field) and compares the polygons with those of the fp32 oracle's map.

Test infrastructure: runs the CPU oracle (oracle/dbnet.py) under autograd; ~5 minutes on 8 cores.
    python tests/golden/make_dbnet_trained_head.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import dbnet as odb  # noqa: E402
from oracle import pipeline as opipe  # noqa: E402
from yomitoku_b200.models import _dbnet_random_state_dict  # noqa: E402
from yomitoku_b200.synth import synthetic_page  # noqa: E402

HEAD = ["decoder.binarize.0.weight", "decoder.binarize.1.weight", "decoder.binarize.1.bias",
        "decoder.binarize.3.weight", "decoder.binarize.3.bias", "decoder.binarize.4.weight",
        "decoder.binarize.4.bias", "decoder.binarize.6.weight", "decoder.binarize.6.bias"]
OUT = os.path.join(ROOT, "tests", "golden", "dbnet_head_trained.npz")


def target_mask(quads, hw, page_hw, shrink=3):
    """DB-style training target: the text-line rectangles shrunk by `shrink` px, at network resolution."""
    H, W = hw
    sy, sx = H / page_hw[0], W / page_hw[1]
    m = np.zeros((H, W), dtype=np.float32)
    for q in quads:
        (x0, y0), (x1, y1) = q[0], q[2]
        m[int(round((y0 + shrink) * sy)):int(round((y1 - shrink) * sy)),
          int(round((x0 + shrink) * sx)):int(round((x1 - shrink) * sx))] = 1.0
    return m


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    sd = _dbnet_random_state_dict(0)
    fuses, masks = [], []
    t0 = time.time()
    for i in range(100, 104):
        page, quads = synthetic_page(i)
        x = opipe.detector_preprocess(page)
        with torch.inference_mode():
            fuses.append(odb.decoder_fuse(sd, odb.backbone_features(sd, x)).clone())
        masks.append(torch.from_numpy(target_mask(quads, x.shape[-2:], page.shape[:2]))[None, None])
    print("features of %d pages in %.1f s; fuse std %.2f" % (len(fuses), time.time() - t0, fuses[0].std().item()))
    params = {k: sd[k].clone() for k in HEAD}
    params["decoder.binarize.0.weight"] /= fuses[0].std()        # unit-variance pre-activations at the start
    for v in params.values():
        v.requires_grad_(True)
    opt = torch.optim.Adam(list(params.values()), lr=2e-3)
    g = torch.Generator().manual_seed(1)
    C = 88          # crop side at 1/4 resolution
    steps = int(os.environ.get("STEPS", "400"))
    for step in range(steps):
        fb, mb = [], []
        for _ in range(6):
            k = int(torch.randint(0, len(fuses), (1,), generator=g))
            y = int(torch.randint(0, fuses[k].shape[2] - C + 1, (1,), generator=g))
            x = int(torch.randint(0, fuses[k].shape[3] - C + 1, (1,), generator=g))
            fb.append(fuses[k][:, :, y:y + C, x:x + C])
            mb.append(masks[k][:, :, 4 * y:4 * (y + C), 4 * x:4 * (x + C)])
        fb, mb = torch.cat(fb), torch.cat(mb)
        local = dict(sd)
        local.update(params)
        logits = odb.binarize_logits(local, fb)
        loss = F.binary_cross_entropy_with_logits(logits, mb)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if step % 25 == 0 or step + 1 == steps:
            with torch.no_grad():
                acc = ((logits > 0) == (mb > 0.5)).float().mean().item()
            print("step %4d loss %.4f pixel acc %.4f (%.0f s)" % (step, loss.item(), acc, time.time() - t0), flush=True)
        if step == int(steps * 0.7):
            for gp in opt.param_groups:
                gp["lr"] = 5e-4
    np.savez_compressed(OUT, **{k: v.detach().numpy().astype(np.float32) for k, v in params.items()})
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    return 0


if __name__ == "__main__":
    main()
