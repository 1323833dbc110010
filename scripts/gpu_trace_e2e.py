"""Where does the wall time of BatchedOCR.stream go?  python scripts/gpu_trace_e2e.py [steps] [pages] [lookahead]"""
import collections
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    look = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    from yomitoku_b200 import TextDetector, TextRecognizer
    from yomitoku_b200 import pipeline as pl
    from yomitoku_b200.synth import synthetic_page, synthetic_prob_map
    det = TextDetector(from_pretrained=False, device="cuda")
    rec = TextRecognizer(model_name="parseq-large-v4_1", from_pretrained=False, device="cuda", dynamic_width=True,
                         batch_bucketing=True)
    pages, quads = zip(*[synthetic_page(i) for i in range(P)])
    pages = list(pages)
    Hn, Wn = det.model.input_size(1200, 1600)
    probs = [synthetic_prob_map(q, (Hn, Wn), (1200, 1600)) for q in quads]
    ncpu = os.cpu_count() or 2
    ocr = pl.BatchedOCR(det, rec, det_batch=8, workers=max(2, min(32, ncpu - 2)))
    for _ in ocr.stream([pages] * 3, lookahead=look, prob_override=[probs] * 3):
        pass
    torch.cuda.synchronize()
    pl.TRACE = []
    t0 = time.perf_counter()
    for res in ocr.stream([pages] * steps, lookahead=look, prob_override=[probs] * steps):
        pass
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("cpus %d workers %d: %.1f pages/s, %.1f ms/step" % (ncpu, ocr.workers, steps * P / dt, dt / steps * 1e3))
    agg = collections.defaultdict(list)
    for name, th, a, b in pl.TRACE:
        agg[(th, name)].append((b - a) * 1e3)
    for k in sorted(agg):
        v = agg[k]
        print("%-14s %-22s n=%3d mean %.1f ms  max %.1f" % (k[0][:14], k[1], len(v), sum(v) / len(v), max(v)))
    ws = [(a - t0, b - t0) for name, th, a, b in pl.TRACE if name == "worker.post"]
    print("first 16 worker jobs: start %s" % " ".join("%.0f" % (a * 1e3) for a, b in ws[:16]))
    for name, th, a, b in pl.TRACE:
        if name in ("submit.detect", "collect.wait_host", "collect.recognize", "recognize.device"):
            print("%8.1f %8.1f %s" % ((a - t0) * 1e3, (b - t0) * 1e3, name))
    ocr.close()


if __name__ == "__main__":
    main()
