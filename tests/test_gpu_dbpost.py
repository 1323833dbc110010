"""GPU: the device-side front half of the DBNet post-processing (csrc/dbpost_ops.cu behind ytk_dbnet_post_front) against
its scipy restatement (oracle/dbpost.py), and end to end: TextDetector / BatchedOCR with `device_post` return exactly
the quads of the host path (OpenCV contours on the downloaded map, reference dbnet_postporcessor.py:39-82), including
pages that fall back because a component has a hole."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle.dbpost import post_front
from yomitoku_b200 import TextDetector
from yomitoku_b200.models import DB_RUN_DTYPE, dbnet_post_front
from yomitoku_b200.postprocessor import DBnetPostProcessor
from yomitoku_b200.synth import synthetic_page, synthetic_prob_map

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_host_logic import _blob_map  # noqa: E402

pytestmark = pytest.mark.gpu


def _front(maps, thresh, **kw):
    dev = torch.from_numpy(np.stack(maps)).cuda()
    return dbnet_post_front(dev, thresh, **kw)


def test_runs_equal_scipy_twin():
    maps = [_blob_map(s, holes=(s % 3 == 0)) for s in range(12)]
    runs, meta = _front(maps, 0.3)
    n_fallback = 0
    for i, prob in enumerate(maps):
        ref, comps, holes = post_front(prob, 0.3)
        assert meta[i][0] == len(ref) and meta[i][1] == comps and meta[i][1] - meta[i][2] // 4 == holes
        assert meta[i][2] % 4 == 0
        if holes:
            assert runs[i] is None
            n_fallback += 1
            continue
        got = np.sort(runs[i], order=("root", "y", "x0"))
        assert got.dtype == DB_RUN_DTYPE
        for f in ("root", "y", "x0", "x1"):
            assert np.array_equal(got[f], ref[f]), f
        assert np.array_equal(got["sum"], ref["sum"])        # fp64 accumulation left to right in both: bit-equal
    assert 0 < n_fallback < len(maps)


def test_full_size_maps_and_overflow():
    maps = []
    for s in range(3):
        _, quads = synthetic_page(40 + s)
        maps.append(synthetic_prob_map(quads, (1184, 1600), (1200, 1600)))
    runs, meta = _front(maps, 0.3)
    pp = DBnetPostProcessor(min_size=2, thresh=0.3, box_thresh=0.4, max_candidates=1500, unclip_ratio=3.5)
    for i, prob in enumerate(maps):
        ref, comps, holes = post_front(prob, 0.3)
        assert holes == 0 and runs[i] is not None and len(runs[i]) == len(ref) and meta[i][1] == comps
        b_host = pp.boxes_from_bitmap(prob, prob > 0.3, 1600, 1200)
        b_dev = pp.boxes_from_runs(runs[i], 1600, 1184, 1600, 1200)
        assert b_host[0] == b_dev[0] and len(b_dev[0]) == 200
        assert np.allclose(b_host[1], b_dev[1], rtol=1e-12, atol=0)
    # a run buffer that is too small is reported, not truncated silently
    runs2, meta2 = _front(maps[:1], 0.3, max_runs=100)
    assert runs2[0] is None and meta2[0][0] > 100


def test_degenerate_maps():
    H, W = 64, 96
    empty = np.zeros((H, W), np.float32)
    full = np.ones((H, W), np.float32)
    ring = np.zeros((H, W), np.float32)
    ring[10:30, 10:40] = 0.9
    ring[15:25, 15:35] = 0.0                # one component with one hole
    diag = np.zeros((H, W), np.float32)
    for k in range(20):
        diag[5 + k, 5 + k] = 0.8            # 8-connected diagonal: one component
        diag[40 - k, 50 + k] = 0.8          # anti-diagonal: one component
    runs, meta = _front([empty, full, ring, diag], 0.3)
    assert len(runs[0]) == 0 and meta[0][1] == 0
    assert len(runs[1]) == H and meta[1][1] == 1 and set(runs[1]["root"]) == {0}
    assert np.array_equal(np.sort(runs[1]["sum"]), np.full(H, float(W)))
    assert runs[2] is None and meta[2][1] == 1 and meta[2][2] == 0
    assert len(runs[3]) == 40 and meta[3][1] == 2 and meta[3][2] == 8
    assert set(runs[3]["root"]) == {5 * W + 5, 21 * W + 69}


def _detector():
    return TextDetector(from_pretrained=False, device="cuda")


def test_text_detector_device_post_equals_host_post():
    """Random-weight DBNet maps are noise (components with holes: host fallback), the blob maps below are hole-free or
    not by construction: both branches of the per-page decision must return the host path's quads.  (The comparison is
    on ONE forward pass.)"""
    det = _detector()
    assert det.device_post
    pages = np.stack([synthetic_page(60 + i)[0] for i in range(2)])
    prob = det.model.detect_pages_u8(torch.from_numpy(pages).cuda())
    got = det.postprocess_device(prob, pages.shape[1:3])
    prob_h = prob.cpu().numpy()
    for i, (quads, scores) in enumerate(got):
        q_ref, s_ref = det.postprocess({"binary": prob_h[i:i + 1, None]}, pages.shape[1:3])
        assert quads == q_ref and np.allclose(scores, s_ref, rtol=1e-12, atol=0)
    # the public calls run (device path) and return the schema
    one, _ = det(pages[0])
    many = det.detect_pages(list(pages))
    assert len(many) == 2 and len(one.points) == len(one.scores)
    det.device_post = False
    assert len(det.detect_pages(list(pages))) == 2

    # the same through postprocess_device on maps with known content (hole-free and with holes)
    maps = [_blob_map(3, False), _blob_map(9, False), _blob_map(201, True)]
    dev = torch.from_numpy(np.stack(maps)).cuda()
    res = det.postprocess_device(dev, (600, 840))
    for prob, (quads, scores) in zip(maps, res):
        q_ref, s_ref = det.postprocess({"binary": prob[None, None]}, (600, 840))
        assert quads == q_ref and np.allclose(scores, s_ref, rtol=1e-12, atol=0) and len(quads) > 0


def test_batched_ocr_device_post_equals_host_post():
    from test_gpu_api import _ocr
    from yomitoku_b200.pipeline import BatchedOCR
    o = _ocr()
    pages, probs = [], []
    for i in range(3):
        p, q = synthetic_page(70 + i)
        pages.append(p)
        probs.append(synthetic_prob_map(q, (1184, 1600), (1200, 1600)))
    probs[1] = probs[1].copy()
    ys, xs = np.nonzero(probs[1] > 0.9)
    probs[1][ys[len(ys) // 2], xs[len(ys) // 2]] = 0.0        # a hole: this page takes the host path
    out = {}
    for mode in (True, False):
        o.detector.device_post = mode
        b = BatchedOCR(o.detector, o.recognizer, workers=2, det_batch=2)
        try:
            out[mode] = b(pages, prob_override=probs)
            if mode:
                assert (b.post_front_pages, b.post_host_pages) == (2, 1)
            else:
                assert b.post_front_pages == 0
        finally:
            b.close()
    for a, c in zip(out[True], out[False]):
        assert [w.points for w in a.words] == [w.points for w in c.words] and len(a.words) >= 199
        assert [w.content for w in a.words] == [w.content for w in c.words]
        assert np.allclose([w.det_score for w in a.words], [w.det_score for w in c.words], rtol=1e-12, atol=0)
