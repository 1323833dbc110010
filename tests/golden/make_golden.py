"""Generates tests/golden/*.npz by running the REFERENCE's own code (loaded by path from /root/reference, see
oracle/refcheck.py) on seeded inputs.  Runs only in the build container; the fixtures travel to the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import parseq as ops  # noqa: E402
from oracle import refcheck, weights  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


# post-processor parameter sets of post_ref.npz (the v2_1 / v2 thresholds of the reference configs + a max_candidates cut)
POST_PARAMS = {
    "dbnetv2_1": dict(min_size=2, thresh=0.3, box_thresh=0.4, max_candidates=1500, unclip_ratio=3.5),
    "dbnetv2": dict(min_size=2, thresh=0.2, box_thresh=0.5, max_candidates=1500, unclip_ratio=5.0),
    "few": dict(min_size=2, thresh=0.3, box_thresh=0.4, max_candidates=20, unclip_ratio=3.5),
}


def crops_ref_page():
    """Page of crops_ref.npz: 3x3-px blocks of seeded noise (pure numpy, so the test rebuilds it instead of storing
    3.5 MB)."""
    g = np.random.default_rng(1717)
    return np.kron(g.integers(0, 256, size=(234, 567, 3), dtype=np.uint8), np.ones((3, 3, 1), np.uint8))[:700, :1700]


def main():
    assert refcheck.available(), "needs /root/reference"
    torch.manual_seed(0)
    # ---- DBNet: reference DBNet.forward on a seeded input
    sd = weights.make_dbnet_state_dict(seed=11)
    ref = refcheck.build_reference_dbnet(sd)
    x = torch.randn(1, 3, 64, 96, generator=torch.Generator().manual_seed(12))
    with torch.inference_mode():
        y = ref(x)["binary"]
    np.savez_compressed(os.path.join(OUT, "dbnet_ref.npz"), x=x.numpy(), prob=y.numpy(), weight_seed=11)
    # ---- PARSeq: reference PARSeq.forward (tiny-dynw, peaked + degenerate-repeat weights) + tokenizer decode
    charset = open(os.path.join(refcheck.SRC, "resource", "charsetv2.txt"), encoding="utf-8").read()
    spec = ops.SPECS["parseq-tiny-dynw-v4"]
    for tag, kw in (("peaked", dict(peaked=True)), ("repeat", dict(peaked=True, degenerate_repeat=True)),
                    ("random", dict())):
        sdp = weights.make_parseq_state_dict(spec, seed=21, **kw)
        m = refcheck.build_reference_parseq(spec, sdp, charset)
        img = torch.rand(3, 3, 32, 96, generator=torch.Generator().manual_seed(22)) * 2 - 1
        with torch.inference_mode():
            lg = m(img)
        strings, scores = m.tokenizer.decode(lg.softmax(-1))
        np.savez_compressed(os.path.join(OUT, "parseq_ref_%s.npz" % tag), img=img.numpy(),
                            ids=lg.argmax(-1).numpy().astype(np.int32),
                            maxlogit=lg.max(-1).values.numpy(), logits_pos0=lg[:, 0].numpy(),
                            strings=np.array(strings, dtype=object), scores=np.array(scores), weight_seed=21)
    # ---- host functions of reference data/functions.py (loaded by path with the absent imports stubbed)
    import types
    for name in ("pypdfium2",):
        sys.modules.setdefault(name, types.ModuleType(name))
    refcheck._pkg("ytk_ref")
    refcheck._pkg("ytk_ref.data")
    refcheck._pkg("ytk_ref.utils")
    refcheck._load("ytk_ref.constants", "constants.py", "ytk_ref")
    refcheck._load("ytk_ref.utils.logger", "utils/logger.py", "ytk_ref.utils")
    fn = refcheck._load("ytk_ref.data.functions", "data/functions.py", "ytk_ref.data")
    rng = np.random.default_rng(5)
    sizes = [(1200, 1600), (1600, 1200), (842, 596), (500, 3000), (40, 50), (2000, 2000)]
    res = []
    for h, w in sizes:
        img = np.zeros((h, w, 3), dtype=np.uint8)
        res.append(fn.resize_shortest_edge(img, 1280, 1600).shape[:2])
    page = rng.integers(0, 256, size=(300, 400, 3), dtype=np.uint8)
    quads = [[[10, 10], [200, 12], [198, 40], [9, 38]], [[50, 60], [70, 60], [70, 200], [50, 200]],
             [[0, 0], [400, 0], [400, 300], [0, 300]], [[390, 10], [420, 10], [420, 30], [390, 30]]]
    crops = {}
    for i, q in enumerate(quads):
        if fn.validate_quads(page, q) is None:
            crops["valid%d" % i] = np.array(0)
            continue
        crops["valid%d" % i] = np.array(1)
        roi = fn.rotate_text_image(fn.extract_roi_with_perspective(page, q), thresh_aspect=2)
        crops["fixed%d" % i] = fn.resize_with_padding(roi, [32, 800])
        crops["dyn%d" % i] = fn.resize_with_dynamic_padding(roi, [32, 800])
    # ---- more crops through the reference's own functions, for the device-side crop extraction (csrc/crop_math.h):
    # rotated rectangles, general quadrilaterals, tall (rotated) and very wide lines, integer shrink ratios
    rng2 = np.random.default_rng(17)
    page2 = crops_ref_page()
    quads2 = []
    for k in range(48):
        kind = k % 6
        if kind == 0:
            w, h = int(rng2.integers(4, 380)), int(rng2.integers(4, 120))
            x, y = int(rng2.integers(0, 1700 - w)), int(rng2.integers(0, 700 - h))
            q = [[x, y], [x + w, y], [x + w, y + h], [x, y + h]]
        elif kind in (1, 2):
            cx, cy = rng2.uniform(160, 1540), rng2.uniform(160, 540)
            w, h = rng2.uniform(6, 280), rng2.uniform(6, 90)
            a = rng2.uniform(-0.5, 0.5) if kind == 1 else rng2.uniform(-3.1, 3.1)
            c, sn = np.cos(a), np.sin(a)
            pts = np.array([[-w / 2, -h / 2], [w / 2, -h / 2], [w / 2, h / 2], [-w / 2, h / 2]])
            q = (pts @ np.array([[c, sn], [-sn, c]]) + [cx, cy] + rng2.uniform(-3, 3, (4, 2)) * (kind == 2)).tolist()
        elif kind == 3:
            w = int(rng2.integers(3, 50))
            h = int(rng2.integers(2 * w + 1, 690))
            x, y = int(rng2.integers(0, 1700 - w)), int(rng2.integers(0, 700 - h))
            q = [[x, y], [x + w, y], [x + w, y + h], [x, y + h]]
        elif kind == 4:
            w, h = int(rng2.integers(801, 1700)), int(rng2.integers(8, 80))
            x, y = int(rng2.integers(0, 1700 - w + 1)), int(rng2.integers(0, 700 - h))
            q = [[x, y], [x + w, y], [x + w, y + h], [x, y + h]]
        else:
            r = int(rng2.integers(2, 5))
            h, w = 32 * r, r * int(rng2.integers(2, 150))
            x, y = int(rng2.integers(0, 1700 - w)), int(rng2.integers(0, 700 - h))
            q = [[x, y], [x + w, y], [x + w, y + h], [x, y + h]]
        quads2.append(q)
    crops2 = {}
    for i, q in enumerate(quads2):
        assert fn.validate_quads(page2, q) is not None
        roi = fn.rotate_text_image(fn.extract_roi_with_perspective(page2, q), thresh_aspect=2)
        dyn = fn.resize_with_dynamic_padding(roi, [32, 800])
        crops2["dyn%d" % i] = dyn
        if i % 8 == 0:
            crops2["fixed%d" % i] = fn.resize_with_padding(roi, [32, 800])
    np.savez_compressed(os.path.join(OUT, "crops_ref.npz"), quads=np.array(quads2, dtype=np.float64), **crops2)
    # ---- DBNet post-processing (row R3): the reference's own DBnetPostProcessor (executed from /root/reference with the
    # oracle's stand-ins for pyclipper / shapely, see oracle/refcheck.py) on seeded probability maps
    post = {}
    cases = refcheck.postprocess_cases()
    for ci, (pu8, ori) in enumerate(cases):
        post["prob%d" % ci] = pu8
        post["ori%d" % ci] = np.array(ori)
    for name, kw in POST_PARAMS.items():
        ref_post = refcheck.build_reference_postprocessor(**kw)
        for ci, (pu8, ori) in enumerate(cases):
            q, sc = refcheck.reference_postprocess(ref_post, pu8.astype(np.float32) / 255.0, ori)
            post["%s_quads%d" % (name, ci)] = np.array(q, dtype=np.int16).reshape(-1, 4, 2)
            post["%s_scores%d" % (name, ci)] = np.array(sc, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "post_ref.npz"), **post)
    # ---- the recognizer's host flow: the reference's own TextRecognizer.__call__ with a stand-in PARSeq
    # (oracle/refcheck.py: build_reference_recognizer_shell, tests/flow_standins.py)
    sys.path.insert(0, os.path.dirname(OUT))
    import flow_standins as FS
    flow = {}
    for name in FS.CASES:
        ref_rec, fpage, fquads = FS.reference_recognizer(name)
        r, _ = ref_rec(fpage, fquads)
        flow[name + "_contents"] = np.array(r["contents"], dtype=object)
        flow[name + "_scores"] = np.array(r["scores"], dtype=np.float64)
        flow[name + "_directions"] = np.array(r["directions"], dtype=object)
        print("flow case %-24s %4d crops" % (name, len(r["contents"])))
    np.savez_compressed(os.path.join(OUT, "flow_ref.npz"), **flow)
    # ---- the detector's host flow: the reference's own TextDetector.__call__ (pre-processing, post-processing) with a
    # stand-in DBNet, on pages that need up-scaling
    ref_det = refcheck.build_reference_detector_shell()
    detflow = {}
    for i, dpage in enumerate(FS.detector_pages()):
        r, _ = ref_det(dpage)
        detflow["points%d" % i] = np.array(r["points"], dtype=np.int16).reshape(-1, 4, 2)
        detflow["scores%d" % i] = np.array(r["scores"], dtype=np.float64)
        print("detector flow page %d: %d boxes" % (i, len(r["points"])))
    np.savez_compressed(os.path.join(OUT, "detflow_ref.npz"), **detflow)
    std = fn.standardization_image(page.astype(np.float32))
    np.savez_compressed(os.path.join(OUT, "host_ref.npz"), sizes=np.array(sizes), resized=np.array(res), page=page,
                        quads=np.array(quads), std=std, **crops)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
