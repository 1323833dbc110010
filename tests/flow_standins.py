"""Stand-in PARSeq for the product's TextRecognizer in the host-flow parity tests: the same function of (crop pixels,
padded width, mini-batch length) as oracle.refcheck.flow_model, which plays PARSeq inside the REFERENCE's
TextRecognizer.__call__ - so that equal outputs mean equal crops, order, mini-batches, padding, pairing and fallback
decisions (rows R4, R5, R10, R11)."""
import ctypes

import numpy as np

from oracle.refcheck import flow_hash, flow_token


def _rows(canvases, padded, groups, n_classes, S):
    n = len(canvases)
    ids = np.zeros((n, S), np.int32)
    probs = np.ones((n, S), np.float32)
    sizes = {}
    for g in groups:
        sizes[g] = sizes.get(g, 0) + 1
    for r, (c, wp, g) in enumerate(zip(canvases, padded, groups)):
        tok, p = flow_token(flow_hash(np.transpose(c, (2, 0, 1))), int(wp), sizes[g], n_classes)
        ids[r, 0], probs[r, 0] = tok, p
    return ids, probs


def install(rec):
    """Replaces both device entry points of rec.model (host-crop path: recognize_crops; device-crop path:
    run_packed_ptr over a host buffer) by the stand-in."""
    S = rec.model.max_label_length + 1
    C = rec.model.num_classes

    def recognize_crops(canvases, padded, groups, n_groups):
        ids, probs = _rows(canvases, padded, groups, C, S)
        return ids, probs, np.full((n_groups,), S, np.int32)

    def run_packed_ptr(ptr, on_device, total, descs, n, n_groups, stream=None):
        raw = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(total,))
        canv = [raw[int(d["pix_off"]):int(d["pix_off"]) + 32 * int(d["w"]) * 3].reshape(32, int(d["w"]), 3) for d in descs]
        ids, probs = _rows(canv, [int(d["wp"]) for d in descs], [int(d["group"]) for d in descs], C, S)
        return ids, probs, np.full((n_groups,), S, np.int32)

    rec.model.recognize_crops = recognize_crops
    rec.model.run_packed_ptr = run_packed_ptr
    return rec


# ---------------------------------------------------------------------------------------------------------- cases
BIG = [[[100, 100], [900, 100], [900, 240], [100, 240]], [[50, 300], [700, 300], [700, 370], [50, 370]],
       [[1000, 100], [1100, 100], [1100, 900], [1000, 900]]]
OUTSIDE = [[-5, 3], [40, 3], [40, 20], [-5, 20]]

# name -> (product model name, TextRecognizer flags, page seed, quads builder)
CASES = {
    "dynw_bucketing": ("parseq-tiny-dynw-v4", dict(dynamic_width=True, batch_bucketing=True), 5, lambda q: q[:70]),
    "fixed_width": ("parseq-tiny-dynw-v4", dict(), 6, lambda q: q[:33]),
    "no_width_budget": ("parseq-small", dict(dynamic_width=True, batch_bucketing=True), 7, lambda q: q[:150]),
    "dropped_quad": ("parseq-tiny-dynw-v4", dict(dynamic_width=True, batch_bucketing=True), 8,
                     lambda q: q[:30] + [OUTSIDE] + q[30:60]),
    "whole_page": ("parseq-tiny-dynw-v4", dict(dynamic_width=True, batch_bucketing=True), 9, lambda q: None),
    "source_downscale": ("parseq-tiny-dynw-v4", dict(dynamic_width=True, batch_bucketing=True, source_downscale=True), 10,
                         lambda q: q[:20] + BIG + q[20:30]),
    "orientation_fallback": ("parseq-tiny-dynw-v4", dict(dynamic_width=True, batch_bucketing=True,
                                                         rec_orientation_fallback=True), 11,
                             lambda q: q[:45] + [BIG[2]]),
    "fallback_and_downscale": ("parseq-tiny-dynw-v4", dict(dynamic_width=True, batch_bucketing=True, source_downscale=True,
                                                           rec_orientation_fallback=True,
                                                           rec_orientation_fallback_thresh=0.8), 12,
                               lambda q: BIG + q[:25]),
}


def case_inputs(name):
    from yomitoku_b200.synth import synthetic_page
    model, flags, seed, build = CASES[name]
    page, quads = synthetic_page(seed)
    if name == "source_downscale":
        page = np.ascontiguousarray(page[:1199, :1597])       # odd size: clipped last column / row of the pyramid
    return model, flags, page, build(quads)


def product_recognizer(name):
    from yomitoku_b200 import TextRecognizer
    model, flags, page, quads = case_inputs(name)
    rec = install(TextRecognizer(model_name=model, from_pretrained=False, device="cpu", **flags))
    return rec, page, quads


def reference_recognizer(name):
    """The reference's own TextRecognizer (oracle.refcheck.build_reference_recognizer_shell) configured like the
    product recognizer of the case."""
    from oracle import refcheck
    rec, page, quads = product_recognizer(name)
    d = rec._cfg.data
    ref = refcheck.build_reference_recognizer_shell(
        rec.charset, img_size=tuple(d.img_size), batch_size=d.batch_size, width_budget=getattr(d, "width_budget", None),
        max_batch_size=getattr(d, "max_batch_size", None), max_label_length=rec._cfg.max_label_length, **CASES[name][1])
    return ref, page, quads


# ------------------------------------------------------------------------------------------------ detector flow
def detector_pages():
    """Pages that need UP-scaling to the detector's input size, so that the product takes its host pre-processing path
    (the reference's way) instead of the fused GPU kernel: landscape and portrait."""
    import cv2
    from yomitoku_b200.synth import synthetic_page
    out = []
    for seed, size in ((4, (800, 600)), (13, (510, 690))):
        page, _ = synthetic_page(seed)
        out.append(cv2.resize(page, size, interpolation=cv2.INTER_AREA))
    return out


def product_detector():
    from oracle.refcheck import flow_detector_model
    from yomitoku_b200 import TextDetector
    det = TextDetector(from_pretrained=False, device="cpu")
    det.model = flow_detector_model
    return det
