"""Oracle: host-side pre/post-processing and the end-to-end CPU pipeline (TEST INFRASTRUCTURE, see __init__.py).

Restates, with the same OpenCV 4.13 the reference pins, the non-model parts of the hot path:
  R1 detector_preprocess   text_detector.py:99-107, data/functions.py:196-264
  R3 dbnet_postprocess     postprocessor/dbnet_postporcessor.py:16-138 (+ Clipper 6.4.2 / shapely restatements, parity
                           unpinned against the real pyclipper / shapely which are not installable offline)
  R4 make_crops            data/dataset.py:44-129, data/functions.py:267-439
  R5 mini_batches          text_recognizer.py:135-203
  R10/R11 recognize / ocr  text_recognizer.py:232-399, ocr.py:6-63
and chains them with oracle.dbnet / oracle.parseq into `ocr_page`, the CPU baseline of bench.py.
"""
import math
import unicodedata

import cv2
import numpy as np
import torch
import torch.nn.functional as F

from . import dbnet as odb
from . import parseq as ops


# ------------------------------------------------------------------------------------------------ R1
def detector_input_size(h, w, shortest=1280, limit=1600):
    s = shortest / min(h, w)
    nh, nw = (shortest, int(w * s)) if h < w else (int(h * s), shortest)
    if max(nh, nw) > limit:
        s2 = float(limit) / max(nh, nw)
        nh, nw = int(nh * s2), int(nw * s2)
    return max(int(nh / 32) * 32, 32), max(int(nw / 32) * 32, 32)


def detector_preprocess(img_bgr, shortest=1280, limit=1600):
    """-> (1,3,H',W') fp32; note the double channel flip: the net sees B,G,R with RGB statistics (Appendix A2)."""
    x = img_bgr.copy()[:, :, ::-1].astype(np.float32)
    nh, nw = detector_input_size(x.shape[0], x.shape[1], shortest, limit)
    x = cv2.resize(x, (nw, nh), interpolation=cv2.INTER_AREA)
    x = x[:, :, ::-1] / 255.0
    x = ((x - np.array((0.485, 0.456, 0.406))) / np.array((0.229, 0.224, 0.225))).astype(np.float32)
    return torch.as_tensor(np.transpose(x, (2, 0, 1)), dtype=torch.float)[None]


# ------------------------------------------------------------------------------------------------ R3
def _round_half_away(v):
    return np.where(v < 0, np.trunc(v - 0.5), np.trunc(v + 0.5)).astype(np.int64)


def clipper_offset_box(box, delta):
    """ClipperOffset.AddPath(box, JT_ROUND, ET_CLOSEDPOLYGON); Execute(delta) for a convex box (Clipper 6.4.2
    DoOffset/DoRound with ArcTolerance 0.25): vectorised over the arc points.  Returns int64 (n,2) vertices."""
    pts = np.trunc(np.asarray(box, dtype=np.float64)).astype(np.int64)
    keep = [0] + [i for i in range(1, len(pts)) if (pts[i] != pts[i - 1]).any()]
    pts = pts[keep]
    while len(pts) > 1 and (pts[0] == pts[-1]).all():
        pts = pts[:-1]
    n = len(pts)
    if n < 3:
        return np.zeros((0, 2), dtype=np.int64)
    x, y = pts[:, 0].astype(np.float64), pts[:, 1].astype(np.float64)
    area = -0.5 * np.sum((np.roll(x, 1) + x) * (np.roll(y, 1) - y))
    if area < 0:
        pts = pts[::-1].copy()
    d = (np.roll(pts, -1, axis=0) - pts).astype(np.float64)
    ln = np.hypot(d[:, 0], d[:, 1])
    ln[ln == 0] = np.inf
    normals = np.stack([d[:, 1] / ln, -d[:, 0] / ln], axis=1)        # GetUnitNormal = (dy, -dx) / len
    ytol = 0.25 if 0.25 <= abs(delta) * 0.25 else abs(delta) * 0.25
    steps = math.pi / math.acos(1 - ytol / abs(delta))
    steps = min(steps, abs(delta) * math.pi)
    sn, cs = math.sin(2 * math.pi / steps), math.cos(2 * math.pi / steps)
    per_rad = steps / (2 * math.pi)
    out = []
    k = n - 1
    for j in range(n):
        nk, nj = normals[k], normals[j]
        sin_a = nk[0] * nj[1] - nj[0] * nk[1]
        cos_a = nk[0] * nj[0] + nj[1] * nk[1]
        p = pts[j].astype(np.float64)
        if abs(sin_a * delta) < 1.0 and cos_a > 0:
            out.append(_round_half_away(p + nk * delta))   # OffsetPoint returns here, BEFORE `k = j` (Clipper 6.4.2)
            continue
        sin_a = min(1.0, max(-1.0, sin_a))
        if sin_a * delta < 0:
            out += [_round_half_away(p + nk * delta), pts[j], _round_half_away(p + nj * delta)]
            k = j
            continue
        m = max(int(_round_half_away(np.float64(per_rad * abs(math.atan2(sin_a, cos_a))))), 1)
        vx, vy = nk
        for _ in range(m):                                               # repeated rotation, like DoRound
            out.append(_round_half_away(p + np.array([vx, vy]) * delta))
            vx, vy = vx * cs - sn * vy, vx * sn + vy * cs
        out.append(_round_half_away(p + nj * delta))
        k = j
    return np.array(out, dtype=np.int64)


def _mini_box(contour):
    rect = cv2.minAreaRect(contour)
    p = sorted(list(cv2.boxPoints(rect)), key=lambda t: t[0])
    i1, i4 = (0, 1) if p[1][1] > p[0][1] else (1, 0)
    i2, i3 = (2, 3) if p[3][1] > p[2][1] else (3, 2)
    return [p[i1], p[i2], p[i3], p[i4]], min(rect[1])


def _box_score(prob, contour):
    h, w = prob.shape
    b = contour.copy()
    x0 = np.clip(np.floor(b[:, 0].min()).astype(int), 0, w - 1)
    x1 = np.clip(np.ceil(b[:, 0].max()).astype(int), 0, w - 1)
    y0 = np.clip(np.floor(b[:, 1].min()).astype(int), 0, h - 1)
    y1 = np.clip(np.ceil(b[:, 1].max()).astype(int), 0, h - 1)
    mask = np.zeros((y1 - y0 + 1, x1 - x0 + 1), dtype=np.uint8)
    b[:, 0] -= x0
    b[:, 1] -= y0
    cv2.fillPoly(mask, b.reshape(1, -1, 2).astype(np.int32), 1)
    return cv2.mean(prob[y0:y1 + 1, x0:x1 + 1], mask)[0]


def dbnet_postprocess(prob, ori_hw, thresh=0.3, box_thresh=0.4, max_candidates=1500, unclip_ratio=3.5, min_size=2):
    """prob: (H,W) float32 numpy.  Returns (quads, scores) like DBnetPostProcessor.__call__ (v2_1 defaults)."""
    H, W = prob.shape
    dest_h, dest_w = ori_hw
    contours, _ = cv2.findContours(((prob > thresh) * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    quads, scores = [], []
    for cnt in contours[:max_candidates]:
        cnt = cnt.squeeze(1)
        pts, sside = _mini_box(cnt)
        if sside < min_size:
            continue
        pts = np.array(pts)
        score = _box_score(prob, cnt)
        if box_thresh > score:
            continue
        bw, bh = pts[:, 0].max() - pts[:, 0].min(), pts[:, 1].max() - pts[:, 1].min()
        ratio = unclip_ratio / math.sqrt(min(bw, bh))
        p64 = pts.astype(np.float64)
        area = 0.5 * abs(np.dot(p64[:, 0], np.roll(p64[:, 1], -1)) - np.dot(p64[:, 1], np.roll(p64[:, 0], -1)))
        peri = np.sqrt(((p64 - np.roll(p64, -1, axis=0)) ** 2).sum(1)).sum()
        grown = clipper_offset_box(pts, area * ratio / peri).reshape(-1, 1, 2)
        box, sside = _mini_box(grown)
        if sside < min_size + 2:
            continue
        box = np.array(box)
        box[:, 0] = np.clip(np.round(box[:, 0] / W * dest_w), 0, dest_w)
        box[:, 1] = np.clip(np.round(box[:, 1] / H * dest_h), 0, dest_h)
        quads.append(box.astype(np.int16).tolist())
        scores.append(score)
    return quads, scores


# ------------------------------------------------------------------------------------------------ R4
def _fit(h, w, target):
    s = min(target[1] / w if w > target[1] else 1.0, target[0] / h if h > target[0] else 1.0)
    return max(1, int(h * s)), max(1, int(w * s))


def make_crop(img_rgb, quad, img_size=(32, 800), dynamic_width=False):
    """One quad -> (canvas u8 (32,Wc,3), content_width) or None if the quad is rejected (validate_quads)."""
    if len(quad) != 4 or any(len(p) != 2 for p in quad):
        return None
    qi = np.array(quad, dtype=int)
    H, W = img_rgb.shape[:2]
    if qi[:, 0].min() < 0 or qi[:, 0].max() > W or qi[:, 1].min() < 0 or qi[:, 1].max() > H:
        return None
    q = np.array(quad, dtype=np.int64)
    x0, y0 = int(q[:, 0].min()), int(q[:, 1].min())
    roi = img_rgb[y0:int(q[:, 1].max()), x0:int(q[:, 0].max()), :]
    q = q - np.array([x0, y0])
    w = int(np.linalg.norm(q[0] - q[1]))
    h = int(np.linalg.norm(q[1] - q[2]))
    M = cv2.getPerspectiveTransform(np.float32(q), np.float32([[0, 0], [w, 0], [w, h], [0, h]]))
    crop = cv2.warpPerspective(roi, M, (w, h))
    if crop.shape[0] > 2 * crop.shape[1]:
        crop = cv2.rotate(crop, cv2.ROTATE_90_COUNTERCLOCKWISE)
    nh, nw = _fit(crop.shape[0], crop.shape[1], img_size)
    small = cv2.resize(crop, (nw, nh), interpolation=cv2.INTER_AREA)
    cw = min(img_size[1], ((nw + 64 + 7) // 8) * 8) if dynamic_width else img_size[1]
    canvas = np.zeros((img_size[0], cw, 3), dtype=np.uint8)
    canvas[:small.shape[0], :small.shape[1]] = small
    return canvas, nw


def to_tensor(canvas):
    t = torch.from_numpy(canvas.transpose(2, 0, 1).copy()).to(torch.float32).div(255)
    return (t - 0.5) / 0.5


# ------------------------------------------------------------------------------------------------ R5
def mini_batches(widths, order, dynamic_width, batch_size, width_budget=None, max_batch_size=None):
    """Index lists of the reference's mini-batches (text_recognizer.py:158-203)."""
    idxs = list(order) if order is not None else list(range(len(widths)))
    if dynamic_width and width_budget:
        res, cur, cur_max = [], [], 0
        for i in idxs:
            new_max = max(cur_max, widths[i])
            if cur and ((len(cur) + 1) * new_max > width_budget or
                        (max_batch_size is not None and len(cur) >= max_batch_size)):
                res.append(cur)
                cur, new_max = [], widths[i]
            cur.append(i)
            cur_max = new_max
        return res + ([cur] if cur else [])
    return [idxs[s:s + batch_size] for s in range(0, len(idxs), batch_size)]


# ------------------------------------------------------------------------------------------------ R4-R11
def recognize(sd, spec, tokenizer, img_bgr, quads, dynamic_width=False, batch_bucketing=False, batch_size=128,
              width_budget=None, max_batch_size=None, return_aux=False):
    """reference TextRecognizer.__call__ (text_recognizer.py:352-399) with the oracle PARSeq on the CPU."""
    rgb = img_bgr[:, :, ::-1]
    made = [make_crop(rgb, q, spec.img_size, dynamic_width) for q in quads]
    data = [m for m in made if m is not None]
    canv = [m[0] for m in data]
    cw = [m[1] for m in data]
    order = None
    if batch_bucketing and len(data) == len(quads) and len(data) > 1:
        order = np.argsort(cw).tolist()
    plan = mini_batches([c.shape[1] for c in canv], order, dynamic_width, batch_size, width_budget, max_batch_size)
    pts = [quads[i] for i in order] if order is not None else quads
    preds, scores, dirs, all_logits, margins = [], [], [], [], []
    off = 0
    for batch in plan:
        ts = [to_tensor(canv[i]) for i in batch]
        if dynamic_width:
            wm = max(t.shape[-1] for t in ts)
            ts = [F.pad(t, (0, wm - t.shape[-1]), value=-1.0) for t in ts]
        logits, paux = ops.parseq_forward(sd, spec, torch.stack(ts, 0), return_aux=True)
        s, p = tokenizer.decode(logits.softmax(-1))
        # smallest top-2 logit gap over every greedy decision the row's string depends on (AR steps + the positions
        # of the final logits up to and including the first EOS): rows below a tolerance are coin flips for any
        # implementation that is not bit-identical fp32
        top2 = logits.topk(2, -1).values
        gap = top2[..., 0] - top2[..., 1]
        ids_ = logits.argmax(-1)
        for b in range(logits.shape[0]):
            row = ids_[b].tolist()
            m = row.index(spec.eos_id) + 1 if spec.eos_id in row else len(row)
            margins.append(min(float(paux["ar_margin"][b].min()), float(gap[b, :m].min())))
        preds += [unicodedata.normalize("NFKC", t) for t in s]
        scores += p
        for q in pts[off:off + len(batch)]:
            q = np.array(q)
            dirs.append("vertical" if np.linalg.norm(q[1] - q[2]) > 2 * np.linalg.norm(q[0] - q[1]) else "horizontal")
        off += len(batch)
        all_logits.append(logits)
    if order is not None:
        inv = np.argsort(order)
        preds, scores, dirs = [preds[i] for i in inv], [scores[i] for i in inv], [dirs[i] for i in inv]
        margins = [margins[i] for i in inv]
    if return_aux:
        return preds, scores, dirs, {"plan": plan, "order": order, "logits": all_logits, "canvases": canv,
                                     "min_margin": margins}
    return preds, scores, dirs


def detect(sd_det, img_bgr, post=None):
    """reference TextDetector.__call__ (text_detector.py:112-146) with the oracle DBNet on the CPU."""
    x = detector_preprocess(img_bgr)
    prob = odb.dbnet_forward(sd_det, x)[0, 0].numpy()
    return dbnet_postprocess(prob, img_bgr.shape[:2], **(post or {}))


def ocr_page(sd_det, sd_rec, spec, tokenizer, img_bgr, quads=None, **rec_kw):
    """reference OCR.__call__ (ocr.py:51-63).  `quads` overrides the detector output (synthetic pages with random
    detector weights feed the ground-truth boxes to the recognizer)."""
    dq, ds = detect(sd_det, img_bgr)
    use = dq if quads is None else quads
    p, s, d = recognize(sd_rec, spec, tokenizer, img_bgr, use, **rec_kw)
    return [{"points": q, "content": c, "direction": dd, "rec_score": sc} for q, c, dd, sc in zip(use, p, d, s)], (dq, ds)
