"""Deterministic synthetic pages for tests and benchmarks.

A page is a white 1200x1600 (H x W) BGR image with 200 axis-aligned text-line boxes on a 40-row x 5-slot grid (row
pitch 29 px, box height 24 px, slot pitch 316 px); box widths are log-normal with a median of 120 px clipped to
[24, 300] (the reference's "median ~120px" remark, cli/main.py:508); each box is filled with random dark vertical
strokes.  `portrait=True` returns the 1600x1200 variant of BASELINE.json's wording (1600 tall, 1200 wide).
"""
import numpy as np


def synthetic_page(page_idx=0, n_rows=40, n_slots=5, height=1200, width=1600, portrait=False):
    """Returns (img uint8 BGR (H,W,3), quads list of [[x,y]*4] clockwise from top-left)."""
    rng = np.random.default_rng(1234 + page_idx)
    if portrait:
        height, width = width, height
        n_rows, n_slots = int(n_rows * 4 / 3), max(1, int(n_slots * 3 / 4))
    img = np.full((height, width, 3), 255, dtype=np.uint8)
    quads = []
    row_pitch, box_h = 29, 24
    slot_pitch = (width - 20) // n_slots
    for r in range(n_rows):
        y = 10 + r * row_pitch
        if y + box_h >= height:
            break
        for s in range(n_slots):
            w = int(np.clip(rng.lognormal(mean=np.log(120.0), sigma=0.5), 24, min(300, slot_pitch - 8)))
            x = 10 + s * slot_pitch
            # strokes
            xx = x + 1
            while xx < x + w - 2:
                bw = int(rng.integers(2, 5))
                top = y + int(rng.integers(1, 6))
                bot = y + box_h - int(rng.integers(1, 6))
                img[top:bot, xx:min(xx + bw, x + w - 1)] = int(rng.integers(0, 81))
                xx += bw + int(rng.integers(1, 4))
            quads.append([[x, y], [x + w, y], [x + w, y + box_h], [x, y + box_h]])
    return img, quads


def synthetic_prob_map(quads, hw, page_hw, blur=5):
    """Blurred ground-truth mask at network resolution: a stand-in probability map with known rectangles (used to
    exercise / time the DBNet post-processor, because random detector weights do not produce ~200 boxes)."""
    import cv2
    H, W = hw
    sy, sx = H / page_hw[0], W / page_hw[1]
    m = np.zeros((H, W), dtype=np.float32)
    for q in quads:
        x0, y0 = q[0]
        x1, y1 = q[2]
        # shrink like DB training targets so that unclip grows the box back to ~the original
        dx, dy = 3, 3
        m[int((y0 + dy) * sy):int((y1 - dy) * sy), int((x0 + dx) * sx):int((x1 - dx) * sx)] = 1.0
    return cv2.GaussianBlur(m, (blur, blur), 0) * 0.9 + 0.02
