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


def peaked_parseq_state_dict(sd, eos_id=0, head_gain=6.0, seed=0):
    """Turns a seeded random PARSeq state_dict (models._parseq_random_state_dict: the reference's init scheme, which
    never emits EOS) into a trained-LIKE one for benchmarks: O(1) activations through the encoder / decoder, a head
    whose logits are far from uniform, and a positional ramp towards the EOS class so that greedy decoding stops at
    varied, image-dependent lengths (~5-40 tokens) like real text lines do.  Same construction as the test weights of
    oracle/weights.py (which the product must not import); the values are synthetic, only their scale is realistic."""
    import math

    import torch
    g = torch.Generator().manual_seed(seed)
    out = {k: v.clone() for k, v in sd.items()}
    D = out["pos_queries"].shape[-1]
    wstd = 1.0 / math.sqrt(D)
    for k, v in out.items():
        if not torch.is_floating_point(v) or v.dim() < 2 or k.endswith("pos_embed") or k == "pos_queries":
            continue
        if k.endswith("patch_embed.proj.weight") or k == "head.weight":
            continue
        fan_in = v.shape[-1]
        gain = 0.5 if (k.endswith("proj.weight") and "attn" in k) or k.endswith("fc2.weight") else 1.0
        out[k] = torch.randn(v.shape, generator=g).clamp_(-2, 2) * (gain / math.sqrt(fan_in))
    out["encoder.pos_embed"] = 0.2 * torch.randn(out["encoder.pos_embed"].shape, generator=g)
    C = out["head.weight"].shape[0]
    out["head.weight"] = torch.randn(C, D, generator=g) * (head_gain / math.sqrt(D))
    out["text_embed.embedding.weight"] = torch.randn(out["text_embed.embedding.weight"].shape, generator=g) * wstd
    S = out["pos_queries"].shape[1]
    pq = torch.randn(1, S, D, generator=g) * 0.5
    w_eos = out["head.weight"][eos_id] / out["head.weight"][eos_id].norm()
    ramp = (torch.arange(S, dtype=torch.float32) - 6.0) * 0.35
    out["pos_queries"] = pq + ramp[None, :, None] * w_eos[None, None, :] * math.sqrt(D) * 0.5
    return out
