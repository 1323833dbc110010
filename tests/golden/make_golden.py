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
    std = fn.standardization_image(page.astype(np.float32))
    np.savez_compressed(os.path.join(OUT, "host_ref.npz"), sizes=np.array(sizes), resized=np.array(res), page=page,
                        quads=np.array(quads), std=std, **crops)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
