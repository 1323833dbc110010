"""Generates tests/golden/rtdetr_ref.npz and tests/golden/rtdetr_wrappers_ref.json from the reference's OWN files,
executed from /root/reference by path (build container only):

  rtdetr_ref.npz            models/rtdetr.py (+ layers/rtdetr_backbone.py, rtdetr_hybrid_encoder.py,
                            rtdetrv2_decoder.py, activate.py): RTDETRv2(cfg).eval() with the seeded weights of
                            oracle.rtdetr.make_state_dict on a seeded input - pred_logits / pred_boxes, the three
                            backbone and encoder feature maps (means of 8x8 blocks), encoder scores and the top-300
                            anchors - for the layout (6 classes) and the table (3 classes) configuration
  rtdetr_wrappers_ref.json  layout_parser.py (LayoutParser.preprocess / postprocess / filtering_elements) and
                            table_structure_recognizer.py (preprocess / postprocess / extract_cell_elements) around the
                            reference's postprocessor/rtdetr_postprocessor.py and utils/misc.py, fed with seeded fake
                            model outputs; modules these files import but this logic never executes (onnx*, base,
                            configs, models, visualizer, logger) are empty stand-ins.
Usage: python tests/golden/make_golden_rtdetr.py
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refcheck as rc  # noqa: E402
from oracle import rtdetr as R  # noqa: E402


def rtdetr_input(seed, n=1):
    """Seeded page-like input in [0, 1]: smooth background + dark boxes (shared with the tests)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, 20, 20, generator=g)
    x = torch.nn.functional.interpolate(x, size=(640, 640), mode="bilinear", align_corners=False) * 0.3 + 0.6
    for b in range(n):
        for _ in range(12):
            x0, y0 = (torch.randint(0, 520, (2,), generator=g)).tolist()
            w, h = (torch.randint(20, 120, (2,), generator=g)).tolist()
            x[b, :, y0:y0 + h, x0:x0 + w] = torch.rand(3, 1, 1, generator=g) * 0.4
    return x.contiguous()


def pooled(t):
    """(1, C, H, W) -> (C, H/8, W/8) block means: a compact fingerprint of a feature map."""
    return torch.nn.functional.avg_pool2d(t, 4 if t.shape[-1] <= 20 else 8)[0].numpy()


def fake_preds(seed, num_classes, n_boxes=14):
    """Model outputs that decode to a handful of confident, partly nested boxes."""
    rng = np.random.default_rng(seed)
    logits = np.full((1, 300, num_classes), -6.0, np.float32)
    boxes = rng.uniform(0.05, 0.95, (1, 300, 4)).astype(np.float32)
    for q in range(n_boxes):
        cx, cy = rng.uniform(0.2, 0.8, 2)
        w, h = rng.uniform(0.08, 0.5, 2)
        boxes[0, q] = (cx, cy, w, h)
        logits[0, q, rng.integers(0, num_classes)] = rng.uniform(0.5, 4.0)
        if q % 3 == 0 and q + 150 < 300:                        # a smaller box inside box q, same or another class
            boxes[0, q + 150] = (cx, cy, w * 0.6, h * 0.6)
            logits[0, q + 150, rng.integers(0, num_classes)] = rng.uniform(0.5, 4.0)
    return {"pred_logits": torch.from_numpy(logits), "pred_boxes": torch.from_numpy(boxes)}


def model_cases():
    out = {}
    for kind in ("layout", "table"):
        spec = R.SPECS[kind]
        sd = R.make_state_dict(spec, seed=11 if kind == "layout" else 12)
        net = rc.build_reference_rtdetr(spec.num_classes, sd)
        x = rtdetr_input(21 if kind == "layout" else 22)
        with torch.no_grad():
            feats = net.backbone(x)
            enc = net.encoder(feats)
            res = net.decoder(enc)
            memory, shapes = net.decoder._get_encoder_input(enc)
            om = net.decoder.enc_output(net.decoder.valid_mask.to(memory.dtype) * memory)
            scores = net.decoder.enc_score_head(om).max(-1).values[0]
        out[kind + "_logits"] = res["pred_logits"][0].numpy()
        out[kind + "_boxes"] = res["pred_boxes"][0].numpy()
        for i in range(3):
            out["%s_c%d" % (kind, i + 3)] = pooled(feats[i])
            out["%s_e%d" % (kind, i + 3)] = pooled(enc[i])
        out[kind + "_enc_scores"] = scores.numpy()
        out[kind + "_topk"] = torch.topk(scores, 300).indices.numpy().astype(np.int32)
    return out


def load_reference_wrappers():
    import importlib.machinery
    rc.load_reference_rtdetr()

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    added = [n for n in ("onnx", "onnxruntime") if n not in sys.modules]
    for n in added:
        stub(n)

    class Schema(dict):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.__dict__.update(kw)

    class Catalog:
        def __init__(self):
            pass

        def register(self, *a):
            pass

    rc._pkg("ytk_ref.utils")
    stub("ytk_ref.constants", ROOT_DIR="/nonexistent")
    stub("ytk_ref.base", BaseModelCatalog=Catalog, BaseModule=object, load_config=None)
    stub("ytk_ref.configs", LayoutParserRTDETRv2Config=None, LayoutParserRTDETRv2V2Config=None,
         TableStructureRecognizerRTDETRv2Config=None)
    sys.modules["ytk_ref.models"].RTDETRv2 = None
    sys.modules["ytk_ref.postprocessor"].RTDETRPostProcessor = sys.modules[
        "ytk_ref.postprocessor.rtdetr_postprocessor"].RTDETRPostProcessor
    rc._load("ytk_ref.utils.misc", "utils/misc.py", "ytk_ref.utils")
    stub("ytk_ref.utils.visualizer", layout_visualizer=None, table_visualizer=None)
    stub("ytk_ref.utils.logger", set_logger=lambda *a, **k: None)
    stub("ytk_ref.schemas", LayoutParserSchema=Schema, TableStructureRecognizerSchema=Schema)
    try:
        lp = rc._load("ytk_ref.layout_parser", "layout_parser.py", "ytk_ref")
        ts = rc._load("ytk_ref.table_structure_recognizer", "table_structure_recognizer.py", "ytk_ref")
    finally:
        for n in added:
            sys.modules.pop(n, None)
    return lp, ts


def reference_layout_parser(lp):
    import torchvision.transforms as T
    from yomitoku_b200.config import LayoutParserRTDETRv2V2Config
    cfg = LayoutParserRTDETRv2V2Config()
    p = object.__new__(lp.LayoutParser)
    p._cfg = rc.AttrDict(data=rc.AttrDict(img_size=[640, 640]))
    p.device, p.visualize, p.infer_onnx = "cpu", False, False
    p.postprocessor = lp.RTDETRPostProcessor(num_classes=6, num_top_queries=300)
    p.transforms = T.Compose([T.Resize([640, 640]), T.ToTensor()])
    p.thresh_score = cfg["thresh_score"]
    p.label_mapper = dict(enumerate(cfg["category"]))
    p.role = cfg["role"]
    return p


def reference_table_recognizer(ts):
    import torchvision.transforms as T
    from yomitoku_b200.config import TableStructureRecognizerRTDETRv2Config
    cfg = TableStructureRecognizerRTDETRv2Config()
    t = object.__new__(ts.TableStructureRecognizer)
    t._cfg = rc.AttrDict(data=rc.AttrDict(img_size=[640, 640]))
    t.device, t.visualize, t.infer_onnx = "cpu", False, False
    t.postprocessor = ts.RTDETRPostProcessor(num_classes=3, num_top_queries=300)
    t.transforms = T.Compose([T.Resize([640, 640]), T.ToTensor()])
    t.thresh_score = cfg["thresh_score"]
    t.label_mapper = dict(enumerate(cfg["category"]))
    return t


def table_preds(seed):
    """Outputs that decode to a grid: rows (class 0), columns (class 1), one span (class 2)."""
    rng = np.random.default_rng(seed)
    logits = np.full((1, 300, 3), -6.0, np.float32)
    boxes = rng.uniform(0.05, 0.95, (1, 300, 4)).astype(np.float32)
    nr, nc = int(rng.integers(2, 6)), int(rng.integers(2, 5))
    q = 0
    for r in range(nr):
        boxes[0, q] = (0.5, (r + 0.5) / nr, 0.98, 1.0 / nr)
        logits[0, q, 0] = rng.uniform(0.5, 4)
        q += 1
    for c in range(nc):
        boxes[0, q] = ((c + 0.5) / nc, 0.5, 1.0 / nc, 0.98)
        logits[0, q, 1] = rng.uniform(0.5, 4)
        q += 1
    boxes[0, q] = (1.0 / nc, 0.5 / nr, 2.0 / nc, 1.0 / nr)          # the first two cells of the first row
    logits[0, q, 2] = 3.0
    boxes[0, q + 1] = (0.5, 0.5 / nr, 0.9, 0.8 / nr)                # a duplicate row inside row 0: filtered
    logits[0, q + 1, 0] = 0.2
    return {"pred_logits": torch.from_numpy(logits), "pred_boxes": torch.from_numpy(boxes)}


def plain(obj):
    if isinstance(obj, dict):
        return {k: plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [plain(v) for v in obj]
    if isinstance(obj, np.generic):
        return obj.item()
    return obj


def wrapper_cases():
    lp, ts = load_reference_wrappers()
    parser, table = reference_layout_parser(lp), reference_table_recognizer(ts)
    out = {"layout": [], "table": []}
    for seed in range(6):
        res = parser.postprocess(fake_preds(100 + seed, 6), (1200 + 40 * seed, 1600 - 30 * seed))
        out["layout"].append({"seed": 100 + seed, "size": [1200 + 40 * seed, 1600 - 30 * seed], "result": plain(dict(res))})
    rng = np.random.default_rng(5)
    page = rng.integers(0, 255, (700, 900, 3), dtype=np.uint8)
    out["preprocess_page_sum"] = float(parser.preprocess(page).double().sum())
    out["preprocess_page_probe"] = parser.preprocess(page)[0, :, ::97, ::89].numpy().tolist()
    for seed in range(6):
        box = [40 + seed, 60, 700 - 10 * seed, 520 + seed]
        data = table.preprocess(page, [box])[0]
        res = table.postprocess(table_preds(200 + seed), data)
        out["table"].append({"seed": 200 + seed, "box": box, "tensor_sum": float(data["tensor"].double().sum()),
                             "result": plain(dict(res))})
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "rtdetr_ref.npz"), **model_cases())
    with open(os.path.join(HERE, "rtdetr_wrappers_ref.json"), "w") as f:
        json.dump(wrapper_cases(), f)
    print("wrote rtdetr_ref.npz, rtdetr_wrappers_ref.json")
