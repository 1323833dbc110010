"""Generates tests/golden/layout_ref.json: outputs of the reference's OWN host logic around the hot path, executed from
/root/reference by path (build container only) -
    src/yomitoku/reading_order.py               prediction_reading_order (three directions)
    src/yomitoku/document_analyzer.py           DocumentAnalyzer.aggregate, extract_words_within_element (+ ruby filter),
                                                _split_text_across_cells
    src/yomitoku/schemas/document_analyzer.py   the pydantic result types
with empty stand-ins for the modules those files import but this logic never executes (text_detector, text_recognizer,
layout_analyzer, export, visualizer: they need omegaconf / onnx / pyclipper, not installable offline).
Inputs are seeded random layouts; the product (yomitoku_b200.reading_order / document_analyzer) must reproduce every
output exactly (tests/test_layout_logic.py).  Usage: python tests/golden/make_golden_layout.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = os.environ.get("YTK_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src", "yomitoku")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    """-> (reading_order module, document_analyzer module, schemas module) of the reference, executed by path."""
    from pydantic import BaseModel, ConfigDict

    def mod(name, path=None, **attrs):
        if path is None:
            m = types.ModuleType(name)
        else:
            spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=None)
            m = importlib.util.module_from_spec(spec)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        if path is not None:
            spec.loader.exec_module(m)
        return m

    class BaseSchema(BaseModel):            # reference base.py:51-57 (Config: extra forbid, validate_assignment)
        model_config = ConfigDict(extra="forbid", validate_assignment=True)

    pkg = mod("yomitoku")
    pkg.__path__ = [SRC]
    mod("yomitoku.base", BaseSchema=BaseSchema)
    noop = lambda *a, **k: None
    mod("yomitoku.export", export_csv=noop, export_html=noop, export_markdown=noop, export_json=noop)
    up = mod("yomitoku.utils")
    up.__path__ = [os.path.join(SRC, "utils")]
    mod("yomitoku.utils.graph", os.path.join(SRC, "utils", "graph.py"))
    mod("yomitoku.utils.misc", os.path.join(SRC, "utils", "misc.py"))
    mod("yomitoku.utils.visualizer", det_visualizer=noop, reading_order_visualizer=noop)
    sp = mod("yomitoku.schemas")
    sp.__path__ = [os.path.join(SRC, "schemas")]
    sd = mod("yomitoku.schemas.document_analyzer", os.path.join(SRC, "schemas", "document_analyzer.py"))
    for k in dir(sd):
        if k.endswith("Schema") or k in ("Element", "WordPrediction"):
            setattr(sp, k, getattr(sd, k))
    ro = mod("yomitoku.reading_order", os.path.join(SRC, "reading_order.py"))
    mod("yomitoku.text_detector", TextDetector=object)
    mod("yomitoku.text_recognizer", TextRecognizer=object)
    mod("yomitoku.layout_analyzer", LayoutAnalyzer=object)
    mod("yomitoku.ocr", OCRSchema=sd.OCRSchema, ocr_aggregate=noop)
    da = mod("yomitoku.document_analyzer", os.path.join(SRC, "document_analyzer.py"))
    return ro, da, sd


# ------------------------------------------------------------------------------------------------ random inputs
def random_boxes(rng, n, page=(1600, 1200), kind="mixed"):
    """Column-ish layouts with overlaps, ties and nesting (the cases the precedence graph has to order)."""
    W, H = page
    out = []
    for _ in range(n):
        if kind == "grid" and rng.random() < 0.7:
            x1 = int(rng.choice([50, 420, 800, 1180])) + int(rng.integers(-5, 6))
            y1 = int(rng.choice(np.arange(40, H - 80, 60)))
            w, h = int(rng.integers(120, 360)), int(rng.integers(20, 56))
        else:
            x1, y1 = int(rng.integers(0, W - 60)), int(rng.integers(0, H - 40))
            w, h = int(rng.integers(20, 500)), int(rng.integers(12, 300))
        out.append([x1, y1, min(W, x1 + w), min(H, y1 + h)])
    return out


def random_words(rng, n, page=(1600, 1200)):
    W, H = page
    words = []
    hira, kata, kanji = "あいうえおかきくけこ", "アイウエオカキクケコ", "漢字日本語文章東京"
    for _ in range(n):
        vertical = rng.random() < 0.2
        small = rng.random() < 0.25
        if vertical:
            w, h = int(rng.integers(10, 28)), int(rng.integers(60, 260))
        else:
            w, h = int(rng.integers(40, 360)), (int(rng.integers(7, 12)) if small else int(rng.integers(18, 34)))
        x, y = int(rng.integers(0, W - w)), int(rng.integers(0, H - h))
        chars = hira if small and rng.random() < 0.6 else (kata if small and rng.random() < 0.5 else kanji)
        content = "".join(rng.choice(list(chars), size=int(rng.integers(1, 7))))
        words.append({"points": [[x, y], [x + w, y], [x + w, y + h], [x, y + h]], "content": content,
                      "direction": "vertical" if vertical else "horizontal", "rec_score": float(rng.random()),
                      "det_score": float(rng.random())})
    return words


def random_table(rng, page=(1600, 1200)):
    W, H = page
    x1, y1 = int(rng.integers(40, W // 2)), int(rng.integers(40, H // 2))
    n_row, n_col = int(rng.integers(1, 5)), int(rng.integers(1, 5))
    cw, rh = int(rng.integers(90, 220)), int(rng.integers(40, 110))
    xs = [x1 + c * cw for c in range(n_col + 1)]
    ys = [y1 + r * rh for r in range(n_row + 1)]
    cells, covered = [], set()
    for r in range(n_row):
        for c in range(n_col):
            if (r, c) in covered:
                continue
            rs = 2 if r + 1 < n_row and rng.random() < 0.15 and (r + 1, c) not in covered else 1
            cs = 2 if c + 1 < n_col and rng.random() < 0.15 and all((r + i, c + 1) not in covered for i in range(rs)) else 1
            for i in range(rs):
                for j in range(cs):
                    covered.add((r + i, c + j))
            cells.append({"col": c + 1, "row": r + 1, "col_span": cs, "row_span": rs,
                          "box": [xs[c], ys[r], xs[c + cs], ys[r + rs]], "contents": None})
    rows = [{"box": [xs[0], ys[r], xs[-1], ys[r + 1]], "score": 0.9} for r in range(n_row)]
    cols = [{"box": [xs[c], ys[0], xs[c + 1], ys[-1]], "score": 0.9} for c in range(n_col)]
    return {"box": [xs[0], ys[0], xs[-1], ys[-1]], "n_row": n_row, "n_col": n_col, "rows": rows, "cols": cols,
            "spans": [], "cells": cells, "order": 0}


def random_layout(rng, words):
    roles = [None, None, None, "section_headings", "page_header", "page_footer", "caption"]
    paragraphs = []
    for b in random_boxes(rng, int(rng.integers(0, 7)), kind="grid"):
        paragraphs.append({"id": None, "box": b, "score": 0.8, "role": roles[int(rng.integers(0, len(roles)))],
                           "contents": None})
    # regions drawn around clusters of words so that elements really contain words
    for _ in range(int(rng.integers(1, 5))):
        if not words:
            break
        w = words[int(rng.integers(0, len(words)))]
        x1, y1 = w["points"][0]
        paragraphs.append({"id": None, "box": [max(0, x1 - 30), max(0, y1 - 40), min(1600, x1 + 500), min(1200, y1 + 160)],
                           "score": 0.9, "role": roles[int(rng.integers(0, len(roles)))], "contents": None})
    tables = [random_table(rng) for _ in range(int(rng.integers(0, 3)))]
    figures = [{"id": None, "box": b, "score": 0.7, "role": None, "contents": None}
               for b in random_boxes(rng, int(rng.integers(0, 3)))]
    return {"paragraphs": paragraphs, "tables": tables, "figures": figures}


def main():
    ro, da, sd = load_reference()
    rng = np.random.default_rng(20260923)
    out = {"reading_order": [], "aggregate": [], "words_in_element": [], "split": []}
    # ---- reading order
    for case in range(120):
        direction = ["top2bottom", "right2left", "left2right"][case % 3]
        n = int(rng.integers(2, 14))
        boxes = random_boxes(rng, n, kind="grid" if case % 2 else "mixed")
        if direction != "top2bottom" and any(b[3] - b[1] == 0 for b in boxes):
            continue
        els = [sd.ParagraphSchema(box=b, contents="p%d" % i, direction="horizontal", order=0, role=None)
               for i, b in enumerate(boxes)]
        try:
            ro.prediction_reading_order(els, direction)
        except ZeroDivisionError:
            continue
        out["reading_order"].append({"direction": direction, "boxes": boxes, "order": [e.order for e in els]})
    # ---- words within an element (+ ruby filter)
    for case in range(60):
        words = [sd.WordPrediction(**w) for w in random_words(rng, int(rng.integers(0, 25)))]
        box = random_boxes(rng, 1)[0]
        box = [max(0, box[0] - 200), max(0, box[1] - 200), min(1600, box[2] + 400), min(1200, box[3] + 300)]
        el = sd.Element(id=None, box=box, score=0.5, role=None, contents=None)
        ignore_ruby = bool(case % 2)
        thr = [2.0, 0.5, 1.2][case % 3]
        text, direction, flags = da.extract_words_within_element(words, el, ignore_ruby=ignore_ruby, ruby_threshold=thr)
        out["words_in_element"].append({"words": [w.model_dump() for w in words], "box": box, "ignore_ruby": ignore_ruby,
                                        "ruby_threshold": thr, "text": text, "direction": direction, "flags": flags})
    # ---- aggregate
    for case in range(60):
        wl = random_words(rng, int(rng.integers(0, 40)))
        ll = random_layout(rng, wl)
        opts = {"ignore_meta": bool(case % 2), "reading_order": ["auto", "top2bottom", "right2left", "left2right"][case % 4],
                "ignore_ruby": bool((case // 2) % 2), "ruby_threshold": 2.0}
        self = types.SimpleNamespace(img=None, **opts)
        ocr = sd.OCRSchema(words=[sd.WordPrediction(**w) for w in wl])
        lay = sd.LayoutAnalyzerSchema(**json.loads(json.dumps(ll)))
        try:
            res = da.DocumentAnalyzer.aggregate(self, ocr, lay)
        except ZeroDivisionError:
            continue
        dumped = sd.DocumentAnalyzerSchema(**res).model_dump()
        for f in dumped["figures"]:
            f.pop("figure_path", None)
        out["aggregate"].append({"words": wl, "layout": ll, "options": opts, "result": dumped})
    # ---- split_text_across_cells
    for case in range(40):
        table = random_table(rng)
        x1, y1, x2, y2 = table["box"]
        pts, scores = [], []
        for _ in range(int(rng.integers(1, 12))):
            if rng.random() < 0.3:     # vertical line inside the table
                w, h = int(rng.integers(16, 30)), int(rng.integers(60, max(61, y2 - y1)))
            else:
                w, h = int(rng.integers(40, max(41, x2 - x1))), int(rng.integers(16, 34))
            x = int(rng.integers(max(0, x1 - 60), max(1, x2 - 10)))
            y = int(rng.integers(max(0, y1 - 40), max(1, y2 - 10)))
            pts.append([[x, y], [x + w, y], [x + w, y + h], [x, y + h]])
            scores.append(float(rng.random()))
        det = sd.TextDetectorSchema(points=pts, scores=scores)
        lay = sd.LayoutAnalyzerSchema(paragraphs=[], tables=[json.loads(json.dumps(table))], figures=[])
        res = da._split_text_across_cells(det, lay)
        out["split"].append({"points": pts, "scores": scores, "table": table,
                             "out_points": res.points, "out_scores": res.scores})
    path = os.path.join(HERE, "layout_ref.json")
    json.dump(out, open(path, "w"), ensure_ascii=False)
    print("wrote", path, {k: len(v) for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
