"""GPU: the public API end to end (OCR, BatchedOCR, DocumentAnalyzer) on synthetic pages."""
import numpy as np
import pytest

from oracle import parseq as ops
from oracle import weights
from yomitoku_b200 import OCR, DocumentAnalyzer
from yomitoku_b200.pipeline import BatchedOCR
from yomitoku_b200.synth import synthetic_page, synthetic_prob_map

pytestmark = pytest.mark.gpu


def _ocr():
    o = OCR(configs={"text_detector": {"from_pretrained": False},
                     "text_recognizer": {"from_pretrained": False, "model_name": "parseq-tiny-dynw-v4",
                                         "dynamic_width": True, "batch_bucketing": True}}, device="cuda")
    spec = ops.SPECS["parseq-tiny-dynw-v4"]
    o.recognizer.model.load_state_dict(weights.make_parseq_state_dict(spec, seed=11, peaked=True))
    return o


def test_ocr_call_returns_schema():
    o = _ocr()
    page, quads = synthetic_page(0)
    res, vis = o(page)
    assert vis is None
    for w in res.words[:5]:
        assert w.direction in ("horizontal", "vertical") and 0.0 <= w.rec_score <= 1.0


def test_batched_ocr_equals_per_page_calls():
    o = _ocr()
    pages, quads, probs = [], [], []
    for i in range(3):
        p, q = synthetic_page(10 + i)
        pages.append(p)
        quads.append(q)
        probs.append(synthetic_prob_map(q, (1184, 1600), (1200, 1600)))
    b = BatchedOCR(o.detector, o.recognizer, workers=2, det_batch=2)
    try:
        res = b(pages, prob_override=probs)
    finally:
        b.close()
    assert len(res) == 3
    for i in range(3):
        assert len(res[i].words) == len(quads[i])
        # the same page through the one-page recognizer API (same quads): identical strings
        det_points = [w.points for w in res[i].words]
        single, _ = o.recognizer(pages[i], det_points)
        assert [w.content for w in res[i].words] == single.contents
        assert np.allclose([w.rec_score for w in res[i].words], single.scores, atol=1e-6)


def test_stream_equals_batched_calls():
    """BatchedOCR.stream (detector / recognizer / assembly threads, two CUDA streams, staging ring reuse over more
    batches than ring slots) yields, in order, exactly what per-batch calls return."""
    o = _ocr()
    batches, overrides = [], []
    for k in range(5):
        pages, probs = [], []
        for i in range(2):
            p, q = synthetic_page(40 + 2 * k + i)
            pages.append(p)
            probs.append(synthetic_prob_map(q, (1184, 1600), (1200, 1600)))
        batches.append(pages)
        overrides.append(probs)
    b = BatchedOCR(o.detector, o.recognizer, workers=3, det_batch=1)
    try:
        ref = [b(pg, prob_override=po) for pg, po in zip(batches, overrides)]
        got = list(b.stream(batches, lookahead=2, prob_override=overrides))
    finally:
        b.close()
    assert len(got) == len(ref) == 5
    for g, r in zip(got, ref):
        assert [[w.content for w in page.words] for page in g] == [[w.content for w in page.words] for page in r]
        assert [[w.points for w in page.words] for page in g] == [[w.points for w in page.words] for page in r]


def test_document_analyzer_default_layout_models():
    """DocumentAnalyzer as the reference builds it: OCR + LayoutAnalyzer (RT-DETRv2 layout parser and table structure
    recognizer on the device engine), random weights."""
    from yomitoku_b200.layout_analyzer import LayoutAnalyzer
    nop = {"from_pretrained": False}
    da = DocumentAnalyzer(configs={"ocr": {"text_detector": nop,
                                           "text_recognizer": {"from_pretrained": False,
                                                               "model_name": "parseq-tiny-dynw-v4"}},
                                   "layout_analyzer": {"layout_parser": nop, "table_structure_recognizer": nop}},
                          device="cuda")
    assert isinstance(da.layout, LayoutAnalyzer)
    from trained_head import load_trained_head
    load_trained_head(da.text_detector.model)       # the detector's own map holds the page's text lines (~250 boxes)
    page, _ = synthetic_page(1)
    res, ocr_vis, layout_vis = da(page)
    assert layout_vis is None and isinstance(res.words, list) and len(res.words) > 0
    off = DocumentAnalyzer(configs={"ocr": {"text_detector": nop,
                                            "text_recognizer": {"from_pretrained": False,
                                                                "model_name": "parseq-tiny-dynw-v4"}}},
                           device="cuda", layout_analyzer=False)
    assert off.layout is None


def test_handles_bind_to_the_requested_device():
    """`TextDetector(device="cuda:k")` / `TextRecognizer(device="cuda:k")` run on GPU k (reference: model.to(self.device),
    base.py:106-121), not on whatever device the calling thread happens to have current."""
    import torch
    from yomitoku_b200 import TextDetector, TextRecognizer, _lib
    k = torch.cuda.device_count() - 1           # the last GPU: differs from the current device (0) on multi-GPU boxes
    det = TextDetector(from_pretrained=False, device="cuda:%d" % k)
    rec = TextRecognizer(model_name="parseq-tiny-dynw-v4", from_pretrained=False, device="cuda:%d" % k)
    L = _lib.lib()
    assert L.ytk_dbnet_device(det.model._ensure()) == k
    assert L.ytk_parseq_device(rec.model._ensure()) == k
    assert det.model.cuda_device().index == k
    # a forward on GPU k from a thread whose current device is 0: same bits as the same weights on GPU 0 (kernels whose
    # shared-memory attribute was first set on the other device included), and the caller's current device is untouched
    cur = torch.cuda.current_device()
    page, _ = synthetic_page(3)
    pk = det.model.detect_pages_u8(page)
    assert torch.cuda.current_device() == cur
    assert pk.device.type == "cpu" or pk.device.index == k
    det0 = TextDetector(from_pretrained=False, device="cuda:0")
    p0 = det0.model.detect_pages_u8(page)
    assert torch.equal(pk.cpu(), p0.cpu())
    assert torch.cuda.current_device() == cur
    # moving a materialised model drops the handle; the next use re-creates it on the new device
    rec.model.to("cuda:0")
    assert (rec.model._handle is None) == (k != 0)
    assert L.ytk_parseq_device(rec.model._ensure()) == 0
    assert torch.cuda.current_device() == cur


def test_document_analyzer_batched_pages_equal_single_page_calls():
    """DocumentAnalyzer with a plugged-in layout analyzer (stub returning a table + a paragraph region): the batched
    multi-page form `analyze_pages` (BatchedOCR underneath) gives, page by page, what the reference-style one-page
    `__call__` gives - with and without split_text_across_cells (lines cut at the cell borders before recognition)."""
    import numpy as np
    from yomitoku_b200 import DocumentAnalyzer
    from yomitoku_b200 import schemas as S
    from yomitoku_b200.synth import synthetic_page

    def layout(img):
        cells = [S.TableCellSchema(col=c + 1, row=r + 1, col_span=1, row_span=1,
                                   box=[20 + 316 * c, 10 + 58 * r, 20 + 316 * (c + 1), 10 + 58 * (r + 1)], contents=None)
                 for r in range(4) for c in range(3)]
        rows = [S.TableLineSchema(box=[20, 10 + 58 * r, 968, 68 + 58 * r], score=0.9) for r in range(4)]
        cols = [S.TableLineSchema(box=[20 + 316 * c, 10, 336 + 316 * c, 242], score=0.9) for c in range(3)]
        table = S.TableStructureRecognizerSchema(box=[20, 10, 968, 242], n_row=4, n_col=3, rows=rows, cols=cols,
                                                 spans=[], cells=cells, order=0)
        para = S.Element(id=None, box=[0, 600, 1600, 900], score=0.9, role=None, contents=None)
        return S.LayoutAnalyzerSchema(paragraphs=[para], tables=[table], figures=[]), None

    cfg = {"ocr": {"text_detector": {"from_pretrained": False},
                   "text_recognizer": {"from_pretrained": False, "model_name": "parseq-tiny-dynw-v4",
                                       "dynamic_width": True, "batch_bucketing": True}}}
    pages = [synthetic_page(60)[0], synthetic_page(61)[0]]
    for split in (False, True):
        an = DocumentAnalyzer(configs=cfg, device="cuda", layout_analyzer=layout, split_text_across_cells=split)
        from trained_head import load_trained_head
        # trained binarize head: the detector finds the pages' text lines (~250 boxes) instead of the > 1000 junk
        # components a random head yields (the test took 100 s on them)
        load_trained_head(an.text_detector.model)
        single = [an(p)[0] for p in pages]
        batched = an.analyze_pages(pages)
        assert len(batched) == 2
        for a, b in zip(single, batched):
            assert [w.points for w in a.words] == [w.points for w in b.words]
            assert [w.content for w in a.words] == [w.content for w in b.words]
            assert [p.contents for p in a.paragraphs] == [p.contents for p in b.paragraphs]
            assert [p.order for p in a.paragraphs] == [p.order for p in b.paragraphs]
            assert [[c.contents for c in t.cells] for t in a.tables] == [[c.contents for c in t.cells] for t in b.tables]
            assert np.allclose([w.rec_score for w in a.words], [w.rec_score for w in b.words], atol=1e-6)
        an._batched.close()
