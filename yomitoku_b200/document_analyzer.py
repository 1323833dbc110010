"""DocumentAnalyzer shell: keeps the reference's constructor / configs-dict / call surface
(src/yomitoku/document_analyzer.py:426-678) and routes the OCR half through the device path.

The layout half (RT-DETRv2 layout parser + table structure recognizer, reading order, paragraph aggregation) is
outside this repo's hot-path scope: a layout analyzer object with the reference's
`__call__(img) -> (LayoutAnalyzerSchema-like, vis)` protocol can be plugged in through `layout_analyzer=`; without one
the result carries the words and no paragraphs / tables / figures.
"""
from concurrent.futures import ThreadPoolExecutor

from .ocr import ocr_aggregate
from .schemas import DocumentAnalyzerSchema, OCRSchema
from .text_detector import TextDetector
from .text_recognizer import TextRecognizer


class DocumentAnalyzer:
    def __init__(self, configs={}, device="cuda", visualize=False, ignore_meta=False, reading_order="auto",
                 split_text_across_cells=False, ignore_ruby=False, ruby_threshold=0.5, layout_analyzer=None):
        default_configs = {
            "ocr": {
                "text_detector": {"device": device, "visualize": visualize},
                "text_recognizer": {"device": device, "visualize": visualize},
            },
            "layout_analyzer": {
                "layout_parser": {"device": device, "visualize": visualize},
                "table_structure_recognizer": {"device": device, "visualize": visualize},
            },
        }
        if isinstance(configs, dict):
            _recursive_update(default_configs, configs)
        else:
            raise ValueError("configs must be a dict. See the https://kotaro-kinoshita.github.io/yomitoku-dev/usage/")
        self.text_detector = TextDetector(**default_configs["ocr"]["text_detector"])
        self.text_recognizer = TextRecognizer(**default_configs["ocr"]["text_recognizer"])
        self.layout = layout_analyzer
        self.visualize = visualize
        self.ignore_meta = ignore_meta
        self.reading_order = reading_order
        self.split_text_across_cells = split_text_across_cells
        self.ignore_ruby = ignore_ruby
        self.ruby_threshold = ruby_threshold

    def _detect_and_recognize(self, img):
        det, vis = self.text_detector(img)
        rec, vis = self.text_recognizer(img, det.points, vis=vis)
        return OCRSchema(words=ocr_aggregate(det, rec)), vis

    def __call__(self, img):
        """Returns (DocumentAnalyzerSchema, ocr_vis, layout_vis) like the reference (:671-678).  OCR and layout run in
        two threads on the same device, as in the reference (:622-659); each C handle serialises on its own mutex."""
        with ThreadPoolExecutor(max_workers=2) as ex:
            f_ocr = ex.submit(self._detect_and_recognize, img)
            f_lay = ex.submit(self.layout, img) if self.layout is not None else None
            ocr, ocr_vis = f_ocr.result()
            layout_vis = None
            if f_lay is not None:
                _, layout_vis = f_lay.result()
        return DocumentAnalyzerSchema(words=[w.model_dump() for w in ocr.words]), ocr_vis, layout_vis


def _recursive_update(original, new_data):
    for key, value in new_data.items():
        if isinstance(original.get(key), dict) and isinstance(value, dict):
            _recursive_update(original[key], value)
        else:
            original[key] = value
    return original
