"""yomitoku_b200: Blackwell-native DBNet -> PARSeq OCR hot path behind yomitoku's module API.

    from yomitoku_b200 import OCR, TextDetector, TextRecognizer, DocumentAnalyzer

The constructors, the `configs` dict and the call contracts mirror kotaro-kinoshita/yomitoku
(src/yomitoku/{text_detector,text_recognizer,ocr,document_analyzer}.py); the models run as hand-written sm_100a CUDA
kernels behind the C ABI in include/yomitoku_b200.h (libytk_b200.so).
"""
from .document_analyzer import DocumentAnalyzer
from .ocr import OCR
from .text_detector import TextDetector
from .text_recognizer import TextRecognizer

__all__ = ["OCR", "TextDetector", "TextRecognizer", "DocumentAnalyzer"]
