"""LayoutAnalyzer = LayoutParser + TableStructureRecognizer (reference src/yomitoku/layout_analyzer.py:7-49): the
`layout_analyzer` DocumentAnalyzer builds by default when a GPU is present."""
from .layout_parser import LayoutParser
from .schemas import LayoutAnalyzerSchema
from .table_structure_recognizer import TableStructureRecognizer


class LayoutAnalyzer:
    def __init__(self, configs={}, device="cuda", visualize=False):
        if not isinstance(configs, dict):
            raise ValueError("configs must be a dict. See the https://kotaro-kinoshita.github.io/yomitoku-dev/usage/")
        parser_kw = {"device": device, "visualize": visualize, **configs.get("layout_parser", {})}
        table_kw = {"device": device, "visualize": visualize, **configs.get("table_structure_recognizer", {})}
        self.layout_parser = LayoutParser(**parser_kw)
        self.table_structure_recognizer = TableStructureRecognizer(**table_kw)

    def __call__(self, img):
        layout, vis = self.layout_parser(img)
        tables, vis = self.table_structure_recognizer(img, [t.box for t in layout.tables], vis=vis)
        return LayoutAnalyzerSchema(paragraphs=layout.paragraphs, tables=tables, figures=layout.figures), vis

    def analyze_pages(self, pages):
        """Batched entry (new surface): every page's layout in one device call, then one table batch per page."""
        out = []
        for page, layout in zip(pages, self.layout_parser.parse_pages(pages)):
            tables, _ = self.table_structure_recognizer(page, [t.box for t in layout.tables])
            out.append(LayoutAnalyzerSchema(paragraphs=layout.paragraphs, tables=tables, figures=layout.figures))
        return out
