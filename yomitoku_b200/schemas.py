"""Result types of the hot path (pydantic), mirroring reference src/yomitoku/schemas/document_analyzer.py:137-180,
234-254 and BaseSchema (base.py:51-57): extra fields forbidden, assignment validated."""
from typing import List, Union

from pydantic import BaseModel, ConfigDict, Field, conlist

Quad = conlist(conlist(int, min_length=2, max_length=2), min_length=4, max_length=4)


class BaseSchema(BaseModel):
    model_config = ConfigDict(extra="forbid", validate_assignment=True)

    def to_json(self, out_path: str, **kwargs):
        import json
        with open(out_path, "w", encoding="utf-8") as f:
            json.dump(self.model_dump(), f, ensure_ascii=False, indent=kwargs.get("indent", 4))


class WordPrediction(BaseSchema):
    points: Quad = Field(..., description="[[x1, y1], [x2, y2], [x3, y3], [x4, y4]]")
    content: str = Field(..., description="Text content of the word")
    direction: str = Field(..., description="'horizontal' or 'vertical'")
    rec_score: float = Field(..., description="Confidence score of the word recognition")
    det_score: float = Field(..., description="Confidence score of the word detection")


class TextDetectorSchema(BaseSchema):
    points: List[Quad] = Field(..., description="Detected text regions as quadrilaterals")
    scores: List[float] = Field(..., description="Confidence score per region")


class TextRecognizerSchema(BaseSchema):
    contents: List[str] = Field(..., description="Recognized text contents")
    directions: List[str] = Field(..., description="'horizontal' or 'vertical' per text")
    scores: List[float] = Field(..., description="Confidence score per text")
    points: List[Quad] = Field(..., description="Quadrilaterals of the recognized texts")


class OCRSchema(BaseSchema):
    words: List[WordPrediction] = Field(..., description="Recognized words")


class ParagraphSchema(BaseSchema):
    box: conlist(int, min_length=4, max_length=4)
    contents: Union[str, None]
    direction: Union[str, None]
    order: Union[int, None]
    role: Union[str, None]


class DocumentAnalyzerSchema(BaseSchema):
    """reference DocumentAnalyzerSchema; tables/figures come from the RT-DETRv2 layout models, which are outside the
    hot path - they are empty unless a layout analyzer is plugged in."""
    paragraphs: List[ParagraphSchema] = Field(default_factory=list)
    tables: List[dict] = Field(default_factory=list)
    words: List[WordPrediction] = Field(default_factory=list)
    figures: List[dict] = Field(default_factory=list)
