"""CPU: the black-box expectations the reference's own tests pin for the functions on this path (SURVEY.md section 8c:
"re-run against the new package"), here for the host data functions - reference tests/test_data.py:83-174.  Same
inputs, same assertions, against yomitoku_b200.data (config / catalog expectations of tests/test_ocr.py and
tests/test_base.py are in test_host_logic.py)."""
import numpy as np

from yomitoku_b200.data import (array_to_tensor, resize_shortest_edge, resize_with_padding, rotate_text_image,
                                standardization_image, validate_quads)


def test_resize_shortest_edge():            # reference tests/test_data.py:83-101
    img = np.zeros((1920, 1920, 3), dtype=np.uint8)
    h, w = resize_shortest_edge(img, 1280, 1500).shape[:2]
    assert min(h, w) == 1280 and h % 32 == 0 and w % 32 == 0
    img = np.zeros((1280, 1920, 3), dtype=np.uint8)
    h, w = resize_shortest_edge(img, 1280, 1600).shape[:2]
    assert max(h, w) == 1600 and h % 32 == 0 and w % 32 == 0
    h, w = resize_shortest_edge(img, 1000, 1000).shape[:2]
    assert h % 32 == 0 and w % 32 == 0


def test_standardization_image():           # :104-108
    img = np.random.randint(0, 255, (100, 100, 3), dtype=np.uint8)
    normalized = standardization_image(img)
    assert normalized.shape == img.shape and normalized.dtype == "float32"


def test_array_to_tensor():                 # :111-114
    img = np.random.randint(0, 255, (100, 50, 3), dtype=np.uint8)
    assert array_to_tensor(img).shape == (1, 3, 100, 50)


def test_rotate_image():                    # :117-124
    img = np.random.randint(0, 255, (100, 30, 3), dtype=np.uint8)
    assert rotate_text_image(img, thresh_aspect=2).shape == (30, 100, 3)
    img = np.random.randint(0, 255, (30, 100, 3), dtype=np.uint8)
    assert rotate_text_image(img, thresh_aspect=2).shape == (30, 100, 3)


def test_resize_with_padding():             # :127-138
    for shape in ((50, 100, 3), (50, 150, 3), (60, 100, 3)):
        img = np.random.randint(0, 255, shape, dtype=np.uint8)
        assert resize_with_padding(img, (50, 100)).shape == (50, 100, 3)


def test_validate_quads():                  # :141-174
    img = np.random.randint(0, 255, (100, 100, 3), dtype=np.uint8)
    for quad in ([[0, 0], [0, 10], [10, 10]], [[0], [0, 10], [10, 10], [10, 0]],
                 [[0, 0], [0, 150], [10, 150], [10, 0]], [[150, 0], [150, 10], [10, 10], [10, 0]],
                 [[-1, 0], [-1, 10], [10, 10], [10, 0]], [[0, -1], [0, 10], [10, 10], [10, -1]]):
        assert validate_quads(img, quad) is None
    for quad in ([[0, 0], [0, 10], [10, 10], [10, 0]], [[0, 0], [0, 20], [10, 20], [10, 0]],
                 [[10, 0], [10, 30], [80, 30], [80, 0]]):
        assert validate_quads(img, quad)
