"""Test helper: the trained `decoder.binarize` head of the seeded random DBNet (tests/golden/dbnet_head_trained.npz,
written by tests/golden/make_dbnet_trained_head.py).  With it the detector's own probability map contains the text
lines of the synthetic pages (~200 boxes per page instead of >1000 junk components of a random head)."""
import os

import numpy as np
import torch

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dbnet_head_trained.npz")


def trained_head():
    z = np.load(PATH)
    return {k: torch.from_numpy(z[k]) for k in z.files}


def load_trained_head(model):
    """model: yomitoku_b200.models.DBNet (or anything with load_state_dict(strict=False))."""
    model.load_state_dict(trained_head(), strict=False)
    return model
