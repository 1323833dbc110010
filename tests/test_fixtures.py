"""CPU: the committed fixtures that GPU tests compare against are what the oracle produces today.

* tests/golden/identity_<model>.npz (oracle outputs for the 2 x 2048 crops of test_gpu_parseq_identity.py): the first 16
  crops of the first group are re-run through the oracle here and must reproduce the stored ids / probabilities / margins.
* tests/golden/dbnet_head_trained.npz (binarize head trained on the seeded backbone): loaded into the seeded detector
  weights, the ORACLE'S own probability map of a held-out synthetic page must contain the page's text lines - the
  premise of test_gpu_dbnet.py::test_polygons_from_the_devices_own_map."""
import os

import numpy as np
import pytest
import torch

import test_gpu_parseq_identity as T
from oracle import dbnet as odb
from oracle import parseq as ops
from oracle import pipeline as opipe
from oracle import weights
from trained_head import trained_head
from yomitoku_b200.config import TextDetectorDBNetV2_1Config
from yomitoku_b200.models import _dbnet_random_state_dict
from yomitoku_b200.postprocessor import DBnetPostProcessor
from yomitoku_b200.synth import synthetic_page


@pytest.mark.parametrize("case", T.CASES, ids=[c[0] for c in T.CASES])
def test_identity_fixture_is_the_oracle(case):
    name, n_groups, gsize, wlo, whi, seed = case
    z = np.load(os.path.join(T.GOLDEN, "identity_%s.npz" % name))
    assert tuple(int(v) for v in z["case"]) == case[1:]
    n = n_groups * gsize
    assert z["ids"].shape == (n, 101) and z["prob"].shape == (n, 101) and z["margin"].shape == (n, 101)
    assert z["ar_margin_min"].shape == (n,)
    spec = ops.SPECS[name]
    sd = weights.make_parseq_state_dict(spec, seed=seed, peaked=True)
    canv, padded, _ = T.make_inputs(n_groups, gsize, wlo, whi, seed)
    k = 16                                           # rows are independent: a slice of the group gives the same rows
    ids, prob, margin, ar = T._oracle_group(sd, spec, canv[:k], padded[0])
    assert np.array_equal(ids, z["ids"][:k].astype(np.int64))
    assert np.allclose(prob, z["prob"][:k], atol=2e-4)
    m = z["margin"][:k]
    assert np.allclose(margin, m, atol=2e-3 + 1e-3 * np.abs(m))
    assert np.allclose(ar.min(-1), z["ar_margin_min"][:k], atol=2e-3)


def test_trained_head_makes_the_oracles_map_a_real_one():
    sd = _dbnet_random_state_dict(0)
    head = trained_head()
    for k, v in head.items():
        assert k.startswith("decoder.binarize.") and tuple(v.shape) == tuple(sd[k].shape), k
    sd.update(head)
    page, quads = synthetic_page(2)
    with torch.inference_mode():
        prob = odb.dbnet_forward(sd, opipe.detector_preprocess(page))[0, 0].numpy()
    pp = DBnetPostProcessor(**TextDetectorDBNetV2_1Config()["post_process"])
    q, s = pp({"binary": prob[None, None]}, page.shape[:2])
    assert 150 <= len(q) <= 400, len(q)

    def rects(qs):
        a = np.asarray(qs, dtype=np.float64).reshape(len(qs), 4, 2)
        return np.stack([a[:, :, 0].min(1), a[:, :, 1].min(1), a[:, :, 0].max(1), a[:, :, 1].max(1)], 1)

    A, B = rects(quads), rects(q)
    hit = 0
    for r in A:
        ix = np.clip(np.minimum(B[:, 2], r[2]) - np.maximum(B[:, 0], r[0]), 0, None)
        iy = np.clip(np.minimum(B[:, 3], r[3]) - np.maximum(B[:, 1], r[1]), 0, None)
        inter = ix * iy
        iou = inter / ((B[:, 2] - B[:, 0]) * (B[:, 3] - B[:, 1]) + (r[2] - r[0]) * (r[3] - r[1]) - inter)
        hit += iou.max() > 0.5
    assert hit >= 170, hit                            # of the page's 200 text lines
