"""CPU: the oracle against fixtures produced by the REFERENCE's own code (tests/golden/make_golden.py), and - when
/root/reference is present - directly against the reference modules (oracle/refcheck.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import dbnet as odb
from oracle import parseq as ops
from oracle import pipeline as opipe
from oracle import refcheck, weights

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_dbnet_oracle_matches_reference_fixture():
    z = np.load(os.path.join(G, "dbnet_ref.npz"))
    sd = weights.make_dbnet_state_dict(seed=int(z["weight_seed"]))
    y = odb.dbnet_forward(sd, torch.from_numpy(z["x"]))
    assert y.shape == (1, 1, 64, 96)
    assert np.abs(y.numpy() - z["prob"]).max() < 1e-6


@pytest.mark.parametrize("tag,kw", [("peaked", dict(peaked=True)), ("repeat", dict(peaked=True, degenerate_repeat=True)),
                                    ("random", dict())])
def test_parseq_oracle_matches_reference_fixture(tag, kw, charset_v2):
    z = np.load(os.path.join(G, "parseq_ref_%s.npz" % tag), allow_pickle=True)
    spec = ops.SPECS["parseq-tiny-dynw-v4"]
    sd = weights.make_parseq_state_dict(spec, seed=int(z["weight_seed"]), **kw)
    lg = ops.parseq_forward(sd, spec, torch.from_numpy(z["img"]))
    assert lg.shape == (3, 101, spec.num_classes)
    assert np.abs(lg[:, 0].numpy() - z["logits_pos0"]).max() < 5e-4
    assert np.array_equal(lg.argmax(-1).numpy(), z["ids"])
    strings, scores = ops.Tokenizer(charset_v2).decode(lg.softmax(-1))
    assert strings == list(z["strings"])
    assert np.allclose(scores, z["scores"], rtol=2e-3, atol=1e-30)
    if tag == "repeat":
        # the repetition stop must have fired: strings are one repeated unit long, not 100 tokens
        assert all(len(s) < 10 for s in strings)


def test_host_functions_match_reference_fixture():
    z = np.load(os.path.join(G, "host_ref.npz"))
    for (h, w), (rh, rw) in zip(z["sizes"], z["resized"]):
        assert opipe.detector_input_size(int(h), int(w)) == (int(rh), int(rw))
    page = z["page"]
    for i, q in enumerate(z["quads"].tolist()):
        for dyn, key in ((False, "fixed%d" % i), (True, "dyn%d" % i)):
            made = opipe.make_crop(page, q, (32, 800), dynamic_width=dyn)
            if int(z["valid%d" % i]) == 0:
                assert made is None
            else:
                assert np.array_equal(made[0], z[key])
    x = opipe.detector_preprocess(page[:, :, ::-1].copy(), shortest=1280, limit=1600)
    assert x.shape[1] == 3 and x.shape[2] % 32 == 0 and x.shape[3] % 32 == 0
    # standardisation arithmetic (reference standardization_image on a float image)
    std = z["std"]
    mine = ((page.astype(np.float32)[:, :, ::-1] / 255.0 - np.array((0.485, 0.456, 0.406))) /
            np.array((0.229, 0.224, 0.225))).astype(np.float32)
    assert np.array_equal(mine, std)


def test_repeat_detector_cases():
    f = ops.detect_repeat_onset
    assert f([1, 2, 3, 4]) is None
    assert f([5] * 8) == (0, 1)
    assert f([9, 5, 5, 5, 5, 5]) is None                 # 5 equal tokens: no period qualifies yet
    assert f([9, 5, 5, 5, 5, 5, 5, 5]) == (2, 2)         # ... but 6 of them are a period-2 unit repeated 3 times
    assert f([9] + [5] * 8) == (1, 1)
    assert f([7, 1, 2, 1, 2, 1, 2]) == (1, 2)              # period 2 repeated 3 times
    assert f([1, 2, 3, 1, 2, 3, 1, 2, 3]) == (0, 3)
    assert f([1, 2, 1, 2]) is None


@pytest.mark.skipif(not refcheck.available(), reason="reference tree not present")
def test_oracle_against_reference_modules_live():
    assert refcheck.main() == 0
