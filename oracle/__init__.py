"""CPU oracle for the yomitoku DBNet -> PARSeq hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain fp32 PyTorch / numpy / OpenCV on the CPU, the algorithm the reference
(kotaro-kinoshita/yomitoku @ ac30118) runs for the path SURVEY.md section 8 scopes: detector pre-processing, DBNet++,
DBNet post-processing, crop extraction, mini-batch formation, PARSeq (encoder, AR decode, refinement, repetition
stop), tokenizer decode.  Every function cites the reference file:line it follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import it, and
only as the checker or the timed CPU baseline.  The product (`yomitoku_b200/`) never imports it: the device path has
no CPU fallback.

Parity pinning: the reference's own tests hold no numeric golden for this path ("parity
unpinned" by the reference).  The oracle is therefore pinned against the reference's own model code instead:
`oracle/refcheck.py` loads the reference's `models/dbnet_plus.py`, `models/parseq.py`,
`models/layers/*.py`, `postprocessor/parseq_tokenizer.py` and `data/functions.py` *by path* from /root/reference
(with small stand-ins for the third-party packages missing in this image) and asserts equality on seeded inputs;
`tests/golden/make_golden.py` stores reference-generated outputs as fixtures that travel to the GPU box.
Beyond the model code, refcheck also executes the reference's own DBnetPostProcessor (pyclipper / shapely replaced by
this oracle's restatements), its ParseqDataset and its TextRecognizer / TextDetector `__call__` methods (around
stand-in models) to pin the host rows - post-processing, crop extraction, batching, pairing, fallback - and
`tests/golden/{post_ref,crops_ref,flow_ref,detflow_ref}.npz` carry those outputs to machines without the reference.
The one native piece, `oracle/crop_host.cpp`, is the PRODUCT's crop arithmetic (csrc/crop_math.h) compiled for the
host so that the CPU tests can pin it bit for bit against OpenCV.
Third-party arithmetic not under /root/reference and absent from the image (pyclipper 1.4.0, shapely 2.1.2,
timm 1.0.27, omegaconf 2.3.0) is restated from its published algorithm; see oracle/postprocess.py and
oracle/parseq.py headers.  Those pieces are "parity unpinned" against the real libraries (they cannot be
installed offline) and say so where they occur.
"""
