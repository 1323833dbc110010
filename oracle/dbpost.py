"""Oracle for the device-side front half of the DBNet post-processing (TEST INFRASTRUCTURE, see __init__.py).

What csrc/dbpost_ops.cu computes, restated with scipy.ndimage on the host: for prob > thresh (reference
postprocessor/dbnet_postporcessor.py:29-30) the 8-connected components (what cv2.findContours walks at :45-47), per
component the raster index of its first pixel, its horizontal runs with the fp64 sum of prob over each run, and the
number of holes (background regions, 4-connected, that do not touch the border) - the quantity that decides whether
OpenCV would list extra (hole) contours for the page.
"""
import numpy as np
from scipy import ndimage as ndi

RUN_DTYPE = np.dtype([("root", "<i4"), ("y", "<i4"), ("x0", "<i4"), ("x1", "<i4"), ("sum", "<f8")])


def post_front(prob, thresh):
    """prob (H, W) float32 -> (runs sorted by (root, y, x0), number of components, number of holes)."""
    bm = prob > thresh
    H, W = bm.shape
    lab, n_comp = ndi.label(bm, structure=np.ones((3, 3), np.int8))
    idx = np.flatnonzero(bm.ravel())
    root = np.full(n_comp + 1, -1, np.int64)
    root[lab.ravel()[idx][::-1]] = idx[::-1]                      # the last write per label is its smallest index
    edge = np.diff(np.pad(bm.astype(np.int8), ((0, 0), (1, 1))), axis=1)
    ys, xs0 = np.nonzero(edge == 1)
    _, xs1 = np.nonzero(edge == -1)
    runs = np.zeros(len(ys), RUN_DTYPE)
    runs["y"], runs["x0"], runs["x1"] = ys, xs0, xs1 - 1
    runs["root"] = root[lab[ys, xs0]]
    for k in range(len(runs)):                                     # left-to-right fp64 accumulation like the kernel
        runs["sum"][k] = np.add.accumulate(prob[ys[k], xs0[k]:xs1[k]].astype(np.float64))[-1]
    bg, n_bg = ndi.label(~bm, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    border = set(np.unique(np.concatenate([bg[0], bg[-1], bg[:, 0], bg[:, -1]]))) - {0}
    return np.sort(runs, order=("root", "y", "x0")), n_comp, n_bg - len(border)
