"""Stage-by-stage comparison of the RT-DETRv2 device engine with the fp32 oracle (prints errors; used to set the
tolerances of tests/test_gpu_rtdetr.py) + timing: python scripts/gpu_probe_rtdetr.py [layout|table] [batch]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_rtdetr import rtdetr_input  # noqa: E402
from oracle import rtdetr as R  # noqa: E402
from yomitoku_b200.config import LayoutParserRTDETRv2V2Config, TableStructureRecognizerRTDETRv2Config, to_config  # noqa: E402
from yomitoku_b200.models import RTDETRv2  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "layout"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
spec = R.SPECS[kind]
cfg = to_config((LayoutParserRTDETRv2V2Config if kind == "layout" else TableStructureRecognizerRTDETRv2Config)())
sd = R.make_state_dict(spec, seed=11 if kind == "layout" else 12)
m = RTDETRv2(cfg=cfg)
m.load_state_dict(sd)
m.to("cuda")
x = rtdetr_input(21 if kind == "layout" else 22, n=n)
if os.environ.get("RT_ONCE"):
    # for ncu launch lists: two forwards from resident inputs, nothing else
    from yomitoku_b200 import _lib
    xd = x.cuda()
    m(xd)
    c0 = _lib.lib().ytk_launch_count()
    m(xd)
    torch.cuda.synchronize()
    print("launches_per_forward %d" % (_lib.lib().ytk_launch_count() - c0))
    sys.exit(0)
aux = {}
t0 = time.time()
ref = R.forward(sd, spec, x, aux)
print("oracle %.2f s" % (time.time() - t0))
out = m(x)
torch.cuda.synchronize()


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12))


def nchw(name):
    return m.debug_tensor(n, name).transpose(0, 3, 1, 2)


for i, nm in enumerate(("c3", "c4", "c5")):
    print(nm, "rel err %.5f" % rel(nchw(nm), aux["backbone"][i].numpy()))
for i, nm in enumerate(("enc_out3", "enc_out4", "enc_out5")):
    print(nm, "rel err %.5f" % rel(nchw(nm), aux["encoder"][i].numpy()))
if n == 1:
    mem = m.debug_tensor(n, "memory")[0, 0]
    print("memory rel err %.5f" % rel(mem, aux["memory"][0].numpy()))
sc_dev = m.debug_tensor(n, "enc.scores").reshape(n, -1)
sc_ref = aux["enc_logits"].max(-1).values.numpy()
print("enc scores max |d| %.4f (std of scores %.3f)" % (np.abs(sc_dev - sc_ref).max(), sc_ref.std()))
tk_dev = m.debug_tensor(n, "topk").view(np.int32).reshape(n, -1)
for b in range(n):
    a, c = set(tk_dev[b].tolist()), set(aux["topk"][b].tolist())
    cut = np.sort(sc_ref[b])[-300]
    worst = max([abs(sc_ref[b][i] - cut) for i in a ^ c], default=0.0)
    print("image %d: top-300 overlap %d / 300, differing anchors within %.4f of the cut" % (b, len(a & c), worst))
    pos = {v: i for i, v in enumerate(aux["topk"][b].tolist())}
    rows = [(i, pos[v]) for i, v in enumerate(tk_dev[b].tolist()) if v in pos]
    di, ri = [r[0] for r in rows], [r[1] for r in rows]
    dl = np.abs(out["pred_logits"][b].cpu().numpy()[di] - ref["pred_logits"][b].numpy()[ri])
    db = np.abs(out["pred_boxes"][b].cpu().numpy()[di] - ref["pred_boxes"][b].numpy()[ri])
    print("   matched queries: logits max |d| %.4f mean %.5f; boxes max |d| %.5f mean %.6f" % (dl.max(), dl.mean(), db.max(), db.mean()))
    s_dev = torch.sigmoid(out["pred_logits"][b].cpu())
    s_ref = torch.sigmoid(ref["pred_logits"][b])
    print("   detections > 0.5: device %d oracle %d" % (int((s_dev > 0.5).sum()), int((s_ref > 0.5).sum())))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
xd = x.cuda()
for _ in range(3):
    m(xd)
ev[0].record()
for _ in range(10):
    m(xd)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
fl = m.flops(n)
print("forward batch %d: %.3f ms = %.1f images/s, %.1f GFLOP algorithmic = %.0f TFLOP/s" % (n, ms, n / ms * 1e3, fl / 1e9, fl / ms / 1e9))
