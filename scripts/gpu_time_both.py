"""Times DBNet (8 pages) and PARSeq-large (3200 crops x 184) device-resident; prints one line (for A/B env toggles)."""
import ctypes, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yomitoku_b200 import TextDetector, TextRecognizer, _lib
L = _lib.lib()
det = TextDetector(from_pretrained=False, device="cuda")
rec = TextRecognizer(model_name="parseq-large-v4_1", from_pretrained=False, device="cuda", dynamic_width=True, batch_bucketing=True)
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pages = torch.from_numpy(rng.integers(0, 256, size=(n, 1200, 1600, 3), dtype=np.uint8)).cuda()
out = torch.empty(n, 1184, 1600, device="cuda")
h = det.model._ensure()
for _ in range(3):
    L.ytk_dbnet_forward_u8(h, pages.data_ptr(), 1, n, 1200, 1600, out.data_ptr(), 1, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    L.ytk_dbnet_forward_u8(h, pages.data_ptr(), 1, n, 1200, 1600, out.data_ptr(), 1, None)
e1.record(); torch.cuda.synchronize()
det_ms = e0.elapsed_time(e1) / 10 / n
B = 3200
canv = [rng.integers(0, 256, size=(32, 184, 3), dtype=np.uint8) for _ in range(B)]
groups = [i // 128 for i in range(B)]
buf, total, descs, ntok = rec.model.pack_crops(canv, [184] * B, groups)
bufd = buf.cuda()
for _ in range(2):
    rec.model.run_packed(bufd, total, descs, B, groups[-1] + 1)
ph = rec.model.last_phase_ms()
print("EPI=%s det_ms_per_page=%.3f parseq_phase_ms=%s" % (os.environ.get("YTK_EPI", "default"), det_ms, {k: round(v, 2) for k, v in ph.items()}), flush=True)
