import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import parseq as ops, weights
from yomitoku_b200 import TextRecognizer
name = "parseq-large-v4_1"
spec = ops.SPECS[name]
sd = weights.make_parseq_state_dict(spec, seed=4, peaked=True)
rec = TextRecognizer(model_name=name, from_pretrained=False, device="cuda", dynamic_width=True, batch_bucketing=True)
rec.model.load_state_dict(sd)
img = torch.rand(8, 3, 32, 160, generator=torch.Generator().manual_seed(5)) * 2 - 1
got = rec.model(img)
ref, aux = ops.parseq_forward(sd, spec, img, return_aux=True)
print("env", {k: v for k, v in os.environ.items() if k.startswith("YTK")}, "steps", aux["ar_steps"])
for b in range(8):
    d = (got[b] - ref[b]).abs()
    pos = int(d.max(-1).values.argmax())
    same = torch.equal(got[b].argmax(-1), ref[b].argmax(-1))
    row = ref[b].argmax(-1).tolist()
    n = row.index(0) + 1 if 0 in row else 101
    print("row", b, "same_ids", same, "len", n, "max|d| %.3f at pos %d" % (d.max().item(), pos),
          "d within len %.3f" % d[:n].max().item(), "ref absmax at pos %.1f" % ref[b, pos].abs().max().item(),
          "min AR margin %.3f" % aux["ar_margin"][b].min().item())
