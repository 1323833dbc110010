"""GPU probe: PARSeq engine vs the CPU oracle (writes gpurun_out/probe_parseq.json)."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import parseq as ops  # noqa: E402
from oracle import weights  # noqa: E402
from yomitoku_b200 import _lib  # noqa: E402

L = _lib.lib()
out = []


def log(**kw):
    out.append(kw)
    print(json.dumps(kw), flush=True)


def make(spec, sd, refine=None):
    tab, keep = _lib.tensor_table(sd)
    cfg = _lib.YtkParseqCfg(spec.embed_dim, spec.enc_heads, spec.enc_depth, spec.patch[0], spec.patch[1],
                            spec.img_size[0], spec.img_size[1], spec.num_tokens, spec.max_label_length, spec.dec_heads,
                            spec.mlp_ratio, spec.dec_mlp_ratio, spec.refine_iters if refine is None else refine,
                            1 if spec.repetition_stop else 0, spec.rep_period_max, spec.rep_min_run_p1,
                            spec.rep_min_repeats)
    h = ctypes.c_void_p()
    st = L.ytk_parseq_create(tab, len(tab), ctypes.byref(cfg), ctypes.byref(h))
    if st != 0:
        raise RuntimeError(L.ytk_last_error().decode())
    return h


def seam(h, spec, img, want_logits=True, want_mem=True):
    B, W = img.shape[0], img.shape[3]
    S, C, D = spec.max_label_length + 1, spec.num_classes, spec.embed_dim
    ntok = (32 // spec.patch[0]) * (W // spec.patch[1])
    logits = torch.zeros(B, S, C) if want_logits else None
    mem = torch.zeros(B * ntok, D) if want_mem else None
    ids = torch.zeros(B, S, dtype=torch.int32)
    probs = torch.zeros(B, S)
    steps = ctypes.c_int(0)
    rep = torch.zeros(B, dtype=torch.int32)
    t0 = time.time()
    st = L.ytk_parseq_forward_f32(h, img.data_ptr(), 0, B, W, logits.data_ptr() if want_logits else None, 0,
                                  ids.data_ptr(), probs.data_ptr(), ctypes.byref(steps), rep.data_ptr(),
                                  mem.data_ptr() if want_mem else None, None)
    if st != 0:
        raise RuntimeError(L.ytk_last_error().decode())
    return dict(logits=logits, mem=mem, ids=ids, probs=probs, steps=steps.value, rep=rep, wall=time.time() - t0)


def compare(name, spec, sd, img):
    h = make(spec, sd)
    r = seam(h, spec, img)
    t0 = time.time()
    o, aux = ops.parseq_forward(sd, spec, img, return_aux=True)
    cpu_s = time.time() - t0
    B = img.shape[0]
    mem_o = aux["memory"].reshape(-1, spec.embed_dim)
    dm = (r["mem"] - mem_o).abs()
    # step-0 logits are independent of any token feedback: clean numeric comparison
    ids_o = o.argmax(-1)
    top2 = o.topk(2, -1).values
    margin = (top2[..., 0] - top2[..., 1])
    lg = r["logits"]
    for b in range(B):  # apply the repetition patch like the reference does (parseq.py:301-309)
        cut = int(r["rep"][b])
        if cut >= 0 and cut < lg.shape[1]:
            lg[b, cut, :] = -30.0
            lg[b, cut, 0] = 30.0
    dl = (lg - o).abs() if lg.shape == o.shape else None
    same_rows = [(r["ids"][b].long() == ids_o[b]).all().item() for b in range(B)]
    first_div = []
    for b in range(B):
        neq = (r["ids"][b].long() != ids_o[b]).nonzero()
        first_div.append(int(neq[0]) if len(neq) else -1)
    po = o.softmax(-1).max(-1).values
    log(name=name, B=B, W=img.shape[3], steps_gpu=r["steps"], steps_cpu=aux["ar_steps"], mem_max_abs=dm.max().item(),
        mem_rel_fro=(dm.norm() / mem_o.norm()).item(),
        logits_max_abs=(dl.max().item() if dl is not None else None),
        logits_mean_abs=(dl.mean().item() if dl is not None else None), logit_std=o.std().item(),
        rows_identical=int(sum(same_rows)), first_div=first_div,
        margin_at_div=[(margin[b, d].item() if d >= 0 else None) for b, d in enumerate(first_div)],
        prob_max_abs=(r["probs"] - po).abs().max().item(), rep_gpu=r["rep"].tolist(), rep_cpu=aux["rep_cut"],
        gpu_wall_s=r["wall"], cpu_s=cpu_s, gflop=L.ytk_parseq_last_flops(h) / 1e9)
    L.ytk_parseq_destroy(h)
    return r, o


def main():
    g = torch.Generator().manual_seed(5)
    spec = ops.SPECS["parseq-tiny-dynw-v4"]
    sd = weights.make_parseq_state_dict(spec, seed=3, peaked=False)
    img = torch.rand(16, 3, 32, 320, generator=g) * 2 - 1
    compare("tiny random 16x320", spec, sd, img)
    sdp = weights.make_parseq_state_dict(spec, seed=3, peaked=True)
    compare("tiny peaked 16x320", spec, sdp, img)
    compare("tiny peaked 5x104", spec, sdp, torch.rand(5, 3, 32, 104, generator=g) * 2 - 1)
    sdr = weights.make_parseq_state_dict(spec, seed=3, peaked=True, degenerate_repeat=True)
    compare("tiny degenerate-repeat 4x200", spec, sdr, torch.rand(4, 3, 32, 200, generator=g) * 2 - 1)
    spec = ops.SPECS["parseq-large-v4_1"]
    sdl = weights.make_parseq_state_dict(spec, seed=4, peaked=True)
    compare("large peaked 8x160", spec, sdl, torch.rand(8, 3, 32, 160, generator=g) * 2 - 1)
    compare("large peaked 3x800", spec, sdl, torch.rand(3, 3, 32, 800, generator=g) * 2 - 1)
    # ---- timing: large model, 512 crops of W=184 (median 120 px content + 64 margin), device-side only
    h = make(spec, sdl)
    B, W = 512, 184
    img = (torch.rand(B, 3, 32, W, generator=g) * 2 - 1).cuda()
    S = 101
    ids = torch.zeros(B, S, dtype=torch.int32)
    probs = torch.zeros(B, S)
    steps = ctypes.c_int(0)
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        st = L.ytk_parseq_forward_f32(h, img.data_ptr(), 1, B, W, None, 0, ids.data_ptr(), probs.data_ptr(),
                                      ctypes.byref(steps), None, None, None)
        torch.cuda.synchronize()
        dt = time.time() - t0
        if st != 0:
            raise RuntimeError(L.ytk_last_error().decode())
        fl = L.ytk_parseq_last_flops(h)
        log(name="time large 512x184 iter%d" % it, s=dt, crops_per_s=B / dt, steps=steps.value, gflop=fl / 1e9,
            tflops=fl / dt / 1e12)
    L.ytk_parseq_destroy(h)


if __name__ == "__main__":
    try:
        main()
    except Exception as e:
        import traceback
        traceback.print_exc()
        log(fatal=repr(e))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe_parseq.json", "w"), indent=1)
