"""Benchmark of the DBNet -> PARSeq OCR hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W               # this repo's CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                        # the reference's CPU implementation (oracle restatement)

One step = the hot path over one batch of synthetic 1200x1600 (H x W) pages per GPU (~200 text lines each):
DBNet (`dbnetv2_1`) on every page + PARSeq (`parseq-large-v4_1`, dynamic_width + batch_bucketing) on every crop.
  value : pages/s with pages and packed crops already resident in HBM (device time only, CUDA events)
  e2e   : pages/s through the public batched API (`BatchedOCR`) from HOST pages: H2D, DBNet, D2H of the
          probability maps, host post-processing + crop extraction (process pool, as the reference does on the host),
          H2D crops, PARSeq, D2H ids/probs, tokenizer decode.  Random detector weights do not produce text boxes, so
          the host post-processor consumes a synthetic probability map of the page's ground-truth boxes (the detector
          still runs and its output still crosses PCIe), per SURVEY.md section 8d.
Weights are seeded random (no checkpoints offline); with random PARSeq weights no row emits EOS, so every AR loop
runs all 101 steps (worst case).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pages/sec (DBNet->PARSeq OCR, synthetic 1600x1200 pages, ~200 crops/page)"
PAGES_PER_GPU = 16
REC_MODEL = "parseq-large-v4_1"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d.get("bf16_tflops", 1590.0), "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [t.strip() for t in o.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_pipeline_rate(n_crops_sample, seconds_budget, det_sd=None, rec_sd=None):
    """Times the reference's CPU implementation of the path (oracle restatement, fp32 eager, all host threads) on a
    bounded sample: one page through DBNet + post-processing and `n_crops_sample` crops through PARSeq with the
    reference's batching; extrapolates the recognizer linearly to the page's crop count.  Returns a dict."""
    from oracle import dbnet as odb
    from oracle import parseq as ops
    from oracle import pipeline as opipe
    from yomitoku_b200.models import _dbnet_random_state_dict, _parseq_random_state_dict
    from yomitoku_b200.config import TextRecognizerPARSeqLargeV41Config, load_config
    from yomitoku_b200.synth import synthetic_page, synthetic_prob_map
    # fp32 eager on many small matrices stops scaling (and collapses when oversubscribed) well before 128 threads:
    # use up to 32 intra-op threads, the best setting measured on the 64-core box
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    page, quads = synthetic_page(0)
    if det_sd is None:
        det_sd = _dbnet_random_state_dict(0)
    cfg = load_config(TextRecognizerPARSeqLargeV41Config)
    if rec_sd is None:
        rec_sd = _parseq_random_state_dict(cfg, 0)
    spec = ops.SPECS[REC_MODEL]
    charset = open(cfg.charset, encoding="utf-8").read()
    tok = ops.Tokenizer(charset)
    t0 = time.perf_counter()
    x = opipe.detector_preprocess(page)
    odb.dbnet_forward(det_sd, x)
    t_det = time.perf_counter() - t0
    prob = synthetic_prob_map(quads, (1184, 1600), (1200, 1600))
    t0 = time.perf_counter()
    dq, _ = opipe.dbnet_postprocess(prob, (1200, 1600))
    t_post = time.perf_counter() - t0
    sample = quads[:n_crops_sample]
    t0 = time.perf_counter()
    opipe.recognize(rec_sd, spec, tok, page, sample, dynamic_width=True, batch_bucketing=True, batch_size=128)
    t_rec = time.perf_counter() - t0
    per_page = t_det + t_post + t_rec * (len(quads) / max(1, len(sample)))
    return {"pages_per_s": 1.0 / per_page, "t_det_s": t_det, "t_post_s": t_post, "t_rec_sample_s": t_rec,
            "crops_sample": len(sample), "crops_per_page": len(quads), "cores": torch.get_num_threads()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    detail = None
    t_start = time.perf_counter()
    warm, steps, i = args.warmup, args.steps, 0
    while i < warm + steps:
        r = cpu_pipeline_rate(16, 30.0)
        if i >= warm:
            vals.append(r["pages_per_s"])
            detail = r
        i += 1
        per = (time.perf_counter() - t_start) / i
        # keep the whole run within a few minutes (a step is a bounded sample, but K and W come from the driver)
        if per * (warm + steps) > 240.0:
            warm = min(warm, 1)
            steps = max(1, int(240.0 / per) - warm)
    if not vals:
        vals, detail = [r["pages_per_s"]], r
    v = float(np.mean(vals))
    sample = ("1 page DBNet fp32 + post-processing, %d of %d crops through PARSeq %s (reference batching), "
              "recognizer time scaled to the page's crop count") % (detail["crops_sample"], detail["crops_per_page"],
                                                                    REC_MODEL)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "pages/s", "n_gpus": args.gpus, "steps": len(vals),
        "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "OCR DBNet(dbnetv2_1)->PARSeq(%s), 1200x1600 synthetic pages" % REC_MODEL,
                   "note": "reference CPU path = oracle restatement of yomitoku's PyTorch fp32 eager modules"},
        "cpu_baseline": {"value": v, "unit": "pages/s", "cores": detail["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pages", type=int, default=PAGES_PER_GPU, help="pages per GPU per step")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-window", action="store_true", help="skip the instrumented per-launch GEMM timing step "
                                                               "(for runs under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from yomitoku_b200 import TextDetector, TextRecognizer, _lib
    from yomitoku_b200.parallel import broadcast_state_dict
    from yomitoku_b200.pipeline import BatchedOCR
    from yomitoku_b200.synth import synthetic_page, synthetic_prob_map

    det = TextDetector(from_pretrained=False, device="cuda")
    rec = TextRecognizer(model_name=REC_MODEL, from_pretrained=False, device="cuda", dynamic_width=True,
                         batch_bucketing=True)
    if world > 1:
        # one-time weight broadcast from rank 0 over NCCL (all ranks then hold identical weights)
        det.model.load_state_dict(broadcast_state_dict(det.model.state_dict(), "cuda"))
        rec.model.load_state_dict(broadcast_state_dict(rec.model.state_dict(), "cuda"))
    L = _lib.lib()
    P = args.pages
    pages, quads = [], []
    for i in range(P):
        pg, q = synthetic_page(rank * P + i)
        pages.append(pg)
        quads.append(q)
    Hn, Wn = det.model.input_size(1200, 1600)
    probs_syn = [synthetic_prob_map(q, (Hn, Wn), (1200, 1600)) for q in quads]
    ncpu = os.cpu_count() or 2
    ocr = BatchedOCR(det, rec, det_batch=8, workers=max(2, min(32, (ncpu - 2 * world) // world)))
    # ---------------- device-resident inputs for `value`
    pages_dev = torch.from_numpy(np.stack(pages)).cuda()
    prob_dev = torch.empty((P, Hn, Wn), dtype=torch.float32, device="cuda")
    from yomitoku_b200.data import ParseqDataset
    per_page = []
    for pg, q in zip(pages, quads):
        ds = ParseqDataset(rec._cfg, pg, q, dynamic_width=True)
        per_page.append((ds.data, ds.content_widths, len(q)))
    # pack all crops exactly as BatchedOCR.recognize_pooled does (reference grouping per page)
    from yomitoku_b200.text_recognizer import plan_mini_batches
    flat_c, flat_p, flat_g = [], [], []
    g0 = 0
    for canv, cw, nq in per_page:
        order = np.argsort(cw).tolist()
        plan = plan_mini_batches([c.shape[1] for c in canv], order, True, rec._cfg.data.batch_size, None, None)
        padded, group = rec._collate_widths(canv, plan)
        for b in plan:
            for i in b:
                flat_c.append(canv[i])
                flat_p.append(padded[i])
                flat_g.append(g0 + group[i])
        g0 += len(plan)
    n_crops = len(flat_c)
    buf, total, descs, n_tok = rec.model.pack_crops(flat_c, flat_p, flat_g)
    buf_dev = buf.cuda()
    def det_step():
        for s in range(0, P, ocr.det_batch):
            e = min(P, s + ocr.det_batch)
            _lib.check(L.ytk_dbnet_forward_u8(det.model._ensure(), pages_dev[s:e].data_ptr(), 1, e - s, 1200, 1600,
                                              prob_dev[s:e].data_ptr(), 1, None))

    sel_geoms = None
    if ocr.device_crops:
        # device-side crop extraction: the canvases of every step are cut on the GPU from the resident pages (same flat
        # order as the packed host canvases above), so `value` covers detector + crop kernels + recognizer
        from yomitoku_b200.data import crop_geometry
        from yomitoku_b200.models import extract_crops_device
        flat_geoms = []
        for pi, ((canv, cw, nq), q) in enumerate(zip(per_page, quads)):
            g, keep = crop_geometry((1200, 1600), q, rec._cfg.data.img_size, True, page=pi)
            order = np.argsort(cw).tolist()
            plan = plan_mini_batches([c.shape[1] for c in canv], order, True, rec._cfg.data.batch_size, None, None)
            flat_geoms.append(g[np.asarray([i for b in plan for i in b], np.int64)])
        sel_geoms = np.concatenate(flat_geoms)
        chk, chk_total = extract_crops_device(pages_dev, sel_geoms)
        if chk_total != total or not torch.equal(chk[:total], buf_dev[:total]):
            raise RuntimeError("device-cut canvases differ from the OpenCV canvases")
        del chk

    def rec_step():
        if sel_geoms is not None:
            canv_dev, _ = extract_crops_device(pages_dev, sel_geoms)
            return rec.model.run_packed(canv_dev, total, descs, n_crops, g0)
        return rec.model.run_packed(buf_dev, total, descs, n_crops, g0)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        det_step()
        rec_step()
    sync_all()
    launches0 = L.ytk_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    det_ms = rec_ms = 0.0
    with ClockSampler(local) as clocks:
        t_all0 = torch.cuda.Event(enable_timing=True)
        t_all1 = torch.cuda.Event(enable_timing=True)
        t_all0.record()
        for _ in range(args.steps):
            ev[0].record()
            det_step()
            ev[1].record()
            rec_step()
            ev[2].record()
            torch.cuda.synchronize()
            det_ms += ev[0].elapsed_time(ev[1])
            rec_ms += ev[1].elapsed_time(ev[2])
        t_all1.record()
        sync_all()
        total_ms = t_all0.elapsed_time(t_all1)
    launches = L.ytk_launch_count() - launches0
    tm = torch.tensor([total_ms, det_ms, rec_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms, det_ms, rec_ms = [float(v) for v in tm.tolist()]
    value = world * P * args.steps / (total_ms / 1e3)
    det_flops = det.model.flops(ocr.det_batch, Hn, Wn) / ocr.det_batch * P
    rec_flops = rec.model.last_flops()
    phase_value = rec.model.last_phase_ms()     # CUDA-event phase times of the last timed recognizer call
    pk = peaks()
    det_tflops = det_flops * args.steps / (det_ms / 1e3) / 1e12
    rec_tflops = rec_flops * args.steps / (rec_ms / 1e3) / 1e12
    # ---------------- per-launch timing of the dominant kernel (one instrumented extra step, outside the timed region)
    def gemm_window(fn):
        f, ms, n = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_longlong(0)
        L.ytk_gemm_profile_begin()
        fn()
        torch.cuda.synchronize()
        _lib.check(L.ytk_gemm_profile_end(ctypes.byref(f), ctypes.byref(ms), ctypes.byref(n)))
        return {"tflop": f.value / 1e12, "ms": ms.value, "launches": int(n.value),
                "achieved": f.value / 1e12 / (ms.value / 1e3) if ms.value > 0 else 0.0}
    if args.no_window:
        g_det = g_rec = {"tflop": 0.0, "ms": 1e-9, "launches": 0, "achieved": 0.0}
    else:
        g_det = gemm_window(det_step)
        g_rec = gemm_window(rec_step)
    # ---------------- e2e through the public batched API from host pages
    e2e = None
    if not args.no_e2e:
        # warm-up through the same entry point: touches every slot of the staging ring, starts the worker pool
        for _ in ocr.stream([pages] * max(3, args.warmup), lookahead=2, prob_override=[probs_syn] * max(3, args.warmup)):
            pass
        sync_all()
        t0 = time.perf_counter()
        n_words = 0
        # documented pipelined use of the public API: `BatchedOCR.stream` runs the detector + host stage of the next
        # batches (own thread, own CUDA stream, process pool) while the recognizer works on the current one; exactly
        # `steps` batches of P pages go through
        for res in ocr.stream([pages] * args.steps, lookahead=2, prob_override=[probs_syn] * args.steps):
            n_words += sum(len(r.words) for r in res)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        # device_crops: the crops never cross PCIe, only their 136-byte records do
        h2d = P * 1200 * 1600 * 3 + (n_crops * 136 if ocr.device_crops else total)
        d2h = P * Hn * Wn * 4 + n_crops * 101 * 8
        e2e = {"value": world * P * args.steps / dt, "unit": "pages/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "words_per_page": n_words / (args.steps * P),
               "host_workers": ocr.workers, "device_crops": bool(ocr.device_crops)}
    ocr.close()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_pipeline_rate(16, 30.0, det.model.state_dict(), rec.model.state_dict())
        cpu = {"value": r["pages_per_s"], "unit": "pages/s", "cores": r["cores"], "kind": "port",
               "sample": "1 page DBNet fp32 (%.2f s) + post-processing (%.3f s) + %d of %d crops PARSeq %s (%.2f s), "
                         "recognizer scaled to the page's crop count; oracle = restated reference fp32 eager path"
                         % (r["t_det_s"], r["t_post_s"], r["crops_sample"], r["crops_per_page"], REC_MODEL,
                            r["t_rec_sample_s"])}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "pages/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "OCR DBNet(dbnetv2_1)->PARSeq(%s), %d synthetic 1200x1600 pages/GPU/step, %d crops"
                                   % (REC_MODEL, P, n_crops),
                       "parallelism": "pages sharded %d/GPU; value: crops recognised on the owning GPU; e2e: mini-batch groups "
                                      "balanced across GPUs (all_to_all crop scatter / result gather)" % P,
                       "l2": "working set (%.1f GB activations per step) >> 126 MB L2; no explicit flush" %
                             (P * 1.2 + 4.0),
                       "ar_steps": int(L.ytk_parseq_last_steps(rec.model._ensure())),
                       "weights": "seeded random init (from_pretrained=False)",
                       "crops": "cut on the GPU from the resident pages (ytk_extract_crops_u8, checked equal to the "
                                "OpenCV canvases)" if sel_geoms is not None else "cut on the host (OpenCV)",
                       "recognizer_phase_ms": phase_value},
            "crops_per_s": world * n_crops * args.steps / (rec_ms / 1e3),
            "det_pages_per_s": world * P * args.steps / (det_ms / 1e3),
            "roofline": {"bound": "tensor",
                         "achieved": (g_det["tflop"] + g_rec["tflop"]) / ((g_det["ms"] + g_rec["ms"]) / 1e3),
                         "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": (g_det["tflop"] + g_rec["tflop"]) / ((g_det["ms"] + g_rec["ms"]) / 1e3) /
                                 pk["bf16_tflops_sustained"],
                         "traffic": None,
                         "kernel": "gemm_tc_kernel (tcgen05 implicit GEMM): every launch of one step (DBNet convs + "
                                   "PARSeq linears), algorithmic FLOPs (2*M*N*K per launch) over the sum of the launch "
                                   "durations; CUDA events around every launch on the launching stream, one "
                                   "instrumented step right after the timed region",
                         "launches_per_step": g_det["launches"] + g_rec["launches"],
                         "kernel_ms_per_step": g_det["ms"] + g_rec["ms"],
                         "share_of_step": (g_det["ms"] + g_rec["ms"]) / (total_ms / args.steps),
                         "by_model": {"dbnet": g_det, "parseq": g_rec},
                         "traffic_note": "shapes differ per launch; ncu --set full DRAM bytes of representative "
                                         "launches are in profiles/README_r01.md (qkv GEMM: 76 MB read + 165 MB "
                                         "written vs 76 + 217 MB algorithmic)",
                         "peak_source": pk["source"] + " bf16_tflops_sustained",
                         "whole_sequence": {"detector": {"achieved": det_tflops,
                                                         "frac": det_tflops / pk["bf16_tflops_sustained"],
                                                         "gflop_per_page": det_flops / P / 1e9},
                                            "recognizer": {"achieved": rec_tflops,
                                                           "frac": rec_tflops / pk["bf16_tflops_sustained"],
                                                           "gflop_per_step": rec_flops / 1e9}}},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
