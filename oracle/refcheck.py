"""Pins the oracle against the reference's OWN model code (TEST INFRASTRUCTURE; runs only where /root/reference
exists, i.e. in the build container - never on the GPU box).

`import yomitoku` is impossible offline (timm, omegaconf, pyclipper, shapely, onnx*, pypdfium2 are not installed), so
the reference's model files are loaded *by path*, unmodified, with stand-ins for the missing third-party imports:
  * timm.models.vision_transformer.{VisionTransformer, PatchEmbed}, timm.models.helpers.named_apply - a minimal
    restatement of timm 1.0.27's ViT (same parameter names) so that `Encoder(VisionTransformer)` and its
    `forward_features_dynamic` (reference code) run;
  * the `cfg` objects (OmegaConf in the reference) are attribute-dicts built from the reference's own config
    dataclasses' values.
  * for the host rows: `pyclipper` / `shapely` -> this oracle's restatements (build_reference_postprocessor); the
    modules the reference's text_recognizer.py / text_detector.py import but this path never executes (omegaconf-based
    `base`, configs, models, visualizer, onnx*, schemas) -> empty stubs, so that the reference's own `__call__`,
    `preprocess`, `_make_mini_batch`, `_collate`, `postprocess`, `_apply_orientation_fallback` and `ParseqDataset`
    run unmodified around stand-in models (build_reference_recognizer_shell / build_reference_detector_shell).
Nothing is copied into the repo: modules are executed from /root/reference in-process.

Usage:  python -m oracle.refcheck            (prints PASS/FAIL lines; exit code 0 iff all pass)
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = os.environ.get("YTK_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src", "yomitoku")


def available():
    return os.path.isdir(SRC)


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


# ---------------------------------------------------------------------------------- timm stand-in
def _install_timm_standin():
    if "timm" in sys.modules:
        return

    class PatchEmbed(nn.Module):
        def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, **kw):
            super().__init__()
            img_size = tuple(img_size) if isinstance(img_size, (list, tuple)) else (img_size, img_size)
            patch_size = tuple(patch_size) if isinstance(patch_size, (list, tuple)) else (patch_size, patch_size)
            self.img_size, self.patch_size = img_size, patch_size
            self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
            self.num_patches = self.grid_size[0] * self.grid_size[1]
            self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
            self.norm = nn.Identity()

        def forward(self, x):
            return self.norm(self.proj(x).flatten(2).transpose(1, 2))

    class Attention(nn.Module):
        def __init__(self, dim, num_heads, qkv_bias):
            super().__init__()
            self.num_heads, self.head_dim = num_heads, dim // num_heads
            self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
            self.proj = nn.Linear(dim, dim)

        def forward(self, x):
            B, N, C = x.shape
            qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
            x = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
            return self.proj(x.transpose(1, 2).reshape(B, N, C))

    class Mlp(nn.Module):
        def __init__(self, dim, hidden):
            super().__init__()
            self.fc1, self.act, self.fc2 = nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim)

        def forward(self, x):
            return self.fc2(self.act(self.fc1(x)))

    class Block(nn.Module):
        def __init__(self, dim, num_heads, mlp_ratio, qkv_bias):
            super().__init__()
            self.norm1 = nn.LayerNorm(dim, eps=1e-6)
            self.attn = Attention(dim, num_heads, qkv_bias)
            self.norm2 = nn.LayerNorm(dim, eps=1e-6)
            self.mlp = Mlp(dim, int(dim * mlp_ratio))

        def forward(self, x):
            x = x + self.attn(self.norm1(x))
            return x + self.mlp(self.norm2(x))

    class VisionTransformer(nn.Module):
        def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, global_pool="token",
                     embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0, qkv_bias=True, class_token=True,
                     drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, embed_layer=PatchEmbed, **kw):
            super().__init__()
            assert not class_token and num_classes == 0 and global_pool == ""
            self.patch_embed = embed_layer(img_size=img_size, patch_size=patch_size, in_chans=in_chans,
                                           embed_dim=embed_dim)
            self.pos_embed = nn.Parameter(torch.randn(1, self.patch_embed.num_patches, embed_dim) * 0.02)
            self.pos_drop = nn.Identity()
            self.patch_drop = nn.Identity()
            self.norm_pre = nn.Identity()
            self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio, qkv_bias) for _ in range(depth)])
            self.norm = nn.LayerNorm(embed_dim, eps=1e-6)

        def no_weight_decay(self):
            return {"pos_embed"}

        def forward_features(self, x):
            x = self.patch_embed(x) + self.pos_embed
            return self.norm(self.blocks(self.norm_pre(x)))

    def named_apply(fn, module, name="", depth_first=True, include_root=False):
        if not depth_first and include_root:
            fn(module=module, name=name)
        for child_name, child in module.named_children():
            child_name = ".".join((name, child_name)) if name else child_name
            named_apply(fn=fn, module=child, name=child_name, depth_first=depth_first, include_root=True)
        if depth_first and include_root:
            fn(module=module, name=name)
        return module

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    vit = types.ModuleType("timm.models.vision_transformer")
    helpers = types.ModuleType("timm.models.helpers")
    vit.VisionTransformer, vit.PatchEmbed = VisionTransformer, PatchEmbed
    helpers.named_apply = named_apply
    timm.models, models.vision_transformer, models.helpers = models, vit, helpers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.vision_transformer": vit,
                        "timm.models.helpers": helpers})


def _load(modname, relpath, package=None):
    path = os.path.join(SRC, relpath)
    spec = importlib.util.spec_from_file_location(modname, path, submodule_search_locations=None)
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def _pkg(name):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    return sys.modules[name]


def load_reference_models():
    """Returns (DBNet class, PARSeq class, ParseqTokenizer class) executed from /root/reference."""
    _install_timm_standin()
    _pkg("ytk_ref")
    _pkg("ytk_ref.models")
    _pkg("ytk_ref.models.layers")
    _pkg("ytk_ref.postprocessor")
    _load("ytk_ref.models.layers.dbnet_feature_attention", "models/layers/dbnet_feature_attention.py",
          "ytk_ref.models.layers")
    _load("ytk_ref.models.layers.parseq_transformer", "models/layers/parseq_transformer.py", "ytk_ref.models.layers")
    db = _load("ytk_ref.models.dbnet_plus", "models/dbnet_plus.py", "ytk_ref.models")
    ps = _load("ytk_ref.models.parseq", "models/parseq.py", "ytk_ref.models")
    tk = _load("ytk_ref.postprocessor.parseq_tokenizer", "postprocessor/parseq_tokenizer.py", "ytk_ref.postprocessor")
    return db.DBNet, ps.PARSeq, tk.ParseqTokenizer


def reference_dbnet_cfg():
    # values of reference configs/cfg_text_detector_dbnet_v2_1.py:5-21
    return AttrDict(backbone=AttrDict(name="resnet50", dilation=True),
                    decoder=AttrDict(in_channels=[256, 512, 1024, 2048], hidden_dim=256, adaptive=True, serial=True,
                                     smooth=False, k=50))


def reference_parseq_cfg(spec):
    return AttrDict(
        max_label_length=spec.max_label_length, decode_ar=spec.decode_ar, refine_iters=spec.refine_iters,
        num_tokens=spec.num_tokens,
        data=AttrDict(img_size=list(spec.img_size)),
        encoder=AttrDict(patch_size=list(spec.patch), num_heads=spec.enc_heads, embed_dim=spec.embed_dim,
                         mlp_ratio=spec.mlp_ratio, depth=spec.enc_depth),
        decoder=AttrDict(embed_dim=spec.embed_dim, num_heads=spec.dec_heads, mlp_ratio=spec.dec_mlp_ratio, depth=1))


def build_reference_dbnet(sd):
    DBNet, _, _ = load_reference_models()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m = DBNet(reference_dbnet_cfg())
    m.load_state_dict(sd, strict=True)
    return m.eval()


def build_reference_parseq(spec, sd, charset):
    _, PARSeq, Tok = load_reference_models()
    m = PARSeq(reference_parseq_cfg(spec))
    m.load_state_dict(sd, strict=True)
    m.tokenizer = Tok(charset)
    return m.eval()


# ---------------------------------------------------------------------------------- RT-DETRv2 (layout / table models)
def load_reference_rtdetr():
    """(RTDETRv2 class, RTDETRPostProcessor class) executed from /root/reference.  `omegaconf` (not installable here) is
    a two-line stand-in: the decoder only does `isinstance(num_points, ListConfig)` (rtdetrv2_decoder.py:26, 78)."""
    if "omegaconf" not in sys.modules:
        om = types.ModuleType("omegaconf")

        class ListConfig(list):
            pass

        om.ListConfig = ListConfig
        sys.modules["omegaconf"] = om
    _pkg("ytk_ref")
    _pkg("ytk_ref.models")
    _pkg("ytk_ref.models.layers")
    _pkg("ytk_ref.postprocessor")
    for n in ("activate", "rtdetr_backbone", "rtdetr_hybrid_encoder", "rtdetrv2_decoder"):
        if "ytk_ref.models.layers." + n not in sys.modules:
            _load("ytk_ref.models.layers." + n, "models/layers/%s.py" % n, "ytk_ref.models.layers")
    m = sys.modules.get("ytk_ref.models.rtdetr") or _load("ytk_ref.models.rtdetr", "models/rtdetr.py", "ytk_ref.models")
    pp = sys.modules.get("ytk_ref.postprocessor.rtdetr_postprocessor") or _load(
        "ytk_ref.postprocessor.rtdetr_postprocessor", "postprocessor/rtdetr_postprocessor.py", "ytk_ref.postprocessor")
    return m.RTDETRv2, pp.RTDETRPostProcessor


def reference_rtdetr_cfg(num_classes):
    # values of reference configs/cfg_layout_parser_rtdtrv2_v2.py:10-61 / cfg_table_structure_recognizer_rtdtrv2.py
    LC = sys.modules["omegaconf"].ListConfig
    return AttrDict(
        PResNet=dict(depth=50, variant="d", freeze_at=0, return_idx=[1, 2, 3], num_stages=4, freeze_norm=True),
        HybridEncoder=dict(in_channels=[512, 1024, 2048], feat_strides=[8, 16, 32], hidden_dim=256, use_encoder_idx=[2],
                           num_encoder_layers=1, nhead=8, dim_feedforward=1024, dropout=0.0, enc_act="gelu",
                           expansion=1.0, depth_mult=1, act="silu"),
        RTDETRTransformerv2=dict(num_classes=num_classes, feat_channels=[256, 256, 256], feat_strides=[8, 16, 32],
                                 hidden_dim=256, num_levels=3, num_layers=6, num_queries=300, num_denoising=100,
                                 label_noise_ratio=0.5, box_noise_scale=1.0, eval_spatial_size=[640, 640], eval_idx=-1,
                                 num_points=LC([4, 4, 4]), cross_attn_method="default", query_select_method="default"))


def build_reference_rtdetr(num_classes, sd=None):
    RTDETRv2, _ = load_reference_rtdetr()
    m = RTDETRv2(reference_rtdetr_cfg(num_classes))
    if sd is not None:
        m.load_state_dict(sd, strict=True)
    return m.eval()


def build_reference_postprocessor(**kwargs):
    """The reference's own DBnetPostProcessor (postprocessor/dbnet_postporcessor.py, executed from /root/reference) with
    the two third-party imports that are not installable offline replaced by this oracle's restatements:
    `pyclipper.PyclipperOffset` (AddPath(JT_ROUND, ET_CLOSEDPOLYGON) + Execute) -> oracle.pipeline.clipper_offset_box,
    `shapely.geometry.Polygon(box).area/.length` -> shoelace / perimeter in float64.  Everything else - contour order,
    minAreaRect corner ordering, box_score_fast, the size / score filters, scaling, rounding, int16 - is the
    reference's code, so this pins the oracle's control flow of row R3 against it."""
    from oracle import pipeline as opipe

    pc = types.ModuleType("pyclipper")
    pc.JT_ROUND, pc.ET_CLOSEDPOLYGON = 1, 0

    class PyclipperOffset:
        def __init__(self):
            self.path = None

        def AddPath(self, path, join_type, end_type):
            assert join_type == pc.JT_ROUND and end_type == pc.ET_CLOSEDPOLYGON
            self.path = np.asarray(path)

        def Execute(self, delta):
            return [opipe.clipper_offset_box(self.path, float(delta)).tolist()]

    pc.PyclipperOffset = PyclipperOffset
    sh, shg = types.ModuleType("shapely"), types.ModuleType("shapely.geometry")

    class Polygon:
        def __init__(self, pts):
            self.p = np.asarray(pts, dtype=np.float64)

        @property
        def area(self):
            p = self.p
            return 0.5 * abs(np.dot(p[:, 0], np.roll(p[:, 1], -1)) - np.dot(p[:, 1], np.roll(p[:, 0], -1)))

        @property
        def length(self):
            p = self.p
            return np.sqrt(((p - np.roll(p, -1, axis=0)) ** 2).sum(1)).sum()

    shg.Polygon = Polygon
    sh.geometry = shg
    saved = {k: sys.modules.get(k) for k in ("pyclipper", "shapely", "shapely.geometry")}
    sys.modules.update({"pyclipper": pc, "shapely": sh, "shapely.geometry": shg})
    try:
        _pkg("ytk_ref")
        _pkg("ytk_ref.postprocessor")
        mod = _load("ytk_ref.postprocessor.dbnet_postporcessor", "postprocessor/dbnet_postporcessor.py",
                    "ytk_ref.postprocessor")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod.DBnetPostProcessor(**kwargs)


def reference_postprocess(post, prob, ori_hw):
    """Runs the reference post-processor on one (H, W) float32 probability map like TextDetector.__call__ does
    (text_detector.py:129-131: preds["binary"] is a (1, 1, H, W) tensor)."""
    return post({"binary": torch.from_numpy(np.ascontiguousarray(prob))[None, None]}, ori_hw)


# ------------------------------------------------------------------- the recognizer's host flow (rows R4, R5, R10, R11)
def flow_hash(u8_chw):
    """Position-sensitive checksum of a crop tensor as uint8 (3, 32, W); black padding contributes nothing, so it does
    not depend on the padded width.  Shared by the stand-in models of the reference flow and of the product flow."""
    c, y, x = np.meshgrid(np.arange(u8_chw.shape[0]), np.arange(u8_chw.shape[1]), np.arange(u8_chw.shape[2]),
                          indexing="ij")
    return int((u8_chw.astype(np.int64) * (1 + (c * 7 + y * 13 + x * 3) % 251)).sum())


def flow_token(h, padded_w, batch_len, n_classes):
    """(token id, probability) the stand-in model emits at position 0 for a crop with checksum h that sits in a
    mini-batch of `batch_len` crops padded to `padded_w`; position 1 is EOS with probability 1."""
    return 1 + (h * 31 + padded_w * 7 + batch_len * 3) % (n_classes - 1), 0.55 + 0.44 * ((h % 1000) / 1000.0)


class _FlowModelOutput:
    def __init__(self, probs):
        self.probs = probs

    def softmax(self, dim):          # the reference calls self.model(data).softmax(-1) (text_recognizer.py:255)
        return self.probs


def flow_model(n_classes, S):
    """Stand-in for PARSeq in the reference flow: data (B, 3, 32, W) normalised float -> object whose .softmax(-1) is a
    (B, S, C) distribution built from flow_token of every row."""
    def model(data):
        B, _, _, W = data.shape
        u8 = torch.round((data * 0.5 + 0.5) * 255.0).to(torch.uint8).numpy()
        probs = torch.zeros((B, S, n_classes), dtype=torch.float32)
        probs[:, 1:, 0] = 1.0                                    # EOS from position 1 on
        for b in range(B):
            tok, p = flow_token(flow_hash(u8[b]), W, B, n_classes)
            probs[b, 0, :] = (1.0 - p) / (n_classes - 1)
            probs[b, 0, tok] = p
        return _FlowModelOutput(probs)
    return model


def build_reference_recognizer_shell(charset, img_size=(32, 800), batch_size=10, width_budget=8000, max_batch_size=64,
                                     max_label_length=25, **flags):
    """An instance of the reference's own TextRecognizer class (text_recognizer.py executed from /root/reference)
    without running its __init__ (which needs the HF hub / omegaconf): the attributes __call__ reads are set by hand,
    `model` is the stand-in above, `tokenizer` the reference's ParseqTokenizer.  Every import of the module that is not
    installable offline or not on this path is a stub module (omegaconf-based `base`, configs, models, visualizer,
    onnx*, schemas); data/dataset.py, data/functions.py and postprocessor/parseq_tokenizer.py are the real files."""
    import importlib.machinery
    absent = [n for n in ("pypdfium2", "onnx", "onnxruntime") if n not in sys.modules]
    for name in absent:       # import-time stubs only; removed again below (a spec-less stub left in sys.modules
        m = types.ModuleType(name)          # confuses importlib.util.find_spec users such as torch._dynamo)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = m
    _pkg("ytk_ref")
    _pkg("ytk_ref.data")
    _pkg("ytk_ref.utils")
    _pkg("ytk_ref.postprocessor")
    _load("ytk_ref.constants", "constants.py", "ytk_ref")
    _load("ytk_ref.utils.logger", "utils/logger.py", "ytk_ref.utils")
    _load("ytk_ref.data.functions", "data/functions.py", "ytk_ref.data")
    _load("ytk_ref.data.dataset", "data/dataset.py", "ytk_ref.data")
    tk = _load("ytk_ref.postprocessor.parseq_tokenizer", "postprocessor/parseq_tokenizer.py", "ytk_ref.postprocessor")
    sys.modules["ytk_ref.postprocessor"].ParseqTokenizer = tk.ParseqTokenizer

    def stub(modname, **attrs):
        m = types.ModuleType(modname)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[modname] = m
        return m

    class BaseModelCatalog:
        def __init__(self):
            self.entries = {}

        def register(self, name, cfg, model):
            self.entries[name] = (cfg, model)

    class Schema(dict):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.__dict__.update(kw)

    stub("ytk_ref.base", BaseModelCatalog=BaseModelCatalog, BaseModule=object)
    stub("ytk_ref.configs", **{n: type(n, (), {}) for n in (
        "TextRecognizerPARSeqConfig", "TextRecognizerPARSeqSmallConfig", "TextRecognizerPARSeqV2Config",
        "TextRecognizerPARSeqTinyConfig", "TextRecognizerPARSeqLargeV41Config", "TextRecognizerPARSeqTinyDynwV4Config")})
    stub("ytk_ref.models", PARSeq=None)
    stub("ytk_ref.utils.misc", load_charset=lambda p: open(p, encoding="utf-8").read())
    stub("ytk_ref.utils.visualizer", rec_visualizer=None)
    stub("ytk_ref.schemas", TextRecognizerSchema=Schema)
    try:
        mod = _load("ytk_ref.text_recognizer", "text_recognizer.py", "ytk_ref")
    finally:
        for name in absent:
            sys.modules.pop(name, None)
    rec = object.__new__(mod.TextRecognizer)
    data = AttrDict(img_size=list(img_size), batch_size=batch_size)
    if width_budget:
        data.width_budget, data.max_batch_size = width_budget, max_batch_size
    rec._cfg = AttrDict(data=data, max_label_length=max_label_length)
    rec.tokenizer = tk.ParseqTokenizer(charset)
    rec.model = flow_model(len(rec.tokenizer) - 2, max_label_length + 1)
    rec.device, rec.infer_onnx, rec.visualize, rec.num_parallel_batches = "cpu", False, False, 1
    rec.dynamic_width = flags.get("dynamic_width", False)
    rec.batch_bucketing = flags.get("batch_bucketing", False)
    rec.source_downscale = flags.get("source_downscale", False)
    rec.rec_orientation_fallback = flags.get("rec_orientation_fallback", False)
    rec.rec_orientation_fallback_thresh = flags.get("rec_orientation_fallback_thresh", 0.75)
    return rec


# ------------------------------------------------------------------- the detector's host flow (rows R1, R3)
def flow_detector_model(tensor):
    """Stand-in for DBNet in the reference / product detector flow: a probability map that is a smooth function of the
    normalised input tensor (1, 3, H, W): text strokes (dark) light up, the white background stays near 0."""
    m = F.avg_pool2d(-tensor.mean(1, keepdim=True), 7, 1, 3)
    return {"binary": torch.sigmoid(2.0 * m + 1.0)}


def build_reference_detector_shell(shortest_size=1280, limit_size=1600, **post):
    """An instance of the reference's own TextDetector (text_detector.py executed from /root/reference) without its
    __init__: `preprocess` / `postprocess` / `__call__` are the reference's code, data/functions.py and the
    post-processor are the real files (the latter with the pyclipper / shapely stand-ins), the model is
    flow_detector_model."""
    import importlib.machinery
    post = post or dict(min_size=2, thresh=0.3, box_thresh=0.4, max_candidates=1500, unclip_ratio=3.5)
    post_obj = build_reference_postprocessor(**post)
    absent = [n for n in ("pypdfium2", "onnx", "onnxruntime") if n not in sys.modules]
    for name in absent:
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = m
    _pkg("ytk_ref")
    _pkg("ytk_ref.data")
    _pkg("ytk_ref.utils")
    _load("ytk_ref.constants", "constants.py", "ytk_ref")
    _load("ytk_ref.utils.logger", "utils/logger.py", "ytk_ref.utils")
    _load("ytk_ref.data.functions", "data/functions.py", "ytk_ref.data")

    def stub(modname, **attrs):
        m = types.ModuleType(modname)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[modname] = m

    class BaseModelCatalog:
        def __init__(self):
            self.entries = {}

        def register(self, name, cfg, model):
            self.entries[name] = (cfg, model)

    class Schema(dict):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.__dict__.update(kw)

    stub("ytk_ref.base", BaseModelCatalog=BaseModelCatalog, BaseModule=object)
    stub("ytk_ref.configs", **{n: type(n, (), {}) for n in (
        "TextDetectorDBNetConfig", "TextDetectorDBNetV2Config", "TextDetectorDBNetV2_1Config")})
    stub("ytk_ref.models", DBNet=None)
    sys.modules["ytk_ref.postprocessor"].DBnetPostProcessor = type(post_obj)
    stub("ytk_ref.utils.visualizer", det_visualizer=None)
    stub("ytk_ref.schemas", TextDetectorSchema=Schema)
    try:
        mod = _load("ytk_ref.text_detector", "text_detector.py", "ytk_ref")
    finally:
        for name in absent:
            sys.modules.pop(name, None)
    det = object.__new__(mod.TextDetector)
    det._cfg = AttrDict(data=AttrDict(shortest_size=shortest_size, limit_size=limit_size))
    det.post_processor = post_obj
    det.model = flow_detector_model
    det.device, det.infer_onnx, det.visualize = "cpu", False, False
    return det


def postprocess_cases():
    """Seeded probability maps for the post-processor checks: blurred rectangles (axis-aligned and rotated), touching
    blobs, tiny specks that the size / score filters drop; values quantised to 1/255 so that they can be stored exactly."""
    import cv2
    cases = []
    for seed, (H, W) in ((0, (592, 800)), (1, (400, 640))):
        rng = np.random.default_rng(900 + seed)
        m = np.zeros((H, W), np.float32)
        for k in range(70):
            cx, cy = rng.uniform(30, W - 30), rng.uniform(20, H - 20)
            w, h = rng.uniform(3, 120), rng.uniform(2, 26)
            ang = 0.0 if k % 3 else rng.uniform(-35, 35)
            box = cv2.boxPoints(((cx, cy), (w, h), ang)).astype(np.int32)
            cv2.fillPoly(m, [box], float(rng.uniform(0.35, 1.0)))
        m = cv2.GaussianBlur(m, (5, 5), 0)
        cases.append((np.round(np.clip(m, 0, 1) * 255).astype(np.uint8), (H * 2 + 16, W * 2)))
    return cases


def main():
    from oracle import dbnet as odb
    from oracle import parseq as ops
    from oracle import weights
    ok = True

    def report(name, cond, detail=""):
        nonlocal ok
        ok = ok and bool(cond)
        print("%s  %s  %s" % ("PASS" if cond else "FAIL", name, detail), flush=True)

    torch.manual_seed(0)
    # ---- DBNet
    sd = weights.make_dbnet_state_dict(seed=1)
    ref = build_reference_dbnet(sd)
    x = torch.randn(1, 3, 96, 160)
    with torch.inference_mode():
        r = ref(x)["binary"]
    o = odb.dbnet_forward(sd, x)
    report("dbnet prob map vs reference DBNet.forward", torch.equal(r, o) or (r - o).abs().max() < 1e-6,
           "max|d|=%.3g" % (r - o).abs().max().item())
    # ---- PARSeq
    charset = open(os.path.join(SRC, "resource", "charsetv2.txt"), encoding="utf-8").read()
    import dataclasses
    for name, W, peaked, over in (("parseq-tiny-dynw-v4", 320, False, {}), ("parseq-tiny-dynw-v4", 200, True, {}),
                                  ("parseq-large-v4_1", 160, True, {}),
                                  # the cfg switches of the decoder: non-AR decode, no / repeated refinement
                                  ("parseq-tiny-dynw-v4", 200, True, {"decode_ar": 0}),
                                  ("parseq-tiny-dynw-v4", 200, True, {"decode_ar": 0, "refine_iters": 0}),
                                  ("parseq-tiny-dynw-v4", 200, True, {"refine_iters": 2}),
                                  ("parseq-tiny-dynw-v4", 200, True, {"refine_iters": 0}),
                                  ("parseq-tiny", 208, True, {})):
        spec = dataclasses.replace(ops.SPECS[name], **over)
        sd = weights.make_parseq_state_dict(spec, seed=3, peaked=peaked)
        ref = build_reference_parseq(spec, sd, charset)
        img = torch.rand(4, 3, 32, W, generator=torch.Generator().manual_seed(5)) * 2 - 1
        with torch.inference_mode():
            r = ref(img)
        o, aux = ops.parseq_forward(sd, spec, img, return_aux=True)
        same_shape = r.shape == o.shape
        d = (r - o).abs().max().item() if same_shape else float("nan")
        ids_same = same_shape and torch.equal(r.argmax(-1), o.argmax(-1))
        report("parseq %s W=%d peaked=%d %s logits vs reference PARSeq.forward" % (name, W, peaked, over or ""),
               same_shape and d < 2e-4 and ids_same, "shape=%s max|d|=%.3g ar_steps=%d" % (tuple(o.shape), d,
                                                                                           aux["ar_steps"]))
        tok_r = ref.tokenizer.decode(r.softmax(-1))
        tok_o = ops.Tokenizer(charset).decode(o.softmax(-1))
        report("tokenizer decode strings+scores", tok_r[0] == tok_o[0] and
               all(abs(a - b) <= 2e-3 * max(abs(a), 1e-30) for a, b in zip(tok_r[1], tok_o[1])))
    # ---- DBNet post-processing (row R3): the reference's own file against the oracle, three threshold sets
    from oracle import pipeline as opipe
    for name, kw in (("dbnetv2_1", dict(min_size=2, thresh=0.3, box_thresh=0.4, max_candidates=1500, unclip_ratio=3.5)),
                     ("dbnetv2", dict(min_size=2, thresh=0.2, box_thresh=0.5, max_candidates=1500, unclip_ratio=5.0)),
                     ("few", dict(min_size=2, thresh=0.3, box_thresh=0.4, max_candidates=20, unclip_ratio=3.5))):
        post = build_reference_postprocessor(**kw)
        for ci, (pu8, ori) in enumerate(postprocess_cases()):
            prob = pu8.astype(np.float32) / 255.0
            rq, rs = reference_postprocess(post, prob, ori)
            oq, os_ = opipe.dbnet_postprocess(prob, ori, thresh=kw["thresh"], box_thresh=kw["box_thresh"],
                                              max_candidates=kw["max_candidates"], unclip_ratio=kw["unclip_ratio"],
                                              min_size=kw["min_size"])
            report("dbnet post-processing %s case %d vs reference DBnetPostProcessor (clipper / shapely stand-ins)"
                   % (name, ci), rq == oq and len(rs) == len(os_) and all(a == b for a, b in zip(rs, os_)),
                   "boxes=%d" % len(rq))
    return 0 if ok else 1


if __name__ == "__main__":
    if not available():
        print("reference not present at %s; nothing to check" % REF)
        sys.exit(0)
    sys.exit(main())
