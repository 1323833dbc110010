"""Seeded random state_dicts with the reference's exact key sets.  TEST INFRASTRUCTURE.

No pretrained weights are reachable offline, so parity and benchmarks run on seeded random weights of the named
architectures (the reference's `from_pretrained=False` path, base.py:80-86).  Distributions are chosen so that
activations stay O(1) through the networks (He-normal convs, near-identity batch-norm statistics with some spread so
that BN folding is actually exercised).  `peaked=True` for PARSeq scales the head and biases the EOS class so that
greedy decoding has comfortable top-1 margins and emits EOS at varied lengths - the setting in which
"character-identical strings" is a meaningful test (hard parts).
"""
import math

import torch


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _conv(g, cout, cin, kh, kw, gain=1.0):
    std = gain * math.sqrt(2.0 / (cin * kh * kw))
    return torch.randn(cout, cin, kh, kw, generator=g) * std


def _bn(sd, g, prefix, c, gamma=1.0):
    sd[prefix + ".weight"] = gamma * (1.0 + 0.1 * torch.randn(c, generator=g))
    sd[prefix + ".bias"] = 0.05 * torch.randn(c, generator=g)
    sd[prefix + ".running_mean"] = 0.05 * torch.randn(c, generator=g)
    sd[prefix + ".running_var"] = 1.0 + 0.2 * torch.rand(c, generator=g)
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def make_dbnet_state_dict(seed=0):
    """Keys of reference DBNet (models/dbnet_plus.py:233-246): backbone.body.* (torchvision resnet50 without
    avgpool/fc), decoder.{input_proj,out_proj,binarize,thresh,concat_attention}.*  - 363 tensors."""
    g = _g(seed)
    sd = {}
    p = "backbone.body."
    sd[p + "conv1.weight"] = _conv(g, 64, 3, 7, 7)
    _bn(sd, g, p + "bn1", 64)
    inplanes = 64
    for li, (planes, blocks) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
        for b in range(blocks):
            q = "%slayer%d.%d." % (p, li, b)
            sd[q + "conv1.weight"] = _conv(g, planes, inplanes, 1, 1)
            _bn(sd, g, q + "bn1", planes)
            sd[q + "conv2.weight"] = _conv(g, planes, planes, 3, 3)
            _bn(sd, g, q + "bn2", planes)
            sd[q + "conv3.weight"] = _conv(g, planes * 4, planes, 1, 1)
            _bn(sd, g, q + "bn3", planes * 4, gamma=0.5)   # damp the residual branch so depth does not blow up
            if b == 0:
                sd[q + "downsample.0.weight"] = _conv(g, planes * 4, inplanes, 1, 1, gain=0.7)
                _bn(sd, g, q + "downsample.1", planes * 4)
            inplanes = planes * 4
    d = "decoder."
    for i, c in enumerate((256, 512, 1024, 2048), start=1):
        sd["%sinput_proj.layer%d.weight" % (d, i)] = _conv(g, 256, c, 1, 1, gain=0.7)
    sd[d + "out_proj.layer1.weight"] = _conv(g, 64, 256, 3, 3)
    for i in (2, 3, 4):
        sd["%sout_proj.layer%d.0.weight" % (d, i)] = _conv(g, 64, 256, 3, 3)
    for name, cin in (("binarize", 256), ("thresh", 257)):
        q = d + name + "."
        sd[q + "0.weight"] = _conv(g, 64, cin, 3, 3)
        _bn(sd, g, q + "1", 64)
        sd[q + "3.weight"] = torch.randn(64, 64, 2, 2, generator=g) * math.sqrt(2.0 / 64)
        sd[q + "3.bias"] = 0.05 * torch.randn(64, generator=g)
        _bn(sd, g, q + "4", 64)
        sd[q + "6.weight"] = torch.randn(64, 1, 2, 2, generator=g) * math.sqrt(2.0 / 64)
        sd[q + "6.bias"] = 0.05 * torch.randn(1, generator=g)
    a = d + "concat_attention."
    sd[a + "conv.weight"] = _conv(g, 64, 256, 3, 3)
    sd[a + "conv.bias"] = 0.05 * torch.randn(64, generator=g)
    e = a + "enhanced_attention."
    sd[e + "channel_wise.1.weight"] = _conv(g, 16, 64, 1, 1)
    sd[e + "channel_wise.3.weight"] = _conv(g, 64, 16, 1, 1)
    sd[e + "spatial_wise.0.weight"] = torch.randn(1, 1, 3, 3, generator=g) * 0.5
    sd[e + "spatial_wise.2.weight"] = torch.randn(1, 1, 1, 1, generator=g)
    sd[e + "attention_wise.0.weight"] = _conv(g, 4, 64, 1, 1)
    return sd


def make_parseq_state_dict(spec, seed=0, peaked=False, degenerate_repeat=False):
    """Keys of reference PARSeq (models/parseq.py:49-96): encoder.* (timm ViT), decoder.layers.0.*, decoder.norm.*,
    head.*, text_embed.embedding.weight, pos_queries.

    peaked: head scaled so logits are far from uniform + EOS bias so rows stop at varied lengths.
    degenerate_repeat: token embeddings and positional queries made position/token independent so every AR step
    emits the same token -> exercises the repetition early-stop (parseq.py:226-242, 301-309)."""
    g = _g(seed)
    D = spec.embed_dim
    ph, pw = spec.patch
    gh, gw = spec.grid

    def lin(out_f, in_f, std=0.02):
        return torch.randn(out_f, in_f, generator=g).clamp_(-2, 2) * std

    def ln(sd, prefix):
        sd[prefix + ".weight"] = 1.0 + 0.05 * torch.randn(D, generator=g)
        sd[prefix + ".bias"] = 0.02 * torch.randn(D, generator=g)

    sd = {}
    e = "encoder."
    sd[e + "patch_embed.proj.weight"] = torch.randn(D, 3, ph, pw, generator=g) * math.sqrt(1.0 / (3 * ph * pw))
    sd[e + "patch_embed.proj.bias"] = 0.02 * torch.randn(D, generator=g)
    sd[e + "pos_embed"] = 0.2 * torch.randn(1, gh * gw, D, generator=g)
    wstd = 1.0 / math.sqrt(D)   # O(1) activations (std 0.02 would make every block a near no-op)
    for i in range(spec.enc_depth):
        p = "%sblocks.%d." % (e, i)
        ln(sd, p + "norm1")
        sd[p + "attn.qkv.weight"] = lin(3 * D, D, wstd)
        sd[p + "attn.qkv.bias"] = 0.02 * torch.randn(3 * D, generator=g)
        sd[p + "attn.proj.weight"] = lin(D, D, 0.5 * wstd)
        sd[p + "attn.proj.bias"] = 0.02 * torch.randn(D, generator=g)
        ln(sd, p + "norm2")
        sd[p + "mlp.fc1.weight"] = lin(spec.mlp_ratio * D, D, wstd)
        sd[p + "mlp.fc1.bias"] = 0.02 * torch.randn(spec.mlp_ratio * D, generator=g)
        sd[p + "mlp.fc2.weight"] = lin(D, spec.mlp_ratio * D, 0.5 / math.sqrt(spec.mlp_ratio * D))
        sd[p + "mlp.fc2.bias"] = 0.02 * torch.randn(D, generator=g)
    ln(sd, e + "norm")
    p = "decoder.layers.0."
    for att in ("self_attn", "cross_attn"):
        sd[p + att + ".in_proj_weight"] = lin(3 * D, D, wstd)
        sd[p + att + ".in_proj_bias"] = 0.02 * torch.randn(3 * D, generator=g)
        sd[p + att + ".out_proj.weight"] = lin(D, D, wstd)
        sd[p + att + ".out_proj.bias"] = 0.02 * torch.randn(D, generator=g)
    H = spec.dec_mlp_ratio * D
    sd[p + "linear1.weight"] = lin(H, D, wstd)
    sd[p + "linear1.bias"] = 0.02 * torch.randn(H, generator=g)
    sd[p + "linear2.weight"] = lin(D, H, 1.0 / math.sqrt(H))
    sd[p + "linear2.bias"] = 0.02 * torch.randn(D, generator=g)
    for n in ("norm1", "norm2", "norm_q", "norm_c"):
        ln(sd, p + n)
    ln(sd, "decoder.norm")
    C = spec.num_classes
    head_std = (6.0 if peaked else 1.0) / math.sqrt(D)
    sd["head.weight"] = torch.randn(C, D, generator=g) * head_std
    sd["head.bias"] = 0.02 * torch.randn(C, generator=g)
    sd["text_embed.embedding.weight"] = torch.randn(spec.num_tokens, D, generator=g) * (1.0 / math.sqrt(D))
    sd["pos_queries"] = torch.randn(1, spec.max_label_length + 1, D, generator=g) * 0.5
    if peaked:
        # EOS ramp: the positional query gains a component along the EOS class direction that grows with the
        # position, so rows stop at varied, image-dependent lengths (roughly 3..40 tokens).
        w_eos = sd["head.weight"][spec.eos_id] / sd["head.weight"][spec.eos_id].norm()
        ramp = (torch.arange(spec.max_label_length + 1, dtype=torch.float32) - 6.0) * 0.35
        sd["pos_queries"] = sd["pos_queries"] + ramp[None, :, None] * w_eos[None, None, :] * math.sqrt(D) * 0.5
    if degenerate_repeat:
        sd["text_embed.embedding.weight"][:] = sd["text_embed.embedding.weight"][:1]
        sd["pos_queries"][:] = sd["pos_queries"][:, :1]
        sd["head.bias"][spec.eos_id] = -50.0
    return sd
