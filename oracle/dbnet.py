"""Oracle: DBNet++ forward, functional fp32 restatement (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows reference src/yomitoku/models/dbnet_plus.py:13-246 and models/layers/dbnet_feature_attention.py:36-79,
115-160, with the backbone being torchvision 0.26 `resnet50(replace_stride_with_dilation=[False, False, True])`
(dbnet_plus.py:30-38) written out as plain conv / batch-norm calls on the reference's state_dict keys
(SURVEY.md Appendix C).  Input: (1,3,H,W) fp32 normalised image; output: (1,1,H,W) fp32 sigmoid probability map
(the `binary` entry of the reference's OrderedDict, dbnet_plus.py:228-230).  The `thresh` branch exists in the
state_dict but is never executed by the reference forward (dbnet_plus.py:118-127 vs :228-230) and is ignored here.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, used by torchvision resnet50 and the decoder


def _bn(sd, prefix, x):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=False, eps=BN_EPS)


def _bottleneck(sd, p, x, stride, dilation):
    """torchvision Bottleneck (stride on conv2, v1.5): conv1 1x1 -> conv2 3x3 -> conv3 1x1 + identity."""
    out = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
    out = F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=dilation, dilation=dilation)
    out = F.relu(_bn(sd, p + ".bn2", out))
    out = _bn(sd, p + ".bn3", F.conv2d(out, sd[p + ".conv3.weight"]))
    if (p + ".downsample.0.weight") in sd:
        x = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
    return F.relu(out + x)


# (blocks, first-block stride, first-block dilation, later-block dilation); layer4 has its stride replaced by
# dilation: block 0 keeps dilation 1 with stride 1, blocks 1-2 use dilation 2.
_LAYERS = {
    "layer1": (3, 1, 1, 1),
    "layer2": (4, 2, 1, 1),
    "layer3": (6, 2, 1, 1),
    "layer4": (3, 1, 1, 2),
}


def backbone_features(sd, x, prefix="backbone.body."):
    """reference dbnet_plus.py:13-38 (IntermediateLayerGetter over resnet50): returns layer1..layer4 maps."""
    x = F.conv2d(x, sd[prefix + "conv1.weight"], stride=2, padding=3)
    x = F.relu(_bn(sd, prefix + "bn1", x))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    feats = {}
    for name, (blocks, stride0, dil0, dil) in _LAYERS.items():
        for i in range(blocks):
            x = _bottleneck(sd, "%s%s.%d" % (prefix, name, i), x, stride0 if i == 0 else 1, dil0 if i == 0 else dil)
        feats[name] = x
    return feats


def _up(x, size=None, scale=None):
    return F.interpolate(x, size=size, scale_factor=scale, mode="bilinear", align_corners=False)


def scale_feature_selection(sd, fuse, feats, prefix="decoder.concat_attention."):
    """reference dbnet_feature_attention.py:150-160 (ScaleFeatureSelection, type scale_channel_spatial) with
    ScaleChannelSpatialAttention.forward :69-79."""
    x = F.conv2d(fuse, sd[prefix + "conv.weight"], sd[prefix + "conv.bias"], padding=1)
    e = prefix + "enhanced_attention."
    g = x.mean(dim=(2, 3), keepdim=True)                                   # AdaptiveAvgPool2d(1)
    g = F.conv2d(F.relu(F.conv2d(g, sd[e + "channel_wise.1.weight"])), sd[e + "channel_wise.3.weight"])
    y = torch.sigmoid(g) + x                                                # :72-74
    m = y.mean(dim=1, keepdim=True)                                         # :76
    s = F.conv2d(F.relu(F.conv2d(m, sd[e + "spatial_wise.0.weight"], padding=1)), sd[e + "spatial_wise.2.weight"])
    z = torch.sigmoid(s) + y                                                # :77
    score = torch.sigmoid(F.conv2d(z, sd[e + "attention_wise.0.weight"]))   # :78, (1,4,H,W)
    return torch.cat([score[:, i:i + 1] * feats[i] for i in range(4)], dim=1)


def decoder_fuse(sd, feats, prefix="decoder."):
    """reference dbnet_plus.py:200-227: FPN + adaptive scale fusion, the (N,256,H/4,W/4) input of the binarize head."""
    names = ["layer1", "layer2", "layer3", "layer4"]
    f = {n: F.conv2d(feats[n], sd[prefix + "input_proj.%s.weight" % n]) for n in names}
    # top-down accumulation, cumulative; interpolate only when the sizes differ (:212)
    for lo, hi in (("layer4", "layer3"), ("layer3", "layer2"), ("layer2", "layer1")):
        b = f[lo]
        if b.shape[-2:] != f[hi].shape[-2:]:
            b = _up(b, size=f[hi].shape[-2:])
        f[hi] = b + f[hi]
    p1 = F.conv2d(f["layer1"], sd[prefix + "out_proj.layer1.weight"], padding=1)
    p2 = _up(F.conv2d(f["layer2"], sd[prefix + "out_proj.layer2.0.weight"], padding=1), scale=2)
    p3 = _up(F.conv2d(f["layer3"], sd[prefix + "out_proj.layer3.0.weight"], padding=1), scale=4)
    p4 = _up(F.conv2d(f["layer4"], sd[prefix + "out_proj.layer4.0.weight"], padding=1), scale=4)
    fp = [p4, p3, p2, p1]                                                   # fp[::-1] in the reference (:225-226)
    return scale_feature_selection(sd, torch.cat(fp, dim=1), fp, prefix + "concat_attention.")


def binarize_logits(sd, fuse, prefix="decoder."):
    """reference dbnet_plus.py:100-116 (the `binarize` Sequential) without the final sigmoid."""
    b = prefix + "binarize."
    x = F.relu(_bn(sd, b + "1", F.conv2d(fuse, sd[b + "0.weight"], padding=1)))
    x = F.relu(_bn(sd, b + "4", F.conv_transpose2d(x, sd[b + "3.weight"], sd[b + "3.bias"], stride=2)))
    return F.conv_transpose2d(x, sd[b + "6.weight"], sd[b + "6.bias"], stride=2)


def decoder_forward(sd, feats, prefix="decoder."):
    """reference dbnet_plus.py:200-230."""
    return torch.sigmoid(binarize_logits(sd, decoder_fuse(sd, feats, prefix), prefix))


@torch.inference_mode()
def dbnet_forward(sd, x):
    """reference DBNet.forward dbnet_plus.py:243-246; returns the `binary` map (1,1,H,W)."""
    return decoder_forward(sd, backbone_features(sd, x))
