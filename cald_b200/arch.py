"""Static description of the detectors the scoring engine runs.

Names are the torchvision ``state_dict`` keys the reference checkpoints carry
(reference: cald_train.py:351-356, 420-426; detection/frcnn_la.py:278-289;
torchvision resnet.py / feature_pyramid_network.py / rpn.py / faster_rcnn.py).
Only shapes and wiring live here; no arithmetic.
"""
from collections import OrderedDict

RESNET_BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}
FPN_CH = 256


def resnet_body_params(depth):
    """Ordered {name: shape} of ``backbone.body.*`` (bottleneck ResNet, stride on the 3x3)."""
    p = OrderedDict()

    def bn(prefix, c):
        for s in ("weight", "bias", "running_mean", "running_var"):
            p["%s.%s" % (prefix, s)] = (c,)

    p["backbone.body.conv1.weight"] = (64, 3, 7, 7)
    bn("backbone.body.bn1", 64)
    inplanes = 64
    for li, nblk in enumerate(RESNET_BLOCKS[depth]):
        planes = 64 * (2 ** li)
        for b in range(nblk):
            pre = "backbone.body.layer%d.%d" % (li + 1, b)
            p[pre + ".conv1.weight"] = (planes, inplanes, 1, 1)
            bn(pre + ".bn1", planes)
            p[pre + ".conv2.weight"] = (planes, planes, 3, 3)
            bn(pre + ".bn2", planes)
            p[pre + ".conv3.weight"] = (planes * 4, planes, 1, 1)
            bn(pre + ".bn3", planes * 4)
            if b == 0:
                p[pre + ".downsample.0.weight"] = (planes * 4, inplanes, 1, 1)
                bn(pre + ".downsample.1", planes * 4)
            inplanes = planes * 4
    return p


def frcnn_params(depth, num_classes):
    """Ordered {name: shape} for FRCNN_Feature(resnet_fpn_backbone(depth), num_classes)."""
    p = resnet_body_params(depth)
    for i, cin in enumerate((256, 512, 1024, 2048)):
        p["backbone.fpn.inner_blocks.%d.0.weight" % i] = (FPN_CH, cin, 1, 1)
        p["backbone.fpn.inner_blocks.%d.0.bias" % i] = (FPN_CH,)
    for i in range(4):
        p["backbone.fpn.layer_blocks.%d.0.weight" % i] = (FPN_CH, FPN_CH, 3, 3)
        p["backbone.fpn.layer_blocks.%d.0.bias" % i] = (FPN_CH,)
    p["rpn.head.conv.0.0.weight"] = (FPN_CH, FPN_CH, 3, 3)
    p["rpn.head.conv.0.0.bias"] = (FPN_CH,)
    p["rpn.head.cls_logits.weight"] = (3, FPN_CH, 1, 1)
    p["rpn.head.cls_logits.bias"] = (3,)
    p["rpn.head.bbox_pred.weight"] = (12, FPN_CH, 1, 1)
    p["rpn.head.bbox_pred.bias"] = (12,)
    p["roi_heads.box_head.fc6.weight"] = (1024, FPN_CH * 49)
    p["roi_heads.box_head.fc6.bias"] = (1024,)
    p["roi_heads.box_head.fc7.weight"] = (1024, 1024)
    p["roi_heads.box_head.fc7.bias"] = (1024,)
    p["roi_heads.box_predictor.cls_score.weight"] = (num_classes, 1024)
    p["roi_heads.box_predictor.cls_score.bias"] = (num_classes,)
    p["roi_heads.box_predictor.bbox_pred.weight"] = (4 * num_classes, 1024)
    p["roi_heads.box_predictor.bbox_pred.bias"] = (4 * num_classes,)
    return p


def retinanet_params(depth, num_classes):
    """Ordered {name: shape} for retinanet_resnet50_fpn_cal (retinanet_cal.py:584-625)."""
    p = resnet_body_params(depth)
    for i, cin in enumerate((512, 1024, 2048)):
        p["backbone.fpn.inner_blocks.%d.0.weight" % i] = (FPN_CH, cin, 1, 1)
        p["backbone.fpn.inner_blocks.%d.0.bias" % i] = (FPN_CH,)
    for i in range(3):
        p["backbone.fpn.layer_blocks.%d.0.weight" % i] = (FPN_CH, FPN_CH, 3, 3)
        p["backbone.fpn.layer_blocks.%d.0.bias" % i] = (FPN_CH,)
    for n in ("p6", "p7"):
        p["backbone.fpn.extra_blocks.%s.weight" % n] = (FPN_CH, FPN_CH, 3, 3)
        p["backbone.fpn.extra_blocks.%s.bias" % n] = (FPN_CH,)
    for head, last, cout in (("classification_head", "cls_logits", 9 * num_classes),
                             ("regression_head", "bbox_reg", 36)):
        for i in (0, 2, 4, 6):
            p["head.%s.conv.%d.weight" % (head, i)] = (FPN_CH, FPN_CH, 3, 3)
            p["head.%s.conv.%d.bias" % (head, i)] = (FPN_CH,)
        p["head.%s.%s.weight" % (head, last)] = (cout, FPN_CH, 3, 3)
        p["head.%s.%s.bias" % (head, last)] = (cout,)
    return p


# torchvision 0.8.2 (the reference's pin, README.md:10-11) names a few tensors
# without the Conv2dNormActivation ".0" level; accept both spellings on load.
_LEGACY = (
    ("backbone.fpn.inner_blocks.%d.weight", "backbone.fpn.inner_blocks.%d.0.weight"),
    ("backbone.fpn.inner_blocks.%d.bias", "backbone.fpn.inner_blocks.%d.0.bias"),
    ("backbone.fpn.layer_blocks.%d.weight", "backbone.fpn.layer_blocks.%d.0.weight"),
    ("backbone.fpn.layer_blocks.%d.bias", "backbone.fpn.layer_blocks.%d.0.bias"),
)


def canonical_key(name):
    """Map a 0.8.2-style key to the current spelling (identity otherwise)."""
    if name == "rpn.head.conv.weight":
        return "rpn.head.conv.0.0.weight"
    if name == "rpn.head.conv.bias":
        return "rpn.head.conv.0.0.bias"
    for old, new in _LEGACY:
        for i in range(4):
            if name == old % i:
                return new % i
    return name
