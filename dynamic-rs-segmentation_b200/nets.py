"""Network specifications of the dilated FCNs and their variable initialisation.

Mirrors the reference's net builders (isprs_dilated_random.py:761-788, 914-959, 962-993, 996-1033)
and ``_conv_layer`` defaults (isprs:700-723): xavier-uniform conv weights, conv biases 0.1, classifier
bias 0.0, BN moving_mean 0 / moving_variance 1.  The compute itself lives in libdrs.so; this module only
describes shapes and names (TF variable scopes) so that checkpoints and parity tests line up.
"""
import math
from collections import OrderedDict

import numpy as np

# (kernel, rate, Co) per conv layer
SPECS = {
    "dilated_icpr_original": dict(act="relu", pool=False, dense=False,
                                  convs=[(5, 1, 64), (5, 1, 64), (4, 2, 128), (4, 2, 128), (3, 4, 256), (3, 4, 256)]),
    "dilated_grsl": dict(act="lrelu", pool=True, dense=False,
                         convs=[(5, 1, 64), (5, 2, 64), (4, 3, 128), (4, 4, 128), (3, 5, 256), (3, 6, 256)]),
    "dilated_icpr_rate6_densely": dict(act="relu", pool=False, dense=True,
                                       convs=[(5, 1, 32), (5, 2, 32), (4, 3, 64), (4, 4, 64), (3, 5, 128), (3, 6, 128)]),
    "dilated_grsl_rate8": dict(act="lrelu", pool=True, dense=False,
                               convs=[(5, 1, 64), (5, 2, 64), (4, 3, 128), (4, 4, 128), (3, 5, 192), (3, 6, 192),
                                      (3, 7, 256), (3, 8, 256)]),
    # plain six-layer stacks: isprs:886-911, 791-816, 852-883
    "dilated_icpr_rate6": dict(act="relu", pool=False, dense=False,
                               convs=[(5, 1, 64), (5, 2, 64), (4, 3, 128), (4, 4, 128), (3, 5, 256), (3, 6, 256)]),
    "dilated_icpr_rate6_small": dict(act="relu", pool=False, dense=False,
                                     convs=[(5, 1, 64), (5, 2, 64), (4, 3, 64), (4, 4, 128), (3, 5, 128), (3, 6, 128)]),
    "dilated_icpr_rate6_nodilation": dict(act="relu", pool=False, dense=False,
                                          convs=[(5, 1, 64), (5, 1, 64), (4, 1, 128), (4, 1, 128), (3, 1, 256), (3, 1, 256)]),
}
# further plain stacks of the other two scripts (same primitives)
SPECS["dilated_icpr_rate1"] = dict(act="relu", pool=False, dense=False,                    # coffee:788-813
                                   convs=[(5, 1, 64), (5, 1, 64), (4, 1, 128), (4, 1, 128), (3, 1, 256), (3, 1, 256)])
SPECS["dilated_icpr_vary_rate"] = dict(act="relu", pool=False, dense=False,                # coffee:816-841
                                       convs=[(5, 1, 64), (5, 2, 64), (4, 4, 128), (4, 1, 128), (3, 2, 256), (3, 4, 256)])
SPECS["dilated_icpr_old"] = dict(act="relu", pool=False, dense=False, scopes=(1, 3, 5),    # contest:574-603
                                 convs=[(5, 1, 64), (4, 2, 128), (3, 4, 256)])
SPECS["dilated_grsl_old"] = SPECS["dilated_grsl"]                                          # contest:606-636 (3 input channels)
# structural variants: same conv stack as dilated_icpr_rate6 with a post-op behind some layers
_R6 = [(5, 1, 64), (5, 2, 64), (4, 3, 128), (4, 4, 128), (3, 5, 256), (3, 6, 256)]
# SAME average pooling 5x5 / 7x7, stride 1, behind conv1..conv5 (isprs:819-849, coffee:721-751; dispatched by the coffee script)
SPECS["dilated_icpr_rate6_avgpool"] = dict(act="relu", pool=False, dense=False, convs=_R6,
                                           post=[("avg", 5), ("avg", 5), ("avg", 5), ("avg", 7), ("avg", 7), None])
# squeeze-and-excitation gate (ratio 4) behind conv2, conv4, conv6 (isprs:1036-1061, 682-697)
SPECS["dilated_icpr_rate6_SE"] = dict(act="relu", pool=False, dense=False, convs=_R6,
                                      post=[None, ("se", 4, "se1"), None, ("se", 4, "se2"), None, ("se", 4, "se3")])
# conv2..conv6 replaced by squeeze modules (isprs:1064-1086, 726-742): 1x1 in->k, then 1x1 k->out/2 and kxk dilated k->out/2, concat
SPECS["dilated_icpr_rate6_squeeze"] = dict(act="relu", pool=False, dense=False, convs=[(5, 1, 64)],
                                           squeeze=[("conv2", 64, 64, 32, 5, 2), ("conv3", 64, 128, 64, 4, 3), ("conv4", 128, 128, 64, 4, 4),
                                                    ("conv5", 128, 256, 64, 3, 5), ("conv6", 256, 256, 128, 3, 6)])
SPECS["dilated8_grsl"] = SPECS["dilated_grsl_rate8"]
NET_TYPES = tuple(SPECS)


def is_pooling(net_type):
    return bool(SPECS[net_type]["pool"])


def scope_prefix(net_type, isprs_scopes):
    # isprs names Dilated6's layers main_conv1..6 (isprs:766-777); coffee uses conv1..6 (coffee:635-662)
    return "main_conv" if (net_type == "dilated_icpr_original" and isprs_scopes) else "conv"


def layer_plan(net_type, channels, isprs_scopes=True):
    """[(scope, k, rate, Ci, Co)], classifier input width."""
    spec = SPECS[net_type]
    prefix = scope_prefix(net_type, isprs_scopes)
    plan, cin = [], channels
    for i, (k, r, co) in enumerate(spec["convs"]):
        plan.append(("%s%d" % (prefix, spec.get("scopes", range(1, 99))[i]), k, r, cin, co))
        if spec["dense"]:
            cin = co if i == 0 else cin + co
        else:
            cin = co
    for (name, in_dim, out_dim, k_dim, ksz, rate) in spec.get("squeeze", ()):      # _squeeze_conv_layer (isprs:726-742)
        plan.append((name + "_s1", 1, rate, in_dim, k_dim))
        plan.append((name + "_s2_1", 1, rate, k_dim, out_dim // 2))
        plan.append((name + "_s2_2", ksz, rate, k_dim, out_dim // 2))
        cin = out_dim
    return plan, cin


def se_blocks(net_type):
    """[(layer index, scope, channels, channels // ratio)] of the squeeze-and-excitation gates (isprs:682-697)."""
    spec = SPECS[net_type]
    out = []
    for i, po in enumerate(spec.get("post", ())):
        if po is not None and po[0] == "se":
            c = spec["convs"][i][2]
            out.append((i, po[2], c, c // po[1]))
    return out


def variable_shapes(net_type, channels, num_classes, isprs_scopes=True):
    plan, cls_in = layer_plan(net_type, channels, isprs_scopes)
    shapes = OrderedDict()
    for scope, k, r, ci, co in plan:
        shapes[scope + "/weights"] = (k, k, ci, co)
        shapes[scope + "/biases"] = (co,)
        shapes[scope + "/moving_mean"] = (co,)
        shapes[scope + "/moving_variance"] = (co,)
    for _, name, c, r in se_blocks(net_type):            # _fc_layer (isprs:666-679)
        shapes[name + "_fc1/weights"] = (c, r)
        shapes[name + "_fc1/biases"] = (r,)
        shapes[name + "_fc2/weights"] = (r, c)
        shapes[name + "_fc2/biases"] = (c,)
    shapes["conv_classifier/weights"] = (1, 1, cls_in, num_classes)
    shapes["conv_classifier/biases"] = (num_classes,)
    return shapes


def initial_variables(net_type, channels, num_classes, seed, isprs_scopes=True):
    """``sess.run(tf.initialize_all_variables())`` (isprs:1698, 1717) with a seeded generator.

    tf.contrib.layers.xavier_initializer_conv2d default: uniform(-l, l), l = sqrt(6 / (fan_in + fan_out)),
    fan_in = kh*kw*Ci, fan_out = kh*kw*Co (isprs:702)."""
    rs = np.random.RandomState(seed)
    out = OrderedDict()
    for name, shape in variable_shapes(net_type, channels, num_classes, isprs_scopes).items():
        if name.endswith("/weights") and len(shape) == 2:
            # tf.truncated_normal_initializer(stddev=0.005): draws beyond two standard deviations are redrawn (isprs:669)
            w = rs.normal(0.0, 0.005, size=shape)
            while np.any(np.abs(w) > 0.01):
                bad = np.abs(w) > 0.01
                w[bad] = rs.normal(0.0, 0.005, size=int(bad.sum()))
            out[name] = w.astype(np.float32)
        elif name.endswith("/weights"):
            kh, kw, ci, co = shape
            lim = math.sqrt(6.0 / (kh * kw * ci + kh * kw * co))
            out[name] = rs.uniform(-lim, lim, size=shape).astype(np.float32)
        elif name == "conv_classifier/biases":
            out[name] = np.zeros(shape, dtype=np.float32)
        elif name.endswith("/biases"):
            out[name] = np.full(shape, 0.1, dtype=np.float32)
        elif name.endswith("/moving_mean"):
            out[name] = np.zeros(shape, dtype=np.float32)
        else:
            out[name] = np.ones(shape, dtype=np.float32)
    return out


def macs_per_pixel(net_type, channels, num_classes):
    """Forward multiply-accumulates per output pixel."""
    plan, cls_in = layer_plan(net_type, channels)
    return sum(k * k * ci * co for _, k, r, ci, co in plan) + cls_in * num_classes


def tensor_core_macs_per_pixel(net_type, channels):
    """MACs per pixel of the layers that run on the tcgen05 kernel (all but conv1 and the classifier)."""
    plan, _ = layer_plan(net_type, channels)
    return sum(k * k * ci * co for _, k, r, ci, co in plan[1:])
