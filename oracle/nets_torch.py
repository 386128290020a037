"""PyTorch-CPU fp32 restatement of the reference's TF-1.x graph (TEST INFRASTRUCTURE).

PARITY UNPINNED: the arithmetic of this half of the path lives in TensorFlow 1.x
(unpinned, not installable here: SURVEY.md section 8c); the reference holds no
golden vectors at the ``sess.run`` boundary.  This file restates the documented
TF semantics and is what the CUDA path is compared with.

Layout conventions are the reference's: activations NHWC, conv weights HWIO
``[kh, kw, Ci, Co]`` (isprs:706), variables named by TF scope
(``conv1/weights``, ``conv1/biases``, ``conv1/moving_mean``, ``conv1/moving_variance``,
``conv_classifier/weights`` ...; isprs Dilated6 uses ``main_conv1..6`` isprs:766-777).
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_DECAY = 0.999   # tf.contrib.layers.batch_norm default (isprs:658)
BN_EPS = 0.001


# (kernel, rate, Co, concat_with_input)  -- SURVEY.md Appendix A
NET_SPECS = {
    # isprs:761-788
    "dilated_icpr_original": dict(act="relu", pool=False, dense=False, scope="main_conv",
                                  convs=[(5, 1, 64), (5, 1, 64), (4, 2, 128), (4, 2, 128), (3, 4, 256), (3, 4, 256)]),
    # isprs:962-993
    "dilated_grsl": dict(act="lrelu", pool=True, dense=False, scope="conv",
                         convs=[(5, 1, 64), (5, 2, 64), (4, 3, 128), (4, 4, 128), (3, 5, 256), (3, 6, 256)]),
    # isprs:914-959
    "dilated_icpr_rate6_densely": dict(act="relu", pool=False, dense=True, scope="conv",
                                       convs=[(5, 1, 32), (5, 2, 32), (4, 3, 64), (4, 4, 64), (3, 5, 128), (3, 6, 128)]),
    # isprs:996-1033
    "dilated_grsl_rate8": dict(act="lrelu", pool=True, dense=False, scope="conv",
                               convs=[(5, 1, 64), (5, 2, 64), (4, 3, 128), (4, 4, 128), (3, 5, 192), (3, 6, 192),
                                      (3, 7, 256), (3, 8, 256)]),
    # isprs:886-911, 791-816, 852-883 (the last one calls tf.nn.conv2d: rate 1)
    "dilated_icpr_rate6": dict(act="relu", pool=False, dense=False, scope="conv",
                               convs=[(5, 1, 64), (5, 2, 64), (4, 3, 128), (4, 4, 128), (3, 5, 256), (3, 6, 256)]),
    "dilated_icpr_rate6_small": dict(act="relu", pool=False, dense=False, scope="conv",
                                     convs=[(5, 1, 64), (5, 2, 64), (4, 3, 64), (4, 4, 128), (3, 5, 128), (3, 6, 128)]),
    "dilated_icpr_rate6_nodilation": dict(act="relu", pool=False, dense=False, scope="conv",
                                          convs=[(5, 1, 64), (5, 1, 64), (4, 1, 128), (4, 1, 128), (3, 1, 256), (3, 1, 256)]),
}
NET_SPECS["dilated_icpr_rate1"] = dict(act="relu", pool=False, dense=False, scope="conv",       # coffee:788-813
                                       convs=[(5, 1, 64), (5, 1, 64), (4, 1, 128), (4, 1, 128), (3, 1, 256), (3, 1, 256)])
NET_SPECS["dilated_icpr_vary_rate"] = dict(act="relu", pool=False, dense=False, scope="conv",   # coffee:816-841
                                           convs=[(5, 1, 64), (5, 2, 64), (4, 4, 128), (4, 1, 128), (3, 2, 256), (3, 4, 256)])
NET_SPECS["dilated_icpr_old"] = dict(act="relu", pool=False, dense=False, scope="conv", scopes=(1, 3, 5),   # contest:574-603
                                     convs=[(5, 1, 64), (4, 2, 128), (3, 4, 256)])
NET_SPECS["dilated_grsl_old"] = NET_SPECS["dilated_grsl"]                                       # contest:606-636
_R6 = [(5, 1, 64), (5, 2, 64), (4, 3, 128), (4, 4, 128), (3, 5, 256), (3, 6, 256)]
NET_SPECS["dilated_icpr_rate6_avgpool"] = dict(act="relu", pool=False, dense=False, scope="conv", convs=_R6,      # isprs:819-849
                                               post=[("avg", 5), ("avg", 5), ("avg", 5), ("avg", 7), ("avg", 7), None])
NET_SPECS["dilated_icpr_rate6_SE"] = dict(act="relu", pool=False, dense=False, scope="conv", convs=_R6,           # isprs:1036-1061
                                          post=[None, ("se", 4, "se1"), None, ("se", 4, "se2"), None, ("se", 4, "se3")])
NET_SPECS["dilated_icpr_rate6_squeeze"] = dict(act="relu", pool=False, dense=False, scope="conv", convs=[(5, 1, 64)],   # isprs:1064-1086
                                               squeeze=[("conv2", 64, 64, 32, 5, 2), ("conv3", 64, 128, 64, 4, 3),
                                                        ("conv4", 128, 128, 64, 4, 4), ("conv5", 128, 256, 64, 3, 5),
                                                        ("conv6", 256, 256, 128, 3, 6)])
NET_SPECS["dilated8_grsl"] = NET_SPECS["dilated_grsl_rate8"]   # isprs CLI key (isprs:1672-1673)   # isprs CLI key (isprs:1672-1673)


def same_pad(k, rate):
    """TF SAME for stride 1: total=(k-1)*rate, before=floor(total/2), after=rest (Appendix A)."""
    total = (k - 1) * rate
    return total // 2, total - total // 2


def layer_plan(net_type, channels):
    """[(scope, k, rate, Ci, Co)] for the conv stack + classifier input width."""
    spec = NET_SPECS[net_type]
    plan = []
    cin = channels
    for i, (k, r, co) in enumerate(spec["convs"]):
        plan.append(("%s%d" % (spec["scope"], spec.get("scopes", range(1, 99))[i]), k, r, cin, co))
        if spec["dense"]:
            cin = co if i == 0 else cin + co     # c1=[conv1,conv2], c2=[c1,conv3] ... (isprs:921-948)
        else:
            cin = co
    for (name, in_dim, out_dim, k_dim, ksz, rate) in spec.get("squeeze", ()):      # _squeeze_conv_layer (isprs:726-742)
        plan.append((name + "_s1", 1, rate, in_dim, k_dim))
        plan.append((name + "_s2_1", 1, rate, k_dim, out_dim // 2))
        plan.append((name + "_s2_2", ksz, rate, k_dim, out_dim // 2))
        cin = out_dim
    return plan, cin


def se_blocks(net_type):
    spec = NET_SPECS[net_type]
    return [(i, po[2], spec["convs"][i][2], spec["convs"][i][2] // po[1]) for i, po in enumerate(spec.get("post", ()))
            if po is not None and po[0] == "se"]


def init_params(net_type, channels, num_classes, seed):
    """Xavier-uniform weights, bias 0.1 (classifier 0.0), BN moving stats 0/1 (Appendix B.2-3)."""
    rs = np.random.RandomState(seed)
    plan, cls_in = layer_plan(net_type, channels)
    p = OrderedDict()
    for scope, k, r, ci, co in plan:
        lim = math.sqrt(6.0 / (k * k * ci + k * k * co))
        p[scope + "/weights"] = rs.uniform(-lim, lim, size=(k, k, ci, co)).astype(np.float32)
        p[scope + "/biases"] = np.full((co,), 0.1, dtype=np.float32)
        p[scope + "/moving_mean"] = np.zeros((co,), dtype=np.float32)
        p[scope + "/moving_variance"] = np.ones((co,), dtype=np.float32)
    lim = math.sqrt(6.0 / (cls_in + num_classes))
    for _, name, c, r in se_blocks(net_type):        # _fc_layer: truncated normal (stddev 0.005), bias 0.1 (isprs:666-679)
        for nm, shape in ((name + "_fc1", (c, r)), (name + "_fc2", (r, c))):
            w = rs.normal(0.0, 0.005, size=shape)
            while np.any(np.abs(w) > 0.01):
                bad = np.abs(w) > 0.01
                w[bad] = rs.normal(0.0, 0.005, size=int(bad.sum()))
            p[nm + "/weights"] = w.astype(np.float32)
            p[nm + "/biases"] = np.full((shape[1],), 0.1, dtype=np.float32)
    p["conv_classifier/weights"] = rs.uniform(-lim, lim, size=(1, 1, cls_in, num_classes)).astype(np.float32)
    p["conv_classifier/biases"] = np.zeros((num_classes,), dtype=np.float32)
    return p


def _conv_same(x_nchw, w_hwio, rate):
    k = w_hwio.shape[0]
    pb, pa = same_pad(k, rate)
    w = w_hwio.permute(3, 2, 0, 1).contiguous()          # HWIO -> OIHW, cross-correlation (Appendix B.1)
    x = F.pad(x_nchw, (pb, pa, pb, pa))
    return F.conv2d(x, w, dilation=rate)


class _RoundBf16(torch.autograd.Function):
    """Storage rounding of the tensor-core path, forward AND backward: the CUDA step keeps Z, the layer outputs and their
    gradients in bf16 (fp32 accumulation in between), so the emulating oracle rounds at the same points."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _rb(x, on):
    return _RoundBf16.apply(x) if on else x


class OracleNet:
    """Holds fp32 parameters as torch tensors; forward in train or eval mode.

    ``emulate_bf16``: round conv operands (input, filter), the raw conv output Z, the layer output and the gradients
    flowing through those points to bf16, like the tensor-core training path stores them.  The arithmetic in between stays
    fp32.  Activation gates and max-pool winners are then decided on the same rounded values as on the GPU, which removes
    the chaotic gate flips from a gradient comparison (tests/test_gpu_parity.py::test_train_step_bf16_vs_emulating_oracle)."""

    def __init__(self, net_type, channels, num_classes, params, bn_unbiased_ema=True, emulate_bf16=False, conv_noise=0.0):
        self.emulate_bf16 = bool(emulate_bf16)
        # relative Gaussian noise on every convolution output BEFORE it is rounded: stands for a different fp32 summation
        # order (tensor cores vs CPU, ~1e-6).  Used by the tests to measure how ill-conditioned a step is: with bf16 storage a
        # 1e-6 perturbation crosses rounding boundaries and moves the filter gradients by several per cent.
        self.conv_noise = float(conv_noise)
        self._noise_gen = torch.Generator().manual_seed(12345)
        self.net_type = net_type
        self.spec = NET_SPECS[net_type]
        self.channels = channels
        self.num_classes = num_classes
        self.plan, self.cls_in = layer_plan(net_type, channels)
        self.p = OrderedDict((k, torch.tensor(np.asarray(v), dtype=torch.float32)) for k, v in params.items())
        self.momentum = OrderedDict()
        self.global_step = 0
        # Appendix B.3: the fused TF implementation feeds the Bessel-corrected batch
        # variance into the EMA while normalising with the biased one (flag).
        self.bn_unbiased_ema = bn_unbiased_ema

    def trainable(self):
        return [k for k in self.p if k.endswith("/weights") or k.endswith("/biases")]

    def _act(self, x):
        if self.spec["act"] == "relu":
            return torch.relu(x)
        return torch.maximum(0.1 * x, x)                   # isprs:620-621

    def forward(self, x_flat, crop, is_training, p=None, update_stats=True, taps=None, ztaps=None):
        """x_flat: [B, crop*crop*C] (isprs:763 reshape).  Returns logits NHWC [B,crop,crop,K]."""
        p = self.p if p is None else p
        B = x_flat.shape[0]
        e = self.emulate_bf16
        x = _rb(x_flat.reshape(B, crop, crop, self.channels).permute(0, 3, 1, 2), e)

        def conv_bn_act(inp, scope, r):
            """_conv_layer (isprs:700-723): atrous conv + bias -> batch_norm(center=False, scale=False) -> activation."""
            z = _conv_same(inp, _rb(p[scope + "/weights"], e), r) + p[scope + "/biases"].view(1, -1, 1, 1)
            if self.conv_noise:
                z = z * (1.0 + self.conv_noise * torch.randn(z.shape, generator=self._noise_gen))
            z = _rb(z, e)
            if ztaps is not None:
                ztaps[scope] = z
            if is_training:
                mean = z.mean(dim=(0, 2, 3))
                var = z.var(dim=(0, 2, 3), unbiased=False)
                if update_stats:
                    n = z.numel() // z.shape[1]
                    var_ema = var * (n / max(n - 1, 1)) if self.bn_unbiased_ema else var
                    with torch.no_grad():
                        self.p[scope + "/moving_mean"].mul_(BN_DECAY).add_((1 - BN_DECAY) * mean.detach())
                        self.p[scope + "/moving_variance"].mul_(BN_DECAY).add_((1 - BN_DECAY) * var_ema.detach())
            else:
                mean = p[scope + "/moving_mean"]
                var = p[scope + "/moving_variance"]
            zh = (z - mean.view(1, -1, 1, 1)) * torch.rsqrt(var.view(1, -1, 1, 1) + BN_EPS)
            return self._act(zh)

        feats = None
        post = self.spec.get("post")
        n_plain = len(self.spec["convs"])
        for i, (scope, k, r, ci, co) in enumerate(self.plan[:n_plain]):
            a = conv_bn_act(x, scope, r)
            if self.spec["pool"]:
                a = F.max_pool2d(a, 3, 1, 1)                # SAME: -inf padding (Appendix B.4)
            po = post[i] if post else None
            if po is not None and po[0] == "avg":
                # tf.nn.avg_pool(SAME, stride 1): the mean over the in-image part of the window (isprs:753-758)
                a = F.avg_pool2d(a, po[1], 1, po[1] // 2, count_include_pad=False)
            elif po is not None and po[0] == "se":
                # _squeeze_excitation_layer (isprs:682-697): global mean -> FC -> ReLU -> FC -> sigmoid -> channel gate
                sq = a.mean(dim=(2, 3))
                ex = torch.relu(sq @ p[po[2] + "_fc1/weights"] + p[po[2] + "_fc1/biases"])
                ex = torch.sigmoid(ex @ p[po[2] + "_fc2/weights"] + p[po[2] + "_fc2/biases"])
                a = a * ex.view(ex.shape[0], -1, 1, 1)
            a = _rb(a, e)
            if taps is not None:
                taps[scope] = a.permute(0, 2, 3, 1).detach()
            if self.spec["dense"]:
                feats = a if i == 0 else torch.cat([feats, a], dim=1)
                x = feats
            else:
                x = a
        for (name, in_dim, out_dim, k_dim, ksz, rate) in self.spec.get("squeeze", ()):
            # _squeeze_conv_layer (isprs:726-742): 1x1 squeeze, then a 1x1 and a kxk dilated expand branch, concatenated
            s1 = _rb(conv_bn_act(x, name + "_s1", rate), e)
            b1 = conv_bn_act(s1, name + "_s2_1", rate)
            b2 = conv_bn_act(s1, name + "_s2_2", rate)
            x = _rb(torch.cat([b1, b2], dim=1), e)
            if taps is not None:
                taps[name] = x.permute(0, 2, 3, 1).detach()
        wc = p["conv_classifier/weights"]
        logits = F.conv2d(x, wc.permute(3, 2, 0, 1).contiguous()) + p["conv_classifier/biases"].view(1, -1, 1, 1)
        return logits.permute(0, 2, 3, 1).contiguous()

    def loss(self, logits, labels, weight_decay, mask=None, p=None):
        """isprs:1089-1099 / contest:881-901: mean CE (+mask) + sum wd*l2_loss(W)."""
        p = self.p if p is None else p
        lg = logits.reshape(-1, self.num_classes)
        lb = labels.reshape(-1).long()
        if mask is not None:
            m = mask.reshape(-1).bool()
            lg, lb = lg[m], lb[m]
        ce = F.cross_entropy(lg, lb, reduction="mean")
        l2 = 0.0
        for k in p:
            if k.endswith("/weights"):
                l2 = l2 + weight_decay * 0.5 * (p[k] ** 2).sum()
        return ce + l2

    def train_step(self, x_flat, y_flat, crop, lr0, weight_decay, decay_steps=50000, decay_rate=0.5,
                   momentum=0.9, mask=None):
        """One ``sess.run([optimizer, loss, pred_up])`` (isprs:1750-1752).

        Returns (loss, pred int64 [B,crop,crop], logits).  loss/pred come from the same
        train-mode forward with pre-update weights (Appendix B.9)."""
        names = self.trainable()
        leaf = OrderedDict((k, (v.clone().requires_grad_(True) if k in names else v)) for k, v in self.p.items())
        logits = self.forward(x_flat, crop, True, p=leaf)
        loss = self.loss(logits, y_flat, weight_decay, mask=mask, p=leaf)
        grads = torch.autograd.grad(loss, [leaf[k] for k in names], allow_unused=True)
        lr = lr0 * (decay_rate ** (self.global_step // decay_steps))   # staircase (isprs:1686)
        with torch.no_grad():
            for k, g in zip(names, grads):
                if g is None:
                    g = torch.zeros_like(self.p[k])
                acc = self.momentum.get(k)
                acc = g.clone() if acc is None else momentum * acc + g   # Appendix B.6 (slot starts at 0)
                self.momentum[k] = acc
                self.p[k] -= lr * acc
        self.global_step += 1
        pred = logits.detach().argmax(dim=3)
        self.last_grads = OrderedDict(zip(names, grads))
        return float(loss.detach()), pred, logits.detach()

    def infer(self, x_flat, crop):
        """``sess.run([pred_up, logits], is_training=False)`` (isprs:1274-1275)."""
        with torch.no_grad():
            logits = self.forward(x_flat, crop, False)
        return logits.argmax(dim=3), logits

    def export_params(self):
        return OrderedDict((k, v.detach().numpy().copy()) for k, v in self.p.items())


def macs_per_pixel(net_type, channels, num_classes):
    """Forward MACs per output pixel."""
    plan, cls_in = layer_plan(net_type, channels)
    return sum(k * k * ci * co for _, k, r, ci, co in plan) + cls_in * num_classes
