"""CPU: the PyTorch restatement of the TensorFlow graph (oracle/nets_torch.py) against an independent, loop-style NumPy
restatement of the documented TF 1.x semantics written from the op definitions, not from torch:

* tf.nn.atrous_conv2d(..., padding='SAME'): stride 1, effective extent (k-1)*rate+1, pad_total = (k-1)*rate,
  pad_before = pad_total // 2, the rest after (asymmetric for the 4x4 kernels at odd rates), HWIO cross-correlation;
* tf.contrib.layers.batch_norm(center=False, scale=False): batch mean / biased variance, eps 1e-3, EMA decay 0.999;
* tf.nn.max_pool 3x3 stride 1 SAME: the maximum over the in-image part of the window;
* sparse_softmax_cross_entropy mean + wd * l2_loss(W) (sum(W**2) / 2), MomentumOptimizer (accum = m*accum + g;
  var -= lr*accum), staircase exponential decay; tf.argmax = first maximum.

TensorFlow itself cannot be installed here (parity of this half is unpinned); this file at least
makes the oracle two independent restatements that have to agree, on small cases where a pure-Python loop finishes fast.
"""
import math

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from oracle import nets_torch  # noqa: E402


def np_atrous_same(x, w, rate):
    B, H, W, Ci = x.shape
    k, _, _, Co = w.shape
    total = (k - 1) * rate
    pb = total // 2
    out = np.zeros((B, H, W, Co), dtype=np.float64)
    for b in range(B):
        for i in range(H):
            for j in range(W):
                acc = np.zeros(Co, dtype=np.float64)
                for u in range(k):
                    ii = i - pb + u * rate
                    if ii < 0 or ii >= H:
                        continue
                    for v in range(k):
                        jj = j - pb + v * rate
                        if jj < 0 or jj >= W:
                            continue
                        acc += x[b, ii, jj, :].astype(np.float64) @ w[u, v].astype(np.float64)
                out[b, i, j] = acc
    return out


def np_maxpool3_same(x):
    B, H, W, C = x.shape
    out = np.empty_like(x)
    for i in range(H):
        for j in range(W):
            out[:, i, j, :] = x[:, max(i - 1, 0):min(i + 2, H), max(j - 1, 0):min(j + 2, W), :].max(axis=(1, 2))
    return out


@pytest.mark.parametrize("k,rate", [(5, 1), (5, 2), (4, 1), (4, 2), (4, 3), (3, 4), (3, 6)])
def test_atrous_same_padding_rule(k, rate):
    rs = np.random.RandomState(k * 10 + rate)
    x = rs.randn(2, 9, 9, 3).astype(np.float32)
    w = rs.randn(k, k, 3, 4).astype(np.float32)
    ref = np_atrous_same(x, w, rate)
    got = nets_torch._conv_same(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(w), rate).permute(0, 2, 3, 1).numpy()
    assert np.allclose(got, ref, atol=1e-4)
    # known answer: an all-ones 4x4 kernel at rate 1 counts the in-image taps; SAME puts 1 row before and 2 after
    if (k, rate) == (4, 1):
        ones = nets_torch._conv_same(torch.ones(1, 1, 5, 5), torch.ones(4, 4, 1, 1), 1)[0, 0].numpy()
        assert ones[0, 0] == 9 and ones[4, 4] == 4 and ones[2, 2] == 16 and ones[0, 4] == 6


def np_forward(net, params, x, crop, C, K, training):
    """Loop-style forward of a whole net (small shapes).  Returns (logits NHWC float64, batch stats per layer)."""
    spec = nets_torch.NET_SPECS[net]
    plan, _ = nets_torch.layer_plan(net, C)
    a = x.reshape(-1, crop, crop, C).astype(np.float64)
    feats, stats = None, {}
    for i, (scope, k, r, ci, co) in enumerate(plan):
        z = np_atrous_same(a, params[scope + "/weights"], r) + params[scope + "/biases"].astype(np.float64)
        if training:
            mean, var = z.mean(axis=(0, 1, 2)), z.var(axis=(0, 1, 2))
            stats[scope] = (mean, var, z.shape[0] * z.shape[1] * z.shape[2])
        else:
            mean, var = params[scope + "/moving_mean"].astype(np.float64), params[scope + "/moving_variance"].astype(np.float64)
        zh = (z - mean) / np.sqrt(var + 1e-3)
        act = np.maximum(zh, 0.0) if spec["act"] == "relu" else np.maximum(0.1 * zh, zh)
        if spec["pool"]:
            act = np_maxpool3_same(act)
        if spec["dense"]:
            feats = act if i == 0 else np.concatenate([feats, act], axis=3)
            a = feats
        else:
            a = act
    logits = a @ params["conv_classifier/weights"][0, 0].astype(np.float64) + params["conv_classifier/biases"].astype(np.float64)
    return logits, stats


@pytest.mark.parametrize("net,C,K", [("dilated_icpr_original", 3, 4), ("dilated_grsl", 4, 3), ("dilated_icpr_rate6_densely", 3, 3),
                                     ("dilated_grsl_rate8", 3, 2), ("dilated_icpr_rate6_small", 4, 3)])
def test_network_forward_and_bn_statistics(net, C, K):
    crop, B = 7, 2
    params = nets_torch.init_params(net, C, K, seed=3)
    # thin the nets: the loop reference is O(pixels * k^2 * Ci * Co) in Python -- keep the graph, shrink the widths
    rs = np.random.RandomState(1)
    x = rs.randn(B, crop * crop * C).astype(np.float32)
    orc = nets_torch.OracleNet(net, C, K, params, bn_unbiased_ema=False)
    # eval mode (moving statistics)
    for name in params:
        if name.endswith("/moving_mean"):
            params[name][:] = rs.randn(*params[name].shape) * 0.1
        if name.endswith("/moving_variance"):
            params[name][:] = rs.rand(*params[name].shape) + 0.5
    orc = nets_torch.OracleNet(net, C, K, params, bn_unbiased_ema=False)
    pred, logits = orc.infer(torch.from_numpy(x), crop)
    ref, _ = np_forward(net, params, x, crop, C, K, training=False)
    assert np.allclose(logits.numpy(), ref, atol=2e-3, rtol=1e-3)
    margin = np.sort(ref, axis=3)
    sure = (margin[..., -1] - margin[..., -2]) > 1e-2
    assert np.array_equal(pred.numpy()[sure], np.argmax(ref, axis=3)[sure])
    # train mode: batch statistics, biased variance; EMA with decay 0.999
    before = {k: v.copy() for k, v in params.items()}
    logits_t = orc.forward(torch.from_numpy(x), crop, True).detach().numpy()
    ref_t, stats = np_forward(net, before, x, crop, C, K, training=True)
    assert np.allclose(logits_t, ref_t, atol=5e-3, rtol=1e-3)
    for scope, (mean, var, n) in stats.items():
        mm = 0.999 * before[scope + "/moving_mean"] + 0.001 * mean
        mv = 0.999 * before[scope + "/moving_variance"] + 0.001 * var
        assert np.allclose(orc.p[scope + "/moving_mean"].numpy(), mm, atol=1e-5)
        assert np.allclose(orc.p[scope + "/moving_variance"].numpy(), mv, atol=1e-5)


def test_loss_and_momentum_step():
    net, C, K, crop, B = "dilated_icpr_original", 3, 3, 5, 2
    params = nets_torch.init_params(net, C, K, seed=8)
    rs = np.random.RandomState(2)
    x = rs.randn(B, crop * crop * C).astype(np.float32)
    y = rs.randint(0, K, size=(B, crop * crop)).astype(np.float32)
    wd, lr0 = 0.005, 0.01
    orc = nets_torch.OracleNet(net, C, K, params)
    before = orc.export_params()
    loss, pred, logits = orc.train_step(torch.from_numpy(x), torch.from_numpy(y), crop, lr0, wd, decay_steps=1, decay_rate=0.5)
    # loss = mean CE over every pixel + sum over `weights` variables of wd * sum(w^2)/2   (isprs:1089-1099, 640-652)
    lg = logits.numpy().reshape(-1, K).astype(np.float64)
    lse = np.log(np.exp(lg - lg.max(1, keepdims=True)).sum(1)) + lg.max(1)
    ce = float(np.mean(lse - lg[np.arange(len(lg)), y.reshape(-1).astype(int)]))
    l2 = sum(wd * 0.5 * float((v.astype(np.float64) ** 2).sum()) for k, v in before.items() if k.endswith("/weights"))
    assert abs(loss - (ce + l2)) < 1e-4
    assert np.array_equal(pred.numpy(), np.argmax(logits.numpy(), axis=3))          # first maximum
    # first step: accum = g, var -= lr*g; the weight-decay gradient wd*W is part of g; biases behind BN get g = 0
    g = orc.last_grads
    for name in ("main_conv3/weights", "conv_classifier/weights", "conv_classifier/biases"):
        assert np.allclose(orc.p[name].numpy(), before[name] - lr0 * g[name].numpy(), atol=1e-7)
    assert float(g["main_conv2/biases"].abs().max()) < 1e-6
    # second step: staircase decay (decay_steps=1 -> lr halves), accum = 0.9*accum + g
    acc1 = {k: v.clone() for k, v in orc.momentum.items()}
    mid = orc.export_params()
    orc.train_step(torch.from_numpy(x), torch.from_numpy(y), crop, lr0, wd, decay_steps=1, decay_rate=0.5)
    name = "main_conv6/weights"
    acc2 = 0.9 * acc1[name] + orc.last_grads[name]
    assert np.allclose(orc.p[name].numpy(), mid[name] - (lr0 * 0.5) * acc2.numpy(), atol=1e-7)
    assert math.isclose(lr0 * 0.5 ** (1 // 1), 0.005)
