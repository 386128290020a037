"""Closed-form stand-in for the network (TEST INFRASTRUCTURE).

The golden vectors of the full-scene and training loops were produced by driving the reference's own
``validate_test`` / ``test`` / ``train`` through a fake ``sess.run`` whose "logits" are this function of the
fed patches (``oracle/make_golden.py``).  IEEE-exact operations only (add, mul, abs, fmod), so any
implementation that feeds bit-identical patches reproduces bit-identical logits.
"""
import numpy as np


def fake_logits(bx, crop, channels, num_classes):
    x = np.asarray(bx, dtype=np.float64).reshape(-1, crop, crop, channels)
    s = np.zeros(x.shape[:3], dtype=np.float64)
    for c in range(channels):
        s = s + x[..., c] * float(c + 1)
    out = np.empty(x.shape[:3] + (num_classes,), dtype=np.float32)
    for k in range(num_classes):
        out[..., k] = np.fmod(np.abs(s) * (3.0 + 2.0 * k) + 0.61 * k, 5.0).astype(np.float32)
    return out


def fake_train_fetches(bx, crop, channels, num_classes):
    """(loss, pred) the fake session returns for a training call."""
    logits = fake_logits(bx, crop, channels, num_classes)
    return np.float32(np.mean(logits.astype(np.float64))), np.argmax(logits, axis=3).astype(np.int64)
