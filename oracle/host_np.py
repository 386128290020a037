"""NumPy restatement of the reference's host data path (TEST INFRASTRUCTURE).

Each function cites the reference lines it follows (``isprs`` =
/root/reference/isprs_dilated_random.py, ``contest`` = contest_dilated_random.py,
``coffee`` = coffee_dilated_random.py).  Written as plain loops so it can be
audited against the reference line by line; pinned by ``tests/golden/*.npz``
(see ``oracle/make_golden.py``).  Never imported by the product.
"""
import math
import random

import numpy as np

VARIANTS = ("isprs", "contest", "coffee")


# --------------------------------------------------------------------------------------
# sliding-window grid  (isprs:337-400, contest:257-328, coffee:296-349)
# --------------------------------------------------------------------------------------
def grid_count(length, crop, stride):
    """Number of window positions along one axis (isprs:344-347, 1253-1256)."""
    if (length - crop) % stride == 0:
        return int((length - crop) / stride) + 1
    return int((length - crop) / stride) + 2


def sliding_stride(crop):
    """isprs:1243 -- stride = floor(crop / 2)."""
    return int(math.floor(crop / 2.0))


def n_batches(h, w, crop, stride, batch):
    """isprs:1264-1265 (intended integer semantics of the py2 expression)."""
    n = grid_count(h, crop, stride) * grid_count(w, crop, stride)
    return int(n / batch) + 1 if n % batch != 0 else int(n / batch)


def patch_positions(h, w, crop, stride, index, batch, variant="isprs"):
    """Positions (top row, left col) of the ``index``-th batch of the sliding window.

    isprs:344-400: row-major walk from (offset_h, offset_w); a window that sticks
    out of the scene is shifted back so that it ends on the border.  The three
    scripts differ only in how the start offsets are derived:
      isprs   -- div and mod by total_index_w                 (isprs:351-352)
      contest -- div by total_index_h, mod by total_index_w   (contest:275-276, bug F10)
      coffee  -- one total_index computed from h for both     (coffee:302-307)
    """
    th = grid_count(h, crop, stride)
    tw = grid_count(w, crop, stride)
    if variant == "isprs":
        div, mod = tw, tw
    elif variant == "contest":
        div, mod = th, tw
    elif variant == "coffee":
        div, mod, tw = th, th, th
    else:
        raise ValueError(variant)
    offset_h = int((index * batch) / div) * stride
    offset_w = int((index * batch) % mod) * stride
    pos = []
    first = True
    for j in range(offset_h, th * stride, stride):
        if not first:
            offset_w = 0
        for k in range(offset_w, tw * stride, stride):
            first = False
            cur_x, cur_y = j, k
            # lengths of the clipped slice data[cur_x:cur_x+crop, cur_y:cur_y+crop]
            len_x = max(0, min(cur_x + crop, h) - cur_x)
            len_y = max(0, min(cur_y + crop, w) - cur_y)
            if len_x != crop:
                cur_x = cur_x - (crop - len_x)
            if len_y != crop:
                cur_y = cur_y - (crop - len_y)
            pos.append((cur_x, cur_y))
            if len(pos) == batch:
                return pos
    return pos


def all_patch_positions(h, w, crop, batch, variant="isprs"):
    """Concatenation of every batch in visiting order (isprs:1264-1267)."""
    stride = sliding_stride(crop)
    out = []
    for i in range(n_batches(h, w, crop, stride, batch)):
        out.extend(patch_positions(h, w, crop, stride, i, batch, variant))
    return out


# --------------------------------------------------------------------------------------
# overlap accumulate + argmax  (isprs:1261-1284, contest:916-941, coffee:1045-1068)
# --------------------------------------------------------------------------------------
def accumulate_argmax(logits, positions, h, w, crop, return_mean=False):
    """Sequential ``prob_im += logits`` in patch order, then argmax(prob/occur).

    logits: float32 [P, crop, crop, K]; positions: P (row, col) pairs in visit order.
    isprs:1261-1262 zero fp32/uint32 images; 1276-1279 ``+=`` per patch; 1282
    occur==0 -> 1; 1284 float64 divide then first-max argmax.
    """
    K = logits.shape[-1]
    prob_im = np.zeros([h, w, K], dtype=np.float32)
    occur_im = np.zeros([h, w, K], dtype=np.uint32)
    for j in range(len(positions)):
        x, y = int(positions[j][0]), int(positions[j][1])
        prob_im[x:x + crop, y:y + crop, :] += logits[j, :, :, :]
        occur_im[x:x + crop, y:y + crop, :] += 1
    occur_im[np.where(occur_im == 0)] = 1
    mean = prob_im / occur_im.astype(float)
    labels = np.argmax(mean, axis=2)
    if return_mean:
        return labels, mean
    return labels


# --------------------------------------------------------------------------------------
# per-crop confusion / accuracy  (isprs:510-531, coffee:390-409, contest:331-350)
# --------------------------------------------------------------------------------------
def confusion_by_crop(true_crop, pred_crop, num_classes, masks=None):
    """Returns (acc, acc_norm, KxK uint32) as isprs.calc_accuracy_by_crop.

    acc_norm divides by num_classes even when classes are absent (isprs:526-529).
    """
    b, h, w = pred_crop.shape
    acc = 0
    cm = np.zeros((num_classes, num_classes), dtype=np.uint32)
    for i in range(b):
        for j in range(h):
            for k in range(w):
                if masks is None or masks[i, j, k]:
                    if true_crop[i, j, k] == pred_crop[i, j, k]:
                        acc += 1
                    cm[true_crop[i, j, k]][pred_crop[i, j, k]] += 1
    _sum = 0.0
    for i in range(num_classes):
        s = np.sum(cm[i])
        _sum += (cm[i][i] / float(s) if s != 0 else 0)
    return acc, _sum / float(num_classes), cm


def confusion_by_crop_contest(true_crop, pred_crop, num_classes, masks):
    """contest:331-350 -- ``mask[i,j,k] is True`` on a numpy bool is always False
    (SURVEY F11), so nothing is ever counted: acc = 0, acc_norm = 0, cm = 0."""
    cm = np.zeros((num_classes, num_classes), dtype=np.uint32)
    return 0, 0.0, cm


def scene_confusion(labels, pred, num_classes, ignore_label=None):
    """Per-pixel scene confusion (isprs:1289-1296 with ignore 6; contest:944-948 with 7)."""
    cm = np.zeros((num_classes, num_classes), dtype=np.uint32)
    lab = labels.reshape(-1).astype(np.int64)
    prd = pred.reshape(-1).astype(np.int64)
    keep = np.ones_like(lab, dtype=bool) if ignore_label is None else (lab != ignore_label)
    np.add.at(cm, (lab[keep], prd[keep]), 1)
    return cm


# --------------------------------------------------------------------------------------
# normalisation  (isprs:74-81)
# --------------------------------------------------------------------------------------
def normalize_images(data, mean_full, std_full):
    """In place; only channels 0..2."""
    for ch in range(3):
        data[:, :, :, ch] = np.subtract(data[:, :, :, ch], mean_full[ch])
    for ch in range(3):
        data[:, :, :, ch] = np.divide(data[:, :, :, ch], std_full[ch])


# --------------------------------------------------------------------------------------
# train-patch gather  (isprs:245-334 without rotation/noise; contest:192-254; coffee:241-293)
# --------------------------------------------------------------------------------------
def shift_back(cur_x, cur_y, crop, h, w):
    """Border rule shared by every gather in the reference (isprs:259-269)."""
    len_x = max(0, min(cur_x + crop, h) - cur_x)
    len_y = max(0, min(cur_y + crop, w) - cur_y)
    if len_x != crop:
        cur_x = cur_x - (crop - len_x)
    if len_y != crop:
        cur_y = cur_y - (crop - len_y)
    return cur_x, cur_y


FLIP_NONE, FLIP_UD, FLIP_LR = 0, 1, 2


def gather_patches(scenes, label_maps, instances, flips, crop):
    """Crop ``crop x crop`` windows at (map, x, y) with shift-back, then flip.

    ``flips[i]``: 0 none, 1 flipud, 2 fliplr (isprs:304-318 numbering).
    Returns patches [B,crop,crop,C] (scene dtype) and labels [B,crop,crop].
    """
    patches, labels = [], []
    for i in range(len(instances)):
        m, x, y = int(instances[i][0]), int(instances[i][1]), int(instances[i][2])
        h, w = scenes[m].shape[0], scenes[m].shape[1]
        x, y = shift_back(x, y, crop, h, w)
        p = scenes[m][x:x + crop, y:y + crop, :]
        l = label_maps[m][x:x + crop, y:y + crop]
        if flips[i] == FLIP_UD:
            p, l = np.flipud(p), np.flipud(l)
        elif flips[i] == FLIP_LR:
            p, l = np.fliplr(p), np.fliplr(l)
        patches.append(p)
        labels.append(l)
    return np.asarray(patches), np.asarray(labels)


def contest_flip_of_index(i, n):
    """contest:197-252 / coffee:245-291: [0,n) none, [n,2n) fliplr, [2n,3n) flipud."""
    if i >= 2 * n:
        return i - 2 * n, FLIP_UD
    if i >= n:
        return i - n, FLIP_LR
    return i, FLIP_NONE


def contest_mask(labels):
    """contest:236-239 -- mask = (label != 7): 0->8, 7->0, astype(bool)."""
    m = np.copy(labels)
    m[m == 0] = 8
    m[m == 7] = 0
    return m.astype(bool)


# --------------------------------------------------------------------------------------
# patch-size policy  (isprs:46-71, 549-608, 1727-1737, 1757-1763)
# --------------------------------------------------------------------------------------
def select_batch(shuffle, batch_size, it, total_size):
    """isprs:46-58 (identical in contest and coffee)."""
    batch = shuffle[it:min(it + batch_size, total_size)]
    if min(it + batch_size, total_size) == total_size or total_size == it + batch_size:
        shuffle = np.asarray(random.sample(range(total_size), total_size))
        it = 0
        if len(batch) < batch_size:
            diff = batch_size - len(batch)
            batch_c = shuffle[it:it + diff]
            batch = np.concatenate((batch, batch_c))
            it = diff
    else:
        it += batch_size
    return shuffle, batch, it


def define_multinomial_probs(values, dif_prob=2):
    """isprs:61-71."""
    interval_size = values[-1] - values[0] + 1
    general_prob = 1.0 / float(interval_size)
    max_prob = general_prob * dif_prob
    probs = np.full(interval_size, (1.0 - max_prob * len(values)) / float(interval_size - len(values)))
    for i in range(len(values)):
        probs[values[i] - values[0]] = max_prob
    return probs


def draw_patch_size(distribution_type, values, probs=None):
    """isprs:1727-1737.  Consumes the legacy global ``np.random`` stream exactly as
    the reference does.  Returns (cur_patch_size, cur_size_int or None)."""
    if distribution_type == "multi_fixed":
        cur_size_int = np.random.randint(len(values))
        return int(values[cur_size_int]), cur_size_int
    if distribution_type == "uniform":
        cur_patch_size = int(np.random.uniform(values[0], values[-1] + 1, 1)[0])
        return cur_patch_size, cur_patch_size - values[0]
    if distribution_type == "multinomial":
        cur_size_int = np.random.multinomial(1, probs).argmax()
        return values[0] + cur_size_int, cur_size_int
    if distribution_type == "single_fixed":
        return int(values[0]), None
    raise ValueError(distribution_type)


def select_best_patch_size(distribution_type, values, patch_acc_loss, patch_occur, is_loss_or_acc="acc",
                           patch_chosen_values=None):
    """isprs:549-608.  Mutates patch_occur (0 -> 1) and patch_chosen_values in place."""
    patch_occur[np.where(patch_occur == 0)] = 1
    patch_mean = patch_acc_loss / patch_occur
    cur_patch_val = None
    if is_loss_or_acc == "acc":
        argmax_acc = np.argmax(patch_mean)
        if distribution_type == "multi_fixed":
            cur_patch_val = int(values[argmax_acc])
        else:
            cur_patch_val = values[0] + argmax_acc
        if patch_chosen_values is not None:
            patch_chosen_values[int(argmax_acc)] += 1
    elif is_loss_or_acc == "loss":
        arg_sort_out = np.argsort(patch_mean)
        n = len(values) if distribution_type == "multi_fixed" else values[-1] - values[0] + 1
        for i in range(n):
            if patch_occur[arg_sort_out[i]] > 0:
                if distribution_type == "multi_fixed":
                    cur_patch_val = int(values[arg_sort_out[i]])
                else:
                    cur_patch_val = values[0] + arg_sort_out[i]
                if patch_chosen_values is not None:
                    patch_chosen_values[arg_sort_out[i]] += 1
                break
    return cur_patch_val


# --------------------------------------------------------------------------------------
# gather + augment + normalise as ONE step: what the fed ``x`` / ``y`` of a batch must be
# (isprs:245-334 then 74-81 then the float32 feed; contest:192-254; coffee:241-293)
# --------------------------------------------------------------------------------------
def apply_plan(scenes, label_maps, inst, flips, crop, mean_full, std_full, noise=None, noise_on=None,
               over_x=None, over_y=None, over_on=None, cast=True, fp16_patches=False):
    """Checker for the gather kernel.  ``inst`` rows are (map, row, col) AFTER shift-back.

    Order of operations is the reference's: crop (or host-rotated override, isprs:292-296) -> + noise
    (isprs:298-301) -> flip (isprs:304-318) -> normalise channels 0..2 in the scene dtype (isprs:74-81) ->
    cast to float32 (feed_dict).  Returns x float32 [B,crop,crop,C], y float32 [B,crop,crop]."""
    xs, ys = [], []
    for b in range(len(inst)):
        m, r, c = int(inst[b][0]), int(inst[b][1]), int(inst[b][2])
        if over_on is not None and over_on[b]:
            p = np.array(over_x[b], dtype=np.float64)
            l = np.array(over_y[b])
        else:
            p = np.array(scenes[m][r:r + crop, c:c + crop, :])
            l = np.array(label_maps[m][r:r + crop, c:c + crop]) if label_maps is not None else np.zeros((crop, crop))
        if noise_on is not None and noise_on[b]:
            p = p + noise[b]
        f = 0 if flips is None else int(flips[b])
        if f == FLIP_UD:
            p, l = np.flipud(p), np.flipud(l)
        elif f == FLIP_LR:
            p, l = np.fliplr(p), np.fliplr(l)
        xs.append(p)
        ys.append(l)
    x = np.array(xs)
    if fp16_patches:
        # coffee:293 -- np.asarray(patches, dtype=np.float16), then normalize_images on the float16 array.  NumPy of the
        # reference era runs the float16 loop: scalar cast to half, operands widened to float32, result rounded to half.
        x = x.astype(np.float16)
        for ch in range(3):
            x[..., ch] = (x[..., ch].astype(np.float32) - np.float32(np.float16(mean_full[ch]))).astype(np.float16)
        for ch in range(3):
            x[..., ch] = (x[..., ch].astype(np.float32) / np.float32(np.float16(std_full[ch]))).astype(np.float16)
        return x.astype(np.float32), np.array(ys).astype(np.float32)
    if x.dtype == np.float32:
        mean_full = np.asarray(mean_full, dtype=np.float32)
        std_full = np.asarray(std_full, dtype=np.float32)
    normalize_images(x, mean_full, std_full)
    if not cast:
        return x, np.array(ys)
    return x.astype(np.float32), np.array(ys).astype(np.float32)
