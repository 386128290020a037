"""Host-side mirror of the reference's L2/L3 logic: batch selection, patch-size
policy, augmentation decisions, instance sampling.  Same function names, argument meaning and RNG
consumption as the reference scripts, so that a run seeded like the reference draws the same patch sizes,
batches and augmentations (bit-exact; pinned by tests/golden/host_golden.npz).

What is NOT here: anything that touches pixels every step.  Patch gathering, normalisation, the network,
the loss, the per-crop confusion matrix, overlap accumulation and argmax run in libdrs.so on the GPU
(``Session``).  The only per-pixel host work left is the nearest-neighbour rotation of the ~50 % of isprs
training patches that the reference rotates with ``scipy.ndimage.rotate`` (row N1 of SURVEY.md section 8f).
"""
import math
import random

import numpy as np

FLIP_NONE, FLIP_UD, FLIP_LR = 0, 1, 2


class BatchColors:      # isprs:20-28
    HEADER = '\033[95m'
    OKBLUE = '\033[94m'
    OKGREEN = '\033[92m'
    WARNING = '\033[93m'
    FAIL = '\033[91m'
    ENDC = '\033[0m'
    BOLD = '\033[1m'
    UNDERLINE = '\033[4m'


# ------------------------------------------------------------------------------------------------
# policy (L3) -- isprs:46-71, 549-608, 1727-1737, 1757-1763
# ------------------------------------------------------------------------------------------------
def select_batch(shuffle, batch_size, it, total_size):
    """isprs:46-58: epoch shuffle with wrap-around refill (python ``random`` stream)."""
    batch = shuffle[it:min(it + batch_size, total_size)]
    if min(it + batch_size, total_size) == total_size or total_size == it + batch_size:
        shuffle = np.asarray(random.sample(range(total_size), total_size))
        it = 0
        if len(batch) < batch_size:
            diff = batch_size - len(batch)
            batch = np.concatenate((batch, shuffle[it:it + diff]))
            it = diff
    else:
        it += batch_size
    return shuffle, batch, it


def define_multinomial_probs(values, dif_prob=2):
    """isprs:61-71: listed sizes get dif_prob/interval each, the rest share the remainder."""
    interval_size = values[-1] - values[0] + 1
    general_prob = 1.0 / float(interval_size)
    max_prob = general_prob * dif_prob
    probs = np.full(interval_size, (1.0 - max_prob * len(values)) / float(interval_size - len(values)))
    for i in range(len(values)):
        probs[values[i] - values[0]] = max_prob
    return probs


def init_score_arrays(distribution_type, values, occur_init=0):
    """isprs:2054-2064 (zeros); contest starts patch_occur at ones (contest:1275-1279)."""
    if distribution_type == 'multi_fixed':
        n = len(values)
    elif distribution_type in ('uniform', 'multinomial'):
        n = values[-1] - values[0] + 1
    else:
        return None, None, None
    patch_acc_loss = np.zeros(n, dtype=np.float32)
    patch_occur = np.full(n, occur_init, dtype=np.int32)
    patch_chosen_values = np.zeros(n, dtype=np.int32)
    return patch_acc_loss, patch_occur, patch_chosen_values


def draw_patch_size(distribution_type, values, probs=None, rng=np.random):
    """isprs:1727-1737: one draw from the legacy global ``np.random`` stream.  -> (size, index|None)"""
    if distribution_type == 'multi_fixed':
        cur_size_int = rng.randint(len(values))
        return int(values[cur_size_int]), cur_size_int
    if distribution_type == 'uniform':
        cur_patch_size = int(rng.uniform(values[0], values[-1] + 1, 1)[0])
        return cur_patch_size, cur_patch_size - values[0]
    if distribution_type == 'multinomial':
        cur_size_int = rng.multinomial(1, probs).argmax()
        return values[0] + cur_size_int, cur_size_int
    if distribution_type == 'single_fixed':
        return int(values[0]), None
    raise ValueError('distribution_type ' + str(distribution_type))


def update_scores(patch_acc_loss, patch_occur, cur_size_int, update_type, batch_loss, acc_norm, loss_scale=1.0):
    """isprs:1757-1763 (loss_scale = epoch_counter/10.0) / contest:1091-1094, coffee:1302-1307 (loss_scale = 1)."""
    patch_acc_loss[cur_size_int] += (batch_loss * loss_scale if update_type == 'loss' else acc_norm)
    patch_occur[cur_size_int] += 1


def select_best_patch_size(distribution_type, values, patch_acc_loss, patch_occur, is_loss_or_acc='acc',
                           patch_chosen_values=None, debug=False):
    """isprs:549-608.  Mutates patch_occur (0 -> 1) and patch_chosen_values in place like the reference."""
    patch_occur[np.where(patch_occur == 0)] = 1
    patch_mean = patch_acc_loss / patch_occur
    cur_patch_val = None
    if is_loss_or_acc == 'acc':
        argmax_acc = np.argmax(patch_mean)
        if distribution_type == 'multi_fixed':
            cur_patch_val = int(values[argmax_acc])
        elif distribution_type in ('uniform', 'multinomial'):
            cur_patch_val = values[0] + argmax_acc
        if patch_chosen_values is not None:
            patch_chosen_values[int(argmax_acc)] += 1
        if debug:
            print('patch_acc_loss', patch_acc_loss)
            print('patch_occur', patch_occur)
            print('patch_mean', patch_mean)
            print('argmax_acc', argmax_acc)
            print('specific', argmax_acc, patch_acc_loss[argmax_acc], patch_occur[argmax_acc], patch_mean[argmax_acc])
    elif is_loss_or_acc == 'loss':
        arg_sort_out = np.argsort(patch_mean)
        if debug:
            print('patch_acc_loss', patch_acc_loss)
            print('patch_occur', patch_occur)
            print('patch_mean', patch_mean)
            print('arg_sort_out', arg_sort_out)
        n = len(values) if distribution_type == 'multi_fixed' else values[-1] - values[0] + 1
        for i in range(n):
            if patch_occur[arg_sort_out[i]] > 0:
                if distribution_type == 'multi_fixed':
                    cur_patch_val = int(values[arg_sort_out[i]])
                else:
                    cur_patch_val = values[0] + arg_sort_out[i]
                if patch_chosen_values is not None:
                    patch_chosen_values[arg_sort_out[i]] += 1
                if debug:
                    print('specific', arg_sort_out[i], patch_acc_loss[arg_sort_out[i]], patch_occur[arg_sort_out[i]],
                          patch_mean[arg_sort_out[i]])
                break
    if debug:
        print('Current patch size ', cur_patch_val)
        if patch_chosen_values is not None:
            print('Distr of chosen sizes ', patch_chosen_values)
    return cur_patch_val


def acc_norm_from_cm(cm, num_classes):
    """Tail of calc_accuracy_by_crop (isprs:526-529): mean per-class recall, divisor always K."""
    _sum = 0.0
    for i in range(num_classes):
        s = np.sum(cm[i])
        _sum += (cm[i][i] / float(s) if s != 0 else 0)
    return _sum / float(num_classes)


# ------------------------------------------------------------------------------------------------
# one-off set-up (C13/C14 of SURVEY.md section 2): kept on the host, reference semantics
# ------------------------------------------------------------------------------------------------
def shift_back(cur_x, cur_y, crop_size, h, w):
    """Border rule of every gather in the reference (isprs:259-269): a window that sticks out is moved back."""
    len_x = max(0, min(cur_x + crop_size, h) - cur_x)
    len_y = max(0, min(cur_y + crop_size, w) - cur_y)
    if len_x != crop_size:
        cur_x = cur_x - (crop_size - len_x)
    if len_y != crop_size:
        cur_y = cur_y - (crop_size - len_y)
    return cur_x, cur_y


def create_distributions_over_classes(labels, crop_size, stride_crop, num_classes, verbose=True):
    """isprs:448-483: (map, x, y) of every reference-size window, bucketed by its majority class."""
    classes = [[] for _ in range(num_classes)]
    for k in range(len(labels)):
        w, h = labels[k].shape
        for i in range(0, w, stride_crop):
            for j in range(0, h, stride_crop):
                cur_x, cur_y = shift_back(i, j, crop_size, w, h)
                patch_class = labels[k][cur_x:cur_x + crop_size, cur_y:cur_y + crop_size]
                if patch_class.shape != (crop_size, crop_size):
                    raise ValueError("Error create_distributions_over_classes: Current patch size is " +
                                     str(len(patch_class)) + "x" + str(len(patch_class[0])))
                count = np.bincount(patch_class.astype(int).flatten())
                classes[int(np.argmax(count))].append((k, cur_x, cur_y))
    if verbose:
        for i in range(len(classes)):
            print(BatchColors.OKBLUE + 'Class ' + str(i + 1) + ' has length ' + str(len(classes[i])) + BatchColors.ENDC)
    return classes


def create_rotation_distribution(training_class_distribution, verbose=True):
    """isprs:486-496: one angle in [0,360) per instance (np.random stream)."""
    rotation = [None] * len(training_class_distribution)
    for i in range(len(training_class_distribution)):
        rotation[i] = np.random.randint(0, 360, size=len(training_class_distribution[i]))
    if verbose:
        for i in range(len(training_class_distribution)):
            print(BatchColors.OKBLUE + 'Class ' + str(i + 1) + ' has length ' + str(len(training_class_distribution[i])) +
                  ' and rotation length ' + str(len(rotation[i])) + BatchColors.ENDC)
    return rotation


def select_super_batch_instances(class_distribution, rotation_distribution=None, batch_size=100, super_batch=500):
    """isprs:403-445: class-balanced draw of batch_size*super_batch (map, x, y, rot) instances."""
    instances = []
    overall_count = 0
    samples_per_class = int((batch_size * super_batch) / len(class_distribution))
    for i in range(len(class_distribution)):
        n = len(class_distribution[i])
        shuffle = np.asarray(random.sample(range(n), (samples_per_class if n >= samples_per_class else n)))
        for j in shuffle:
            cur_map, cur_x, cur_y = class_distribution[i][j][0], class_distribution[i][j][1], class_distribution[i][j][2]
            cur_rot = (rotation_distribution[i][j] if (rotation_distribution is not None) else 0)
            instances.append((cur_map, cur_x, cur_y, cur_rot))
            overall_count += 1
    if overall_count != (batch_size * super_batch):
        lack = (batch_size * super_batch) - overall_count
        for i in range(lack):
            rand_class = np.random.randint(len(class_distribution))
            rand_map = np.random.randint(len(class_distribution[rand_class]))
            cur_map = class_distribution[rand_class][rand_map][0]
            cur_x = class_distribution[rand_class][rand_map][1]
            cur_y = class_distribution[rand_class][rand_map][2]
            cur_rot = (rotation_distribution[rand_class][rand_map] if (rotation_distribution is not None) else 0)
            instances.append((cur_map, cur_x, cur_y, cur_rot))
            overall_count += 1
    assert overall_count == (batch_size * super_batch), "Could not select ALL instances"
    return np.asarray(instances)


def compute_image_mean(data):
    """isprs:84-88: mean over everything per channel; std across patches at pixel (0,0) only."""
    mean_full = np.mean(np.mean(np.mean(data, axis=0), axis=0), axis=0)
    std_full = np.std(data, axis=0, ddof=1)[0, 0, :]
    return mean_full, std_full


def dynamically_calculate_mean_and_std(data, indexes, crop_size):
    """isprs:151-184: chunked (5000 patches) mean/std over all reference-crop windows."""
    total = []
    for cls in indexes:
        total = total + list(cls)
    mean_full, std_full, all_patches = [], [], []
    for i in range(len(total)):
        cur_map, cur_x, cur_y = total[i][0], total[i][1], total[i][2]
        all_patches.append(data[cur_map][cur_x:cur_x + crop_size, cur_y:cur_y + crop_size, :])
        if i > 0 and i % 5000 == 0:
            mean, std = compute_image_mean(np.asarray(all_patches))
            mean_full.append(mean)
            std_full.append(std)
            all_patches = []
    mean, std = compute_image_mean(np.asarray(all_patches))
    mean_full.append(mean)
    std_full.append(std)
    return np.mean(mean_full, axis=0), np.mean(std_full, axis=0)


# ------------------------------------------------------------------------------------------------
# per-step augmentation plan (decisions on the host, pixels on the GPU)
# ------------------------------------------------------------------------------------------------
class BatchPlan:
    """What ``dynamically_create_patches`` decided for one batch; consumed by Session.gather_dev."""
    __slots__ = ("inst", "flips", "noise", "noise_on", "over_x", "over_y", "over_on", "acc_mask", "crop", "rot", "rot_on",
                 "noise_slot", "slot", "local")

    def __init__(self):
        self.rot = self.rot_on = None
        self.local = False          # True: the plan already holds only this rank's patches (data-parallel local planning)
        self.noise_slot = None      # compact noise: block noise_slot[b] of ``noise`` belongs to patch b (-1: none)
        self.slot = None            # the PlanSlot whose (pinned) buffers the arrays above are views of


def rotate_affine(angle, crop_size):
    """The affine map scipy.ndimage.rotate(x, angle, reshape=False) hands to its C loop for a crop x crop plane:
    [m00, m01, m10, m11, off0, off1] (same functions, same operation order as scipy's Python front end), consumed by the
    gather kernel (drs_gather_rot_dev)."""
    from scipy import special
    c, s = special.cosdg(angle), special.sindg(angle)
    m = np.array([[c, s], [-s, c]])
    shp = np.asarray([crop_size, crop_size])
    out_center = m @ ((shp - 1) / 2)
    in_center = (shp - 1) / 2
    off = in_center - out_center
    return np.array([m[0, 0], m[0, 1], m[1, 0], m[1, 1], off[0], off[1]], dtype=np.float64)


_ROT_TABLES = {}


def rotation_table(crop_size):
    """[360, 6] rotate_affine of every integer angle create_rotation_distribution can draw (isprs:489), built once per patch
    size with the scalar code path above (so the values are the ones scipy itself would compute) and then only indexed."""
    t = _ROT_TABLES.get(crop_size)
    if t is None:
        t = np.stack([rotate_affine(a, crop_size) for a in range(360)])
        _ROT_TABLES[crop_size] = t
    return t


def plan_isprs_batch(data, mask_data, training_instances_batch, crop_size, is_train=True, rotate_on_device=False):
    """isprs:245-334.  Consumes np.random exactly like the reference, per patch:
    randint(0,2) rotate?  randint(0,2) noise? (+ normal(0,0.01,shape) if yes)  randint(0,3) flip.
    Rotated patches (scipy nearest-neighbour, isprs:294-296): with rotate_on_device the plan only carries the affine map
    per patch and the gather kernel rotates patch, labels and accuracy mask (acc_mask stays None: it is produced on the
    device); otherwise they are produced here with scipy and handed to the gather as overrides.  Every other patch is
    cut from the HBM-resident scene by the gather kernel."""
    import scipy.ndimage
    B = len(training_instances_batch)
    C = data[0].shape[-1]
    p = BatchPlan()
    p.crop = crop_size
    p.inst = np.zeros((B, 3), dtype=np.int32)
    p.flips = np.zeros(B, dtype=np.uint8)
    p.noise_on = np.zeros(B, dtype=np.uint8)
    p.over_on = np.zeros(B, dtype=np.uint8)
    p.noise = None
    p.over_x = None
    p.over_y = None
    p.acc_mask = None
    for i in range(B):
        cur_map = int(training_instances_batch[i][0])
        h, w = data[cur_map].shape[0], data[cur_map].shape[1]
        cur_x, cur_y = shift_back(int(training_instances_batch[i][1]), int(training_instances_batch[i][2]), crop_size, h, w)
        if cur_x < 0 or cur_y < 0 or cur_x + crop_size > h or cur_y + crop_size > w:
            raise ValueError(BatchColors.FAIL + "Error: Current PATCH size is " + str(min(h, crop_size)) + "x" +
                             str(min(w, crop_size)) + BatchColors.ENDC)
        p.inst[i] = (cur_map, cur_x, cur_y)
        if not is_train:
            continue
        cur_rot = training_instances_batch[i][3]
        possible_rotation = np.random.randint(0, 2)
        if possible_rotation == 1 and rotate_on_device:
            if p.rot is None:
                p.rot = np.zeros((B, 6), dtype=np.float64)
                p.rot_on = np.zeros(B, dtype=np.uint8)
            a = int(cur_rot)
            p.rot[i] = rotation_table(crop_size)[a] if (a == cur_rot and 0 <= a < 360) else rotate_affine(cur_rot, crop_size)
            p.rot_on[i] = 1
        elif possible_rotation == 1:
            if p.over_x is None:
                p.over_x = np.zeros((B, crop_size, crop_size, C), dtype=np.float64)
                p.over_y = np.zeros((B, crop_size, crop_size), dtype=np.uint8)
                p.acc_mask = np.ones((B, crop_size, crop_size), dtype=np.uint8)
            cur_patch = data[cur_map][cur_x:cur_x + crop_size, cur_y:cur_y + crop_size, :]
            cur_mask_patch = mask_data[cur_map][cur_x:cur_x + crop_size, cur_y:cur_y + crop_size]
            p.over_x[i] = scipy.ndimage.rotate(cur_patch, cur_rot, order=0, reshape=False)
            p.over_y[i] = scipy.ndimage.rotate(cur_mask_patch, cur_rot, order=0, reshape=False)
            p.acc_mask[i] = scipy.ndimage.rotate(np.ones((crop_size, crop_size), dtype=bool), cur_rot, order=0,
                                                 reshape=False)
            p.over_on[i] = 1
        possible_noise = np.random.randint(0, 2)
        if possible_noise == 1:
            if p.noise is None:
                p.noise = np.zeros((B, crop_size, crop_size, C), dtype=np.float64)
            p.noise[i] = np.random.normal(0, 0.01, (crop_size, crop_size, C))
            p.noise_on[i] = 1
        possible_flip = np.random.randint(0, 3)
        p.flips[i] = (FLIP_NONE, FLIP_UD, FLIP_LR)[possible_flip]     # isprs:304-318
    # the accuracy mask is flipped together with the patch (isprs:308-317)
    if p.acc_mask is not None:
        for i in range(B):
            if p.flips[i] == FLIP_UD:
                p.acc_mask[i] = np.flipud(p.acc_mask[i])
            elif p.flips[i] == FLIP_LR:
                p.acc_mask[i] = np.fliplr(p.acc_mask[i])
    return p


class PlanSlot:
    """Reusable buffers for one native plan; page-locked when a CUDA device is present so that the gather's uploads are
    asynchronous copies straight from here (drs_gather_plan_dev)."""

    def __init__(self, batch_max, crop_max, channels, pinned=None):
        self.B, self.crop_max, self.C = int(batch_max), int(crop_max), int(channels)
        n_noise = self.B * self.crop_max * self.crop_max * self.C + 2

        def buf(n, dtype):
            if pinned is None:
                return np.zeros(n, dtype=dtype)
            import torch
            t = torch.zeros(n, dtype=getattr(torch, np.dtype(dtype).name)).pin_memory()
            self._keep.append(t)
            return t.numpy()

        self._keep = []
        self.inst = buf(self.B * 3, np.int32).reshape(self.B, 3)
        self.flips = buf(self.B, np.uint8)
        self.rot_on = buf(self.B, np.uint8)
        self.rot = buf(self.B * 6, np.float64).reshape(self.B, 6)
        self.noise_on = buf(self.B, np.uint8)
        self.noise_slot = buf(self.B, np.int32)
        self.noise = buf(n_noise, np.float64)


class NativePlanner:
    """``plan_isprs_batch`` (isprs:245-334) through the library's host planner (csrc/host_plan.cpp).

    Same decisions, same consumption of the global ``np.random`` stream and bit-identical noise values as the Python
    function above (tests/test_host_plan.py), at ~1/10 of the cost: the state of the legacy global generator is handed to
    the C code and written back afterwards.  Rotation always runs on the device with this planner."""

    def __init__(self, threads=None):
        import ctypes as C
        import os
        from . import lib as L
        self._C, self._L = C, L
        self._lib = L.load()
        self._h = C.c_void_p()
        if self._lib.drs_planner_create(C.byref(self._h)) != 0:
            raise L.DrsError("drs_planner_create failed")
        self.threads = int(threads if threads is not None else os.environ.get("DRS_PLAN_THREADS", min(4, os.cpu_count() or 1)))
        self._ms = L.MtState()
        self._key_view = np.frombuffer(self._ms.key, dtype=np.uint32)

    def close(self):
        if self._h:
            self._lib.drs_planner_destroy(self._h)
            self._h = self._C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _load_state(self):
        st = np.random.get_state()
        self._key_view[:] = st[1]
        self._ms.pos, self._ms.has_gauss, self._ms.gauss = int(st[2]), int(st[3]), float(st[4])

    def _store_state(self):
        np.random.set_state(('MT19937', self._key_view.copy(), int(self._ms.pos), int(self._ms.has_gauss), float(self._ms.gauss)))

    def normal(self, loc, scale, n):
        """np.random.normal(loc, scale, n) on the global stream (test entry)."""
        out = np.empty(int(n), dtype=np.float64)
        self._load_state()
        rc = self._lib.drs_mt_normal(self._h, self._C.byref(self._ms), float(loc), float(scale), out.ctypes.data, int(n), self.threads)
        if rc:
            raise self._L.DrsError("drs_mt_normal failed (%d)" % rc)
        self._store_state()
        return out

    def randint(self, n, count):
        out = np.empty(int(count), dtype=np.int32)
        self._load_state()
        rc = self._lib.drs_mt_randint(self._C.byref(self._ms), int(n), out.ctypes.data, int(count))
        if rc:
            raise self._L.DrsError("drs_mt_randint failed (%d)" % rc)
        self._store_state()
        return out

    def local_stream(self, seed):
        """A generator state of its own (not the global ``np.random`` one) for data-parallel local planning: every rank draws
        the augmentation of ITS patches from its own stream, so no rank has to scan the noise of the whole global batch."""
        st = np.random.RandomState(int(seed) & 0xffffffff).get_state()
        ms = self._L.MtState()
        np.frombuffer(ms.key, dtype=np.uint32)[:] = st[1]
        ms.pos, ms.has_gauss, ms.gauss = int(st[2]), int(st[3]), float(st[4])
        return ms

    def plan(self, scene_hw, training_instances_batch, crop_size, channels, slot=None, own=None, stream=None):
        """-> BatchPlan with rotate_on_device semantics.  scene_hw: int32 [n_scenes, 2]; own = (b0, b1): the patches whose
        noise values are needed (data-parallel rank slice), default all.  stream: an MtState from local_stream() to draw from
        instead of the global generator."""
        C = self._C
        inst_in = np.ascontiguousarray(training_instances_batch, dtype=np.int64)
        B = inst_in.shape[0]
        if inst_in.ndim != 2 or inst_in.shape[1] != 4:
            raise ValueError("training instances must be (map, x, y, rotation) rows")
        if slot is None or slot.B < B or slot.crop_max < crop_size or slot.C != channels:
            slot = PlanSlot(B, crop_size, channels)
        b0, b1 = (0, B) if own is None else own
        n_noise = C.c_int32()
        ms = stream if stream is not None else self._ms
        if stream is None:
            self._load_state()
        rc = self._lib.drs_plan_isprs_batch(self._h, C.byref(ms), inst_in.ctypes.data, B, scene_hw.ctypes.data, scene_hw.shape[0],
                                            int(crop_size), int(channels), 1, rotation_table(int(crop_size)).ctypes.data,
                                            slot.inst.ctypes.data, slot.flips.ctypes.data, slot.rot_on.ctypes.data, slot.rot.ctypes.data,
                                            slot.noise_on.ctypes.data, slot.noise_slot.ctypes.data, slot.noise.ctypes.data,
                                            slot.noise.size, C.byref(n_noise), self.threads, int(b0), int(b1))
        if rc == 3:
            raise ValueError(BatchColors.FAIL + "Error: Current PATCH size is out of the scene" + BatchColors.ENDC)
        if rc:
            raise self._L.DrsError("drs_plan_isprs_batch failed (%d)" % rc)
        if stream is None:
            self._store_state()
        p = BatchPlan()
        p.crop = int(crop_size)
        p.inst, p.flips = slot.inst[:B], slot.flips[:B]
        p.rot, p.rot_on = slot.rot[:B], slot.rot_on[:B]
        p.noise_on, p.noise_slot = slot.noise_on[:B], slot.noise_slot[:B]
        nn = int(n_noise.value)
        p.noise = slot.noise[:nn * crop_size * crop_size * channels] if nn else None
        p.over_x = p.over_y = p.over_on = p.acc_mask = None
        p.slot = slot
        return p


def plan_index_flip_batch(class_distribution, shuffle_batch, crop_size, shapes, with_map=False):
    """contest:192-254 / coffee:241-293: flip encoded by the shuffle index range
    [0,N) none, [N,2N) fliplr, [2N,3N) flipud; window start shifted back at the border (negative-offset slices)."""
    n = len(class_distribution)
    B = len(shuffle_batch)
    p = BatchPlan()
    p.crop = crop_size
    p.inst = np.zeros((B, 3), dtype=np.int32)
    p.flips = np.zeros(B, dtype=np.uint8)
    p.noise = p.noise_on = p.over_x = p.over_y = p.over_on = p.acc_mask = None
    for b, i in enumerate(shuffle_batch):
        i = int(i)
        if i >= 2 * n:
            cur_pos, flip = i - 2 * n, FLIP_UD
        elif i >= n:
            cur_pos, flip = i - n, FLIP_LR
        else:
            cur_pos, flip = i, FLIP_NONE
        if with_map:      # coffee: (map, (x, y))
            cur_map = int(class_distribution[cur_pos][0])
            cur_x, cur_y = int(class_distribution[cur_pos][1][0]), int(class_distribution[cur_pos][1][1])
        else:             # contest: (x, y), single scene
            cur_map = 0
            cur_x, cur_y = int(class_distribution[cur_pos][0]), int(class_distribution[cur_pos][1])
        h, w = shapes[cur_map]
        cur_x, cur_y = shift_back(cur_x, cur_y, crop_size, h, w)
        p.inst[b] = (cur_map, cur_x, cur_y)
        p.flips[b] = flip
    return p


def sliding_stride(crop_size):
    return int(math.floor(crop_size / 2.0))       # isprs:1243


# ------------------------------------------------------------------------------------------------
# contest / coffee one-off set-up (C13/C14): reference semantics incl. their quirks
# ------------------------------------------------------------------------------------------------
def contest_create_distributions_over_classes(labels, crop_size, stride_crop, num_classes=7, verbose=True):
    """contest:172-189.  ``count`` has length max_label+1, so ``count[-1]`` is the highest class PRESENT in the window
    (not label 7): single-class windows are skipped and the highest class present never wins; partial border windows
    are dropped (no shift-back)."""
    classes = [[] for _ in range(num_classes)]
    w, h = labels.shape
    for i in range(0, w, stride_crop):
        for j in range(0, h, stride_crop):
            patch_class = labels[i:i + crop_size, j:j + crop_size]
            if patch_class.shape == (crop_size, crop_size):
                count = np.bincount(patch_class.astype(int).flatten())
                if count[-1] == crop_size * crop_size:
                    continue
                classes[int(np.argmax(count[:-1]))].append((i, j))
    if verbose:
        for i in range(len(classes)):
            print("Class " + str(i + 1) + " has " + str(len(classes[i])) + " instances")
    out = []
    for c in classes:
        out += c
    return out


def contest_create_mean_and_std(data, class_distribution, crop_size):
    """contest:104-116."""
    all_patches = [data[x:x + crop_size, y:y + crop_size, :] for (x, y) in class_distribution]
    return compute_image_mean(np.asarray(all_patches))


def coffee_create_distributions_over_classes(labels, crop_size, stride_crop, num_classes=2):
    """coffee:358-372: (tile, (row, col)) of every full window, bucketed by majority class."""
    classes = [[] for _ in range(num_classes)]
    for k in range(len(labels)):
        lab = np.asarray(labels[k])
        w, h = lab.shape[0], lab.shape[1]
        for i in range(0, w, stride_crop):
            for j in range(0, h, stride_crop):
                patch_class = np.squeeze(lab[i:i + crop_size, j:j + crop_size])
                if patch_class.shape == (crop_size, crop_size):
                    count = np.bincount(patch_class.astype(int).flatten())
                    classes[int(np.argmax(count))].append((k, (i, j)))
    out = []
    for c in classes:
        out += c
    return out


def coffee_create_mean_and_std(training_data, crop_size, stride_crop):
    """coffee:352-355 over create_crops_stride(..., is_train=True) (coffee:176-238): every full window with the
    alternating stride/stride+1 walk of odd crop sizes, plus its two mirrored copies (which leave mean/std-at-(0,0)
    dependent on the flips, so they are generated like the reference does)."""
    crops = []
    for i in range(len(training_data)):
        tile = training_data[i]
        j, count_x = 0, 0
        while j < tile.shape[0]:
            k, count_y = 0, 0
            while k < tile.shape[1]:
                if j + crop_size <= tile.shape[0] and k + crop_size <= tile.shape[1]:
                    crop = tile[j:j + crop_size, k:k + crop_size, :]
                    crops.append(crop)
                    crops.append(np.fliplr(crop))
                    crops.append(np.flipud(crop))
                k += (stride_crop + 1) if (crop_size % 2 != 0 and count_y % 2 != 0) else stride_crop
                count_y += 1
            j += (stride_crop + 1) if (crop_size % 2 != 0 and count_x % 2 != 0) else stride_crop
            count_x += 1
    return compute_image_mean(np.asarray(crops))
