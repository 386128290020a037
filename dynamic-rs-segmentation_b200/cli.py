"""Shared pieces of the three drop-in command lines (positional argv exactly as the reference)."""
import os
import sys

import numpy as np

from .host import BatchColors


def print_params(list_params):       # isprs:31-35
    print('\n------------------------------------------------')
    for i in range(1, len(sys.argv)):
        print(list_params[i - 1] + '= ' + sys.argv[i])
    print('------------------------------------------------\n')


def dist_setup():
    """One process per GPU under torchrun (RANK / LOCAL_RANK / WORLD_SIZE); single process otherwise."""
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def rank0_first(fn):
    """Cache files in the working directory (isprs:2087-2115): under torchrun rank 0 runs ``fn`` (which may create and save),
    the others wait at a barrier and run it afterwards (finding the files and loading them).  Writers save to a temporary
    name and os.replace it, so a reader never sees a half-written file."""
    _, world, _ = dist_setup()
    if world <= 1:
        return fn()
    from . import dist as ddist
    return ddist.rank0_first(fn)


def save_atomic(path, array, **kw):
    tmp = path + '.tmp.npy'
    np.save(tmp, array, **kw)
    os.replace(tmp, path)


def load_npy_scenes(path, instances, prefix=""):
    """Scenes as ``<path>/<prefix><instance>_image.npy`` ([H,W,C] float64/float32) + ``..._labels.npy`` ([H,W] uint8).

    The reference reads the ISPRS tif/jpg files with scipy.misc / GDAL (isprs:187-242) and Torch7-ASCII dumps for
    contest/coffee (contest:146-169, coffee:84-125); dataset I/O is outside the hot path and neither the
    data nor those readers exist in this image, so the drop-in reads pre-converted ``.npy`` arrays of the same dtype."""
    data, labels = [], []
    for inst in instances:
        img = os.path.join(path, "%s%s_image.npy" % (prefix, inst))
        lab = os.path.join(path, "%s%s_labels.npy" % (prefix, inst))
        if not os.path.isfile(img):
            print(BatchColors.FAIL + "Error! Scene file not found: " + img + BatchColors.ENDC)
            raise FileNotFoundError(img)
        data.append(np.load(img))
        labels.append(np.load(lab).astype(np.uint8) if os.path.isfile(lab) else np.zeros(data[-1].shape[:2], dtype=np.uint8))
    return data, labels


def make_backend(net_type, channels, num_classes, weight_decay, lr_initial, decay_rate, scenes, label_maps, mean_full, std_full,
                 isprs_scopes, training, seed=None, train_fp16_patches=False):
    """Session (libdrs.so) + GpuBackend with every scene resident in HBM.  Training runs the bf16 tensor-core path, whole-
    scene inference the fp16 one (10-bit mantissa, the TF32 class); DRS_PRECISION=fp32 selects the exact-order mode."""
    import drs_b200
    from .backend import GpuBackend
    from . import dist as ddist
    rank, world, local = dist_setup()
    prec = os.environ.get("DRS_PRECISION", "bf16" if training else "f16")
    s = drs_b200.Session(net_type, channels, num_classes, weight_decay=weight_decay, lr_initial=lr_initial, decay_rate=decay_rate,
                         precision=prec, device=local, isprs_scopes=isprs_scopes,
                         seed=int(os.environ.get("DRS_SEED", "0")) if seed is None else seed)
    be = GpuBackend(s, scenes, label_maps, mean_full, std_full, device=local, rank=rank, world=world,
                    train_fp16_patches=train_fp16_patches)
    if world > 1:
        sync_bn = bool(int(os.environ.get("DRS_SYNC_BN", "0")))
        if os.environ.get("DRS_COMM", "nccl") == "torch":
            ddist.attach_allreduce(s, sync_bn=sync_bn)       # exchange through a torch.distributed callback
        else:
            ddist.attach_nccl(s, sync_bn=sync_bn)            # the library's own NCCL communicator (csrc/drs_comm.cuh)
    return be
