#!/usr/bin/env python
"""bench.py -- the hot path's headline metrics on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|infer]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload = BASELINE.json configs[1]: dilated_grsl (Dilated6Pooling), multinomial patch sizes
{25,29,...,49}, update_type acc, batch 64 per GPU, Vaihingen-shaped synthetic scene resident in HBM.
One "step" = the reference's training iteration (isprs:1726-1763): draw a patch size, select the batch, gather +
normalise the patches, sess.run([optimizer, loss, pred_up]) (forward, backward, exchange, momentum update),
per-crop confusion matrix, score update.  metric = train patches/s (whole job).
The same JSON line carries the second headline metric as "inference": full-scene sliding-window inference of a
Potsdam-shaped 6000x6000x5 scene with dilated_grsl_rate8 (configs[3]) in Mpixel/s.

--impl reference times the reference-equivalent CPU path (TensorFlow is not installable here: the oracle's
PyTorch-CPU restatement of the graph + the NumPy restatement of the host loops) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRAIN_CFG = dict(net="dilated_grsl", dataset="vaihingen", C=4, K=6, batch=64, values=[25, 29, 33, 37, 41, 45, 49],
                 distribution="multinomial", update="acc", lr=0.01, wd=0.005)
INFER_CFG = dict(net="dilated_grsl_rate8", dataset="potsdam", C=5, K=6, crop=25, batch=64, H=6000, W=6000)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback")


class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region (B200_PROFILING.md) through NVML: a burst of samples
    10 ms apart for short regions, then one every 200 ms (NVML calls take a driver lock that kernel launches also need;
    polling every 5 ms slowed the launch-heavy scene pass by 50 %).  Falls back to one nvidia-smi query."""

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread, self.nv = index, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.01 if len(self.samples) < 8 else 0.2)

    def stop(self):
        if self.nv is None:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,power.draw",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout.split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "power_w_max": float(out[2]), "samples": 1,
                        "reasons": [], "note": "single nvidia-smi sample after the timed region (NVML unavailable)"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        nv = self.nv
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        reasons = set()
        for _, _, r in self.samples:
            for k, bit in names.items():
                if r & bit:
                    reasons.add(k)
        sm = [s[0] for s in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "power_w_max": max([s[1] for s in self.samples]) if self.samples else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def profiled_traffic(which):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        return float(json.load(open(p))[which]["dram_bytes_per_launch"])
    except Exception:
        return None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ----------------------------------------------------------------------------------------------------------------
# shared host-side step logic (the reference's policy, seeded identically on every rank)
# ----------------------------------------------------------------------------------------------------------------
class TrainHost:
    """Patch-size draws, batch selection and flips of the training loop (isprs:1726-1763), seeded."""

    def __init__(self, cfg, shapes, global_batch, seed=77):
        import random
        from drs_b200 import host, synth
        self.host, self.cfg, self.shapes, self.gb = host, cfg, shapes, global_batch
        np.random.seed(seed)
        random.seed(seed)
        self.values = cfg["values"]
        self.probs = host.define_multinomial_probs(self.values)
        self.pal, self.occ, self.chosen = host.init_score_arrays(cfg["distribution"], self.values)
        self.instances = synth.random_instances(np.random.RandomState(seed), global_batch * 100, shapes)
        self.total = len(self.instances)
        self.shuffle = np.asarray(random.sample(range(self.total), self.total))
        self.it = 0
        self.shape_arr = np.asarray(shapes, dtype=np.int64)
        self.rs_flips = np.random.RandomState(seed + 1)   # own stream: the patch-size sequence must not depend on the batch size

    def next_plan(self):
        host = self.host
        # the multinomial spreads 44 % of its mass over the unlisted sizes of [25, 49], exactly like the reference
        crop, idx = host.draw_patch_size(self.cfg["distribution"], self.values, self.probs)
        self.shuffle, batch, self.it = host.select_batch(self.shuffle, self.gb, self.it, self.total)
        # border rule of every gather in the reference (isprs:259-269), vectorised: a window that sticks out is moved back
        sel = self.instances[batch]
        hw = self.shape_arr[sel[:, 0]]
        inst = np.empty((len(batch), 3), dtype=np.int32)
        inst[:, 0] = sel[:, 0]
        inst[:, 1] = np.minimum(sel[:, 1], hw[:, 0] - int(crop))
        inst[:, 2] = np.minimum(sel[:, 2], hw[:, 1] - int(crop))
        flips = self.rs_flips.randint(0, 3, size=len(batch)).astype(np.uint8)  # isprs:304 flip decision per patch
        # isprs:289-296 rotation decision per patch; the nearest-neighbour rotation itself runs in the gather kernel
        rot_on = self.rs_flips.randint(0, 2, size=len(batch)).astype(np.uint8)
        rot = host.rotation_table(int(crop))[sel[:, 3] % 360]
        self.angles = sel[:, 3] % 360
        return int(crop), idx, inst, flips, rot, rot_on

    def update(self, idx, loss, cm):
        acc_norm = self.host.acc_norm_from_cm(cm, self.cfg["K"])
        self.host.update_scores(self.pal, self.occ, idx, self.cfg["update"], loss, acc_norm)


def conv_flops_train(net, C, K, M):
    """Algorithmic FLOPs of one training step: fprop + dgrad + wgrad of every conv, minus conv1's dgrad."""
    from drs_b200 import nets
    plan, cls_in = nets.layer_plan(net, C)
    macs = 0
    for i, (_, k, r, ci, co) in enumerate(plan):
        macs += k * k * ci * co * (3 if i > 0 else 2)
    macs += 3 * cls_in * K
    return 2.0 * macs * M


# ----------------------------------------------------------------------------------------------------------------
# ours
# ----------------------------------------------------------------------------------------------------------------
def run_train_ours(args, rank, world, local):
    import torch
    import torch.distributed as dist
    import drs_b200
    from drs_b200 import dist as ddist, nets, synth
    cfg = TRAIN_CFG
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    img, lab = synth.scene(cfg["dataset"])
    mean, std = synth.normalisation(img)
    s = drs_b200.Session(cfg["net"], cfg["C"], cfg["K"], weight_decay=cfg["wd"], lr_initial=cfg["lr"], precision="bf16",
                         device=local, seed=5)
    s.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    s.upload_scene(0, img, lab)
    s.set_normalization(mean, std)
    if world > 1:
        ddist.attach_allreduce(s, sync_bn=False)
    B = cfg["batch"]
    s.reserve(B, cfg["values"][-1], training=True)      # the patch-size interval is known up front (probValues)
    th = TrainHost(cfg, [img.shape[:2]], B * world)
    cmax = max(cfg["values"][-1], 49)
    x = torch.empty(B * cmax * cmax * cfg["C"], dtype=torch.float32, device=dev)
    y = torch.empty(B * cmax * cmax, dtype=torch.float32, device=dev)
    pred = torch.empty(B * cmax * cmax, dtype=torch.uint8, device=dev)
    cm_dev = torch.zeros(cfg["K"] ** 2 + 1, dtype=torch.int32, device=dev)
    amask = torch.empty(B * cmax * cmax, dtype=torch.uint8, device=dev)
    from drs_b200 import host as _host
    for c in range(cfg["values"][0], cfg["values"][-1] + 1):
        _host.rotation_table(c)                          # affine maps of the 360 angles per patch size of the interval

    def step():
        crop, idx, inst, flips, rot, rot_on = th.next_plan()
        sl = slice(rank * B, (rank + 1) * B)
        s.gather_rot_dev(inst[sl], flips[sl], crop, x, y, rot=rot[sl], rot_on=rot_on[sl], amask_out_dev=amask)
        loss = s.train_step_dev(x, y, B, crop, pred_dev=pred, cm_dev=cm_dev, acc_mask_dev=amask)
        cm = cm_dev.cpu().numpy()[:cfg["K"] ** 2].reshape(cfg["K"], cfg["K"])
        th.update(idx, float(loss), cm)
        return crop

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    host_state = (np.random.get_state(), th.rs_flips.get_state(), th.shuffle.copy(), th.it)
    import random as _random
    py_state = _random.getstate()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = s.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    px = 0
    crops = []
    for _ in range(args.steps):
        c = step()
        crops.append(c)
        px += B * c * c
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = s.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = B * world * args.steps / (ms / 1e3)

    # ---- kernel timing pass: the SAME K steps again (host policy rewound to the same patch sizes and batches) with a CUDA
    # event pair around every tensor-core launch.  Timing single launches needs them serialised, so this pass runs the
    # filter gradients on the main stream instead of overlapping them with the backward's HBM-bound kernels (which is what
    # the timed region above does); its step time is reported next to the kernel time.
    np.random.set_state(host_state[0]); th.rs_flips.set_state(host_state[1]); th.shuffle = host_state[2]; th.it = host_state[3]
    _random.setstate(py_state)
    s.set_profiling(True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_prof = e0.elapsed_time(e1)
    conv_ms, conv_n, conv_fl = s.profile_read()
    s.set_profiling(False)

    # ---- end to end through the sess.run seam: host x / y in pinned memory, host loss / pred / confusion back
    e2e = None
    host_batches = {}
    rs = np.random.RandomState(3)
    for c in sorted(set(crops)):
        xh = torch.from_numpy(rs.randn(B, c * c * cfg["C"]).astype(np.float32)).pin_memory()
        yh = torch.from_numpy(rs.randint(0, cfg["K"], size=(B, c * c)).astype(np.float32)).pin_memory()
        host_batches[c] = (xh.numpy(), yh.numpy(), xh, yh)
    for c in crops[:max(1, min(3, len(crops)))]:
        s.train_step(host_batches[c][0], host_batches[c][1], c, want_cm=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    bi = bo = 0
    dbg = []
    for c in crops:
        t1 = time.perf_counter()
        s.train_step(host_batches[c][0], host_batches[c][1], c, want_cm=True)
        dbg.append(round((time.perf_counter() - t1) * 1e3, 1))
        bi += B * c * c * (cfg["C"] + 1) * 4
        bo += B * c * c * 8 + 4 + (cfg["K"] ** 2 + 1) * 4
    e1.record()
    barrier()
    ms2 = e0.elapsed_time(e1)
    if os.environ.get("BENCH_DEBUG") and rank == 0:
        print("e2e per-step ms:", dbg, file=sys.stderr)
    if world > 1:
        t = torch.tensor([ms2], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2 = float(t.item())
    e2e = {"value": B * world * args.steps / (ms2 / 1e3), "unit": "patches/s", "h2d_bytes_per_step": bi // len(crops),
           "d2h_bytes_per_step": bo // len(crops), "api": "Session.train_step == drs_train_step_host (sess.run seam, host feeds/fetches)"}
    s.close()

    pk = peaks()
    M_total = px
    roof = {"bound": "tensor", "kernel": "tcgen05 kernels of the step: conv_tc_kernel (fprop + dgrad) and wgrad_tc_kernel",
            "achieved": (conv_fl / (conv_ms * 1e-3) / 1e12) if conv_ms > 0 else None, "peak": pk["tc_sustained"],
            "unit": "TFLOP/s", "frac": (conv_fl / (conv_ms * 1e-3) / 1e12 / pk["tc_sustained"]) if conv_ms > 0 else None,
            "traffic": profiled_traffic("train"), "peak_source": pk["source"] + " bf16 sustained (kernel timed inside a long step)",
            "launches": conv_n, "kernel_ms_per_step": conv_ms / args.steps, "ms_per_step_timing_pass": ms_prof / args.steps,
            "step_tflops_all_kernels": conv_flops_train(cfg["net"], cfg["C"], cfg["K"], M_total) / (ms * 1e-3) / 1e12}
    return dict(metric="train patches/s", value=value, unit="patches/s", ms_per_step=ms / args.steps, dtype="bf16",
                scaling="weak", e2e=e2e, gpu_launches=launches, clocks=clocks, roofline=roof,
                config={"workload": "configs[1]: dilated_grsl multinomial {25..49} acc, batch 64/GPU, Vaihingen-shaped 2000x2500x4 "
                                    "float64 scene resident in HBM", "net": cfg["net"], "batch_per_gpu": B, "global_batch": B * world,
                        "patch_sizes_drawn": crops, "parallelism": "dp%d" % world, "sync_bn": False,
                        "l2": "per-step working set (activations %.0f-%.0f MB) exceeds the 126 MB L2; no explicit flush" %
                              (B * 25 * 25 * 896 * 6 / 1e6, B * 49 * 49 * 896 * 6 / 1e6),
                        "augment": "flips and nearest-neighbour rotation (scipy order 0, SURVEY N1) in the gather kernel, decisions drawn on the "
                                   "host; the additive np.random.normal noise (host RNG stream) is not in the timed step"})


def run_infer_ours(args, rank, world, local, steps=1):
    import torch
    import torch.distributed as dist
    import drs_b200
    from drs_b200 import dist as ddist, nets, synth
    cfg = INFER_CFG
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    H, W = cfg["H"], cfg["W"]
    if args.small:
        H = W = 1500
    img, _ = synth.scene(cfg["dataset"], H=H, W=W)
    mean, std = synth.normalisation(img)
    s = drs_b200.Session(cfg["net"], cfg["C"], cfg["K"], precision="f16", device=local, seed=9)
    s.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    s.set_normalization(mean, std)
    r0, r1 = ddist.stripe_bounds(H, world, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # a rank keeps only the scene rows its stripe needs (the stripe plus the patch rows straddling its borders)
    u0, u1 = ddist.stripe_rows_needed(H, cfg["crop"], r0, r1) if world > 1 else (None, None)
    s.upload_scene(0, img, None, u0, u1)
    s.scene_infer(0, cfg["crop"], cfg["batch"], H, W, row_begin=r0, row_end=min(r1, r0 + 40))   # warm-up stripe (kernels)
    s.scene_infer(0, cfg["crop"], cfg["batch"], H, W, row_begin=r0, row_end=r1)   # warm-up pass: stripe-sized buffers allocated
    barrier()
    l0 = s.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0.record()
    for _ in range(steps):
        stripe = s.scene_infer(0, cfg["crop"], cfg["batch"], H, W, row_begin=r0, row_end=r1)
        full = ddist.gather_label_stripes(stripe, H, W, rank, world, device=dev)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    launches = s.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    # kernel timing pass (roofline): the same pass once more with a CUDA-event pair around every tensor-core launch
    s.set_profiling(True)
    s.scene_infer(0, cfg["crop"], cfg["batch"], H, W, row_begin=r0, row_end=r1)
    barrier()
    conv_ms, conv_n, conv_fl = s.profile_read()
    conv_ms, conv_n, conv_fl = conv_ms * steps, conv_n * steps, conv_fl * steps      # (reported per step below)
    s.set_profiling(False)
    # end to end: host scene -> HBM -> label map on the host
    barrier()
    e0.record()
    if world == 1:
        # a fresh tile: the upload is streamed ahead of the chunks that read it (drs_scene_infer_host)
        stripe = s.scene_infer_host(0, img, cfg["crop"], cfg["batch"])
    else:
        s.upload_scene(0, img, None, u0, u1)
        stripe = s.scene_infer(0, cfg["crop"], cfg["batch"], H, W, row_begin=r0, row_end=r1)
    full = ddist.gather_label_stripes(stripe, H, W, rank, world, device=dev)
    e1.record()
    barrier()
    ms2 = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms, ms2], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms2 = float(t[0].item()), float(t[1].item())
    s.close()
    pk = peaks()
    ach = (conv_fl / steps / (conv_ms / steps * 1e-3) / 1e12) if conv_ms > 0 else None
    return dict(metric="full-scene inference Mpixel/s", value=H * W / 1e6 / (ms / 1e3), unit="Mpixel/s", ms_per_step=ms,
                dtype="f16", scaling="strong",
                e2e={"value": H * W / 1e6 / (ms2 / 1e3), "unit": "Mpixel/s", "h2d_bytes_per_step": int(img.nbytes if u0 is None else img[u0:u1].nbytes),
                     "d2h_bytes_per_step": int(H * W), "api": "Session.scene_infer_host (host scene in, upload streamed under the pass, host label map out)" if world == 1 else
                            "Session.upload_scene + Session.scene_infer (host stripe in, host label map out)"},
                gpu_launches=launches, clocks=clocks,
                roofline={"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 fprop)", "achieved": ach, "peak": pk["tc_sustained"],
                          "unit": "TFLOP/s", "frac": ach / pk["tc_sustained"] if ach else None, "traffic": profiled_traffic("inference"),
                          "peak_source": pk["source"] + " bf16 sustained", "launches": conv_n,
                          "kernel_ms_per_step": conv_ms / steps},
                config={"workload": "configs[3]: dilated_grsl_rate8 full-scene sliding-window inference, Potsdam-shaped %dx%dx5 "
                                    "float64, crop 25 stride 12, row stripes over %d GPU(s)" % (H, W, world),
                        "net": cfg["net"], "crop": cfg["crop"], "patches": int(len(drs_b200.grid_positions(H, W, cfg["crop"], cfg["batch"]))),
                        "parallelism": "stripes%d" % world, "l2": "scene (%.0f MB) and activations exceed L2" % (img.nbytes / 1e6)})


# ----------------------------------------------------------------------------------------------------------------
# reference-equivalent CPU path (oracle): the checker timed as the baseline
# ----------------------------------------------------------------------------------------------------------------
def cpu_train_baseline(steps, warmup, budget_s=25.0):
    import torch
    from drs_b200 import synth
    from oracle import host_np, nets_torch
    cfg = TRAIN_CFG
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    img, lab = synth.scene(cfg["dataset"], H=600, W=700)     # gather cost does not depend on scene size
    mean, std = synth.normalisation(img)
    B = cfg["batch"]
    th = TrainHost(cfg, [img.shape[:2]], B)
    orc = nets_torch.OracleNet(cfg["net"], cfg["C"], cfg["K"], nets_torch.init_params(cfg["net"], cfg["C"], cfg["K"], seed=5))

    def step():
        crop, idx, inst, flips, _, rot_on = th.next_plan()
        # the reference's own rotation (isprs:294-296) for the patches the plan rotates
        import scipy.ndimage
        over_x = np.zeros((B, crop, crop, cfg["C"]), dtype=np.float64)
        over_y = np.zeros((B, crop, crop), dtype=np.uint8)
        for b in np.nonzero(rot_on)[0]:
            r, c = int(inst[b, 1]), int(inst[b, 2])
            over_x[b] = scipy.ndimage.rotate(img[r:r + crop, c:c + crop], int(th.angles[b]), order=0, reshape=False)
            over_y[b] = scipy.ndimage.rotate(lab[r:r + crop, c:c + crop], int(th.angles[b]), order=0, reshape=False)
        xs, ys = host_np.apply_plan([img], [lab], inst, flips, crop, mean, std, None, None, over_x, over_y, rot_on)
        loss, pred, _ = orc.train_step(torch.from_numpy(xs.reshape(B, -1)), torch.from_numpy(ys.reshape(B, -1)), crop,
                                       cfg["lr"], cfg["wd"])
        acc, acc_norm, cm = host_np.confusion_by_crop(ys.astype(np.int64), pred.numpy(), cfg["K"])
        th.update(idx, loss, cm)
        return crop

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done, crops = 0, []
    for _ in range(steps):
        crops.append(step())
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(value=B * done / dt, unit="patches/s", cores=cores, kind="port", ms_per_step=dt / done * 1e3, steps=done,
                sample="%d full steps (batch 64, crops %s) of the same seeded loop: NumPy gather + scipy order-0 rotation + normalise, PyTorch-CPU fp32 graph "
                       "fwd+bwd+momentum, Python-loop calc_accuracy_by_crop; TensorFlow not installable (SURVEY F13)" % (done, crops))


def cpu_infer_baseline(batches=6):
    import torch
    from drs_b200 import synth
    from oracle import host_np, nets_torch
    cfg = INFER_CFG
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    img, _ = synth.scene(cfg["dataset"], H=300, W=400)
    mean, std = synth.normalisation(img)
    orc = nets_torch.OracleNet(cfg["net"], cfg["C"], cfg["K"], nets_torch.init_params(cfg["net"], cfg["C"], cfg["K"], seed=9))
    crop, B = cfg["crop"], cfg["batch"]
    stride = host_np.sliding_stride(crop)
    prob = np.zeros((300, 400, cfg["K"]), dtype=np.float32)
    occ = np.zeros((300, 400, cfg["K"]), dtype=np.uint32)

    def one(i):
        pos = host_np.patch_positions(300, 400, crop, stride, i, B, "isprs")
        xs, _ = host_np.apply_plan([img], None, [(0, r, c) for r, c in pos], None, crop, mean, std)
        _, logits = orc.infer(torch.from_numpy(xs.reshape(len(pos), -1)), crop)
        lg = logits.numpy()
        for j, (r, c) in enumerate(pos):
            prob[r:r + crop, c:c + crop] += lg[j]
            occ[r:r + crop, c:c + crop] += 1
        return len(pos)

    one(0)
    t0 = time.perf_counter()
    n = sum(one(i) for i in range(1, 1 + batches))
    dt = time.perf_counter() - t0
    # scene pixels per second = patches/s * (H*W / patches of the full scene); extrapolated linearly by patch count
    total_patches = 499 * 499
    sec_full = total_patches / (n / dt)
    return dict(value=cfg["H"] * cfg["W"] / 1e6 / sec_full, unit="Mpixel/s", cores=cores, kind="port", patches_per_s=n / dt,
                sample="%d batches of 64 patches (crop 25) through the PyTorch-CPU fp32 graph + NumPy accumulate, extrapolated "
                       "linearly to the 249001 patches of a 6000x6000 scene" % batches)


def run_reference(args, rank, world):
    if rank != 0:
        return None
    if args.workload == "infer":
        b = cpu_infer_baseline(batches=max(4, min(args.steps, 16)))
        line = dict(metric="full-scene inference Mpixel/s", value=b["value"], unit=b["unit"], ms_per_step=None, dtype="f32",
                    scaling="strong", config={"workload": "configs[3] on the host cores (bounded sample, extrapolated)"})
    else:
        b = cpu_train_baseline(args.steps, min(args.warmup, 1), budget_s=150.0)
        line = dict(metric="train patches/s", value=b["value"], unit=b["unit"], ms_per_step=b["ms_per_step"], dtype="f32",
                    scaling="weak", config={"workload": "configs[1] on the host cores: " + b["sample"]})
    line.update(impl="reference", cpu_baseline=b, gpu_launches=0,
                e2e={"value": b["value"], "unit": b["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "infer"])
    ap.add_argument("--no-secondary", action="store_true", help="skip the second headline metric")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--small", action="store_true", help="1500x1500 inference scene (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world, local = dist_env()
    base = dict(n_gpus=world, steps=args.steps, warmup=args.warmup, higher_is_better=True, vs_baseline=None, data="synthetic")

    if args.impl == "reference":
        line = run_reference(args, rank, world)
        if line is not None:
            base.update(line)
            print(json.dumps(base), flush=True)
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU baseline)")
    import torch.distributed as dist
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        if args.workload == "train":
            line = run_train_ours(args, rank, world, local)
            second = None if args.no_secondary else run_infer_ours(args, rank, world, local)
        else:
            line = run_infer_ours(args, rank, world, local, steps=max(1, min(args.steps, 3)))
            second = None
        if rank == 0:
            base.update(line)
            base["impl"] = "ours"
            if second is not None:
                base["inference"] = second
            if world == 1 and not args.no_cpu:
                base["cpu_baseline"] = cpu_train_baseline(4, 1, budget_s=20.0) if args.workload == "train" else cpu_infer_baseline()
                if second is not None:
                    base["inference"]["cpu_baseline"] = cpu_infer_baseline()
            print(json.dumps(base), flush=True)
    finally:
        if world > 1:
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
