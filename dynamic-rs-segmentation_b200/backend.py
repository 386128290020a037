"""GPU backend of the drop-in scripts: executes a host ``BatchPlan`` (host.py) on the device.

The scripts' loops call a backend with three operations; ``GpuBackend`` is the product implementation
(libdrs.so through ``Session``, torch only for device buffers).  Tests drive the same loops with a
closed-form stand-in to pin the host logic against reference-generated traces without a GPU.

    train_on_plan(plan, loss_mask=None)  -> (loss, cm[K,K] uint32, n_correct)   gather + normalise + sess.run(train) + confusion
    eval_on_plan(plan, scene_offset=0)   -> (pred int64 [B,c,c], labels [B,c,c]) gather + normalise + sess.run(pred_up)
    scene_labels(scene_id, crop, batch, variant) -> uint8 [H,W]   whole sliding-window inference
    save(path) / restore(path)                                    tf.train.Saver stand-in (one .npz keyed by TF names)
"""
import numpy as np


class GpuBackend:
    def __init__(self, session, scenes, label_maps, mean_full, std_full, device=0, rank=0, world=1, train_fp16_patches=False):
        import torch
        self.train_fp16_patches = train_fp16_patches      # coffee:293
        self.rank, self.world = rank, world
        self.torch = torch
        self.s = session
        self.dev = torch.device("cuda", device)
        self.K = session.num_classes
        self.C = session.channels
        self.shapes = []
        self.has_labels = label_maps is not None
        for i, sc in enumerate(scenes):
            lab = None if label_maps is None else label_maps[i]
            session.upload_scene(i, np.ascontiguousarray(sc), lab)
            self.shapes.append(sc.shape[:2])
        session.set_normalization(mean_full, std_full)
        self._x = self._y = self._pred = self._mask = None
        self._amask_on_dev = False
        self.rotate_on_device = True      # isprs rotation augmentation in the gather kernel (plan_isprs_batch flag)
        self._cm = torch.zeros(self.K * self.K + 1, dtype=torch.int32, device=self.dev)
        session.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)

    def _buffers(self, B, crop):
        t = self.torch
        n = B * crop * crop
        if self._x is None or self._x.numel() < n * self.C:
            self._x = t.empty(n * self.C, dtype=t.float32, device=self.dev)
            self._y = t.empty(n, dtype=t.float32, device=self.dev)
            self._pred = t.empty(n, dtype=t.uint8, device=self.dev)
            self._mask = t.empty(n, dtype=t.uint8, device=self.dev)
            self._amask = t.empty(n, dtype=t.uint8, device=self.dev)
        return self._x, self._y, self._pred

    def _rank_rows(self, n):
        """Data parallel: this rank's contiguous share of a global batch."""
        per = n // self.world
        return slice(self.rank * per, (self.rank + 1) * per) if self.world > 1 else slice(0, n)

    def _gather(self, plan, scene_offset=0, shard=False):
        if shard and self.world > 1:
            sl = self._rank_rows(plan.inst.shape[0])
            sub = type(plan)()
            for k in plan.__slots__:
                v = getattr(plan, k)
                setattr(sub, k, v[sl] if isinstance(v, np.ndarray) else v)
            plan = sub
        self._plan = plan
        B, crop = plan.inst.shape[0], plan.crop
        x, y, pred = self._buffers(B, crop)
        inst = plan.inst
        if scene_offset:
            inst = inst.copy()
            inst[:, 0] += scene_offset
        self._amask_on_dev = False
        if getattr(plan, "rot", None) is not None:
            # rotation on the device: the kernel also writes the accuracy mask
            self.s.gather_rot_dev(inst, plan.flips, crop, x, y, noise=plan.noise, noise_on=plan.noise_on, rot=plan.rot,
                                  rot_on=plan.rot_on, amask_out_dev=self._amask)
            self._amask_on_dev = True
        else:
            self.s.gather_dev(inst, plan.flips, crop, x, y, noise=plan.noise, noise_on=plan.noise_on,
                              over_x=plan.over_x, over_y=plan.over_y, over_on=plan.over_on)
        return x, y, pred, B, crop

    def train_on_plan(self, plan, loss_mask=None):
        t = self.torch
        if self.train_fp16_patches:
            self.s.set_gather_fp16(True)
        x, y, pred, B, crop = self._gather(plan, shard=True)
        if self.train_fp16_patches:
            self.s.set_gather_fp16(False)
        plan = self._plan
        n = B * crop * crop
        mask_dev = None
        if isinstance(loss_mask, str):
            # "label!=K": the session derives the mask on the device from the gathered labels (Session.set_ignore_label)
            self.s.set_ignore_label(int(loss_mask.split("!=")[1]))
        elif loss_mask is not None:
            self._mask[:n].copy_(t.from_numpy(np.ascontiguousarray(loss_mask, dtype=np.uint8).reshape(-1)))
            mask_dev = self._mask
        amask_dev = self._amask if self._amask_on_dev else None
        if plan.acc_mask is not None:
            self._amask[:n].copy_(t.from_numpy(np.ascontiguousarray(plan.acc_mask, dtype=np.uint8).reshape(-1)))
            amask_dev = self._amask
        loss = self.s.train_step_dev(x, y, B, crop, mask_dev=mask_dev, pred_dev=pred, cm_dev=self._cm,
                                     acc_mask_dev=amask_dev)
        cm = self._cm.cpu().numpy().astype(np.uint32)
        K = self.K
        return loss, cm[:K * K].reshape(K, K), int(cm[K * K])

    def eval_on_plan(self, plan, scene_offset=0):
        x, y, pred, B, crop = self._gather(plan, scene_offset)
        self.s.infer_dev(x, B, crop, None, pred)
        self.s.synchronize()
        n = B * crop * crop
        return (pred[:n].cpu().numpy().astype(np.int64).reshape(B, crop, crop),
                y[:n].cpu().numpy().astype(np.int64).reshape(B, crop, crop))

    def scene_labels(self, scene_id, crop, batch, variant="isprs"):
        H, W = self.shapes[scene_id]
        if self.world == 1:
            return self.s.scene_infer(scene_id, crop, batch, H, W, variant=variant)
        from . import dist as ddist
        r0, r1 = ddist.stripe_bounds(H, self.world, self.rank)
        stripe = self.s.scene_infer(scene_id, crop, batch, H, W, variant=variant, row_begin=r0, row_end=r1)
        return ddist.gather_label_stripes(stripe, H, W, self.rank, self.world, device=self.dev, all_ranks=True)

    def scene_mean_logits(self, scene_id, crop, batch, variant="isprs"):
        """``prob_im / occur_im`` (float64 [H,W,K]) of one sliding-window pass: the per-scale map of the multi-scale
        evaluation (isprs:1380-1411)."""
        H, W = self.shapes[scene_id]
        if self.world != 1:
            raise NotImplementedError("multi-scale evaluation runs on one rank")
        _, mean = self.s.scene_infer(scene_id, crop, batch, H, W, variant=variant, want_mean=True)
        return mean

    def scene_confusion(self, scene_id, num_classes, ignore_label=None):
        """Confusion counts of the label map scene_labels() just produced against the resident ground truth, on the device
        (isprs:1289-1296).  None when the ground truth is not resident or the map is striped over ranks (host path then)."""
        if not self.has_labels or self.world != 1:
            return None
        cm, _ = self.s.scene_confusion(scene_id, num_classes, ignore_label)
        return cm.astype(np.uint32)

    def save(self, path):
        self.s.save(path)

    def restore(self, path):
        self.s.restore(path)
