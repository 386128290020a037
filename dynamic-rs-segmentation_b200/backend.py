"""GPU backend of the drop-in scripts: executes a host ``BatchPlan`` (host.py) on the device.

The scripts' loops call a backend with three operations; ``GpuBackend`` is the product implementation
(libdrs.so through ``Session``, torch only for device buffers).  Tests drive the same loops with a
closed-form stand-in to pin the host logic against reference-generated traces without a GPU.

    train_on_plan(plan)  -> (loss, cm[K,K] uint32, n_correct)   gather + normalise + sess.run(train) + per-crop confusion
    eval_on_plan(plan)   -> (pred int64 [B,c,c], labels [B,c,c]) gather + normalise + sess.run(pred_up)
    scene_labels(scene_id, crop, batch, H, W, variant) -> uint8 [H,W]   whole sliding-window inference
"""
import numpy as np


class GpuBackend:
    def __init__(self, session, scenes, label_maps, mean_full, std_full, device=0):
        import torch
        self.torch = torch
        self.s = session
        self.dev = torch.device("cuda", device)
        self.K = session.num_classes
        self.C = session.channels
        self.shapes = []
        for i, sc in enumerate(scenes):
            lab = None if label_maps is None else label_maps[i]
            session.upload_scene(i, np.ascontiguousarray(sc), lab)
            self.shapes.append(sc.shape[:2])
        session.set_normalization(mean_full, std_full)
        self._x = self._y = self._pred = self._mask = None
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

    def _gather(self, plan):
        B, crop = plan.inst.shape[0], plan.crop
        x, y, pred = self._buffers(B, crop)
        self.s.gather_dev(plan.inst, plan.flips, crop, x, y, noise=plan.noise, noise_on=plan.noise_on,
                          over_x=plan.over_x, over_y=plan.over_y, over_on=plan.over_on)
        return x, y, pred, B, crop

    def train_on_plan(self, plan, loss_mask=None):
        t = self.torch
        x, y, pred, B, crop = self._gather(plan)
        n = B * crop * crop
        mask_dev = None
        if loss_mask is not None:
            self._mask[:n].copy_(t.from_numpy(np.ascontiguousarray(loss_mask, dtype=np.uint8).reshape(-1)))
            mask_dev = self._mask
        amask_dev = None
        if plan.acc_mask is not None:
            self._amask[:n].copy_(t.from_numpy(np.ascontiguousarray(plan.acc_mask, dtype=np.uint8).reshape(-1)))
            amask_dev = self._amask
        loss = self.s.train_step_dev(x, y, B, crop, mask_dev=mask_dev, pred_dev=pred, cm_dev=self._cm,
                                     acc_mask_dev=amask_dev)
        cm = self._cm.cpu().numpy().astype(np.uint32)
        K = self.K
        return loss, cm[:K * K].reshape(K, K), int(cm[K * K])

    def eval_on_plan(self, plan):
        x, y, pred, B, crop = self._gather(plan)
        self.s.infer_dev(x, B, crop, None, pred)
        self.s.synchronize()
        n = B * crop * crop
        return (pred[:n].cpu().numpy().astype(np.int64).reshape(B, crop, crop),
                y[:n].cpu().numpy().astype(np.int64).reshape(B, crop, crop))

    def scene_labels(self, scene_id, crop, batch, variant="isprs"):
        H, W = self.shapes[scene_id]
        return self.s.scene_infer(scene_id, crop, batch, H, W, variant=variant)
