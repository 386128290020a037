"""A/B two builds of libdrs.so on the same box: scene inference (Mpixel/s, conv roofline fraction) and training
ms/step at fixed patch sizes.

    python tools/ab_lib.py libdrs_old.so libdrs.so [--hw 3000] [--net dilated_grsl_rate8]

Each library runs in its own process (DRS_LIB selects it), so nothing is shared between the arms."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "dynamic-rs-segmentation_b200")


def child():
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import drs_b200
    from drs_b200 import synth
    hw = int(os.environ.get("AB_HW", "3000"))
    net = os.environ.get("AB_NET", "dilated_grsl_rate8")
    tag = os.path.basename(os.environ.get("DRS_LIB", "libdrs.so"))
    if os.environ.get("AB_INFER", "1") == "1":
        img, _ = synth.scene("potsdam", H=hw, W=hw)
        mean, std = synth.normalisation(img)
        s = drs_b200.Session(net, 5, 6, precision="f16", seed=9)
        s.set_stream(torch.cuda.current_stream().cuda_stream)
        s.set_normalization(mean, std)
        s.upload_scene(0, img)
        s.scene_infer(0, 25, 64, hw, hw)            # full warm-up pass: scene-sized buffers are allocated here
        torch.cuda.synchronize()
        s.set_profiling(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.scene_infer(0, 25, 64, hw, hw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        cms, cn, cfl = s.profile_read()
        tag += " waves=%s" % os.environ.get("DRS_CHUNK_WAVES", "default")
        print("%s infer %dx%d %s: %.2f Mpx/s  %.1f ms  conv %.1f ms  %.0f TFLOP/s" %
              (tag, hw, hw, net, hw * hw / 1e3 / ms, ms, cms, cfl / cms / 1e9), flush=True)
        s.close()
    B, C, K = 64, 4, 6
    for tnet in os.environ.get("AB_TRAIN_NETS", "dilated_grsl").split(","):
        if not tnet:
            continue
        s = drs_b200.Session(tnet, C, K, precision="bf16", seed=1)
        s.set_stream(torch.cuda.current_stream().cuda_stream)
        s.reserve(B, 49)
        for crop in (25, 37, 49):
            x = torch.randn(B * crop * crop * C, device="cuda")
            y = torch.randint(0, K, (B * crop * crop,), device="cuda").float()
            pred = torch.empty(B * crop * crop, dtype=torch.uint8, device="cuda")
            cm = torch.zeros(K * K + 1, dtype=torch.int32, device="cuda")
            for _ in range(4):
                s.train_step_dev(x, y, B, crop, pred_dev=pred, cm_dev=cm)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                s.train_step_dev(x, y, B, crop, pred_dev=pred, cm_dev=cm, want_loss=False)
            e1.record()
            torch.cuda.synchronize()
            print("%s train %s crop %d: %.3f ms/step" % (tag, tnet, crop, e0.elapsed_time(e1) / 20), flush=True)
        s.close()


if __name__ == "__main__":
    if os.environ.get("AB_CHILD"):
        child()
    else:
        args = sys.argv[1:]
        env = dict(os.environ, AB_CHILD="1")
        libs = []
        while args:
            a = args.pop(0)
            if a == "--hw":
                env["AB_HW"] = args.pop(0)
            elif a == "--net":
                env["AB_NET"] = args.pop(0)
            elif a == "--train-nets":
                env["AB_TRAIN_NETS"] = args.pop(0)
            elif a == "--no-infer":
                env["AB_INFER"] = "0"
            else:
                libs.append(a)
        for lib in libs or ["libdrs.so"]:
            e = dict(env, DRS_LIB=lib if os.path.isabs(lib) else os.path.join(PKG, lib))
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=e, check=False)
