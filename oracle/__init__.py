"""CPU oracle for the dynamic-rs-segmentation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import it, and only as the checker / the timed CPU baseline.  The product path
(``dynamic-rs-segmentation_b200``) never imports this package and fails loudly
when its CUDA library is missing.

Contents
--------
* ``host_np``    -- loop-style NumPy restatement of the reference's host data path
                    (sliding-window grid, overlap accumulate + argmax, per-crop
                    confusion, normalisation, patch gather, patch-size policy).
                    PINNED: checked against golden vectors produced by running the
                    reference's own functions (``oracle/make_golden.py`` imports
                    ``/root/reference`` with stubbed tensorflow/gdal/skimage).
* ``nets_torch`` -- PyTorch-CPU fp32 restatement of the TF-1.x graph (conv + BN +
                    activation + pool + concat nets, loss, momentum step).
                    PARITY UNPINNED: TensorFlow is not installable in this image
                    and the reference ships no test vectors at the sess.run
                    boundary, so this half restates TF's
                    documented semantics.
* ``ref_import`` -- loader for the real reference modules; usable only where
                    ``/root/reference`` exists (the build container), never on
                    the GPU box.
"""
