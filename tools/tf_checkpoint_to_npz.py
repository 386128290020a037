"""Convert a TensorFlow-1.x checkpoint written by the reference scripts (saver.save(sess, output_path + 'model',
global_step=step), isprs_dilated_random.py:1797-1802) into the .npz that drs_load / Session.restore read.

    python tools/tf_checkpoint_to_npz.py <output_path>/model-150000 model-150000.npz

Needs TensorFlow only for reading (tf.train.load_checkpoint / tf.train.NewCheckpointReader); the writing side is the
library's own container (drs_npz_write), so the result is exactly what drs_save would have produced: one float32 array per
variable, '/' in the TF name written as '__', filters kept in TF's HWIO layout.  Variables the library does not hold
(beta1_power-style optimizer scalars of other optimizers, savers' bookkeeping) are listed and skipped.
"""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# what a variable of the reference's graphs looks like: <scope>/weights|biases|moving_mean|moving_variance[/Momentum], the
# squeeze-and-excitation gates' <name>_fc1|2/weights|biases, and the step counter (isprs:1685: tf.Variable(0, name='main_global_step'))
_KEEP = re.compile(r"^(?:[A-Za-z0-9_]+/)?(?:weights|biases|moving_mean|moving_variance)(?:/Momentum)?$")
_STEP = ("global_step", "main_global_step")


def npz_key(tf_name):
    """'conv1/weights/Momentum:0' -> 'conv1__weights__Momentum'; the step counter -> 'global_step'; None = not ours."""
    name = tf_name.split(":")[0]
    if name in _STEP:
        return "global_step"
    if not _KEEP.match(name):
        return None
    return name.replace("/", "__")


def convert(reader_names, load, out_path, npz_write):
    """reader_names: iterable of TF variable names; load(name) -> ndarray; npz_write(path, {key: array})."""
    import numpy as np
    arrays, skipped = {}, []
    for name in sorted(reader_names):
        key = npz_key(name)
        if key is None:
            skipped.append(name)
            continue
        a = np.asarray(load(name), dtype=np.float32)
        arrays[key] = a.reshape(1) if key == "global_step" else a
    npz_write(out_path, arrays)
    return arrays, skipped


def main(argv):
    if len(argv) != 3:
        print(__doc__)
        return 2
    try:
        import tensorflow as tf
    except ImportError:
        print("TensorFlow is needed to read the checkpoint (any 1.x / 2.x with tf.train.load_checkpoint)")
        return 1
    import drs_b200
    reader = tf.train.load_checkpoint(argv[1])
    names = list(reader.get_variable_to_shape_map())
    arrays, skipped = convert(names, reader.get_tensor, argv[2], drs_b200.npz_write)
    print("wrote %s: %d variables, %d skipped%s" % (argv[2], len(arrays), len(skipped), (" (" + ", ".join(skipped) + ")") if skipped else ""))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
