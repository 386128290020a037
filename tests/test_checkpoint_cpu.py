"""CPU: the checkpoint container of drs_save / drs_load (an uncompressed .npz written and read by the library's own host
code, csrc/npz_io.cpp) against NumPy in both directions.  Replaces tf.train.Saver's files (isprs:1693-1717, 1797-1802);
the GPU round trip (train, save, restore, continue bit-identically) is tests/test_gpu_parity.py."""
import os
import zipfile

import numpy as np
import pytest


def test_library_written_npz_is_read_by_numpy(drs, tmp_path):
    rs = np.random.RandomState(3)
    arrays = {"conv1__weights": rs.randn(5, 5, 4, 64).astype(np.float32), "conv1__biases": rs.randn(64).astype(np.float32),
              "conv1__weights__Momentum": rs.randn(5, 5, 4, 64).astype(np.float32),
              "se1_fc1__weights": rs.randn(64, 4).astype(np.float32), "global_step": np.array([50000.0], np.float32),
              "scalar": np.float32(2.5), "empty": np.zeros((0, 3), np.float32)}
    path = str(tmp_path / "model-5.npz")
    drs.npz_write(path, arrays)
    assert not [f for f in os.listdir(tmp_path) if ".tmp." in f], "temp file left behind"
    with np.load(path) as z:
        assert z.files == list(arrays)                       # archive order = the order given
        for k, v in arrays.items():
            assert z[k].dtype == np.float32 and z[k].shape == np.shape(v) and np.array_equal(z[k], v), k
    with zipfile.ZipFile(path) as zf:
        assert zf.testzip() is None                          # CRC-32 of every member
        assert all(i.compress_type == zipfile.ZIP_STORED for i in zf.infolist())
    # overwriting goes through a temp file + rename
    drs.npz_write(path, {"a": np.arange(3, dtype=np.float32)})
    with np.load(path) as z:
        assert z.files == ["a"]


def test_numpy_written_npz_is_read_by_the_library(drs, tmp_path):
    rs = np.random.RandomState(4)
    arrays = {"conv2__weights": rs.randn(5, 5, 64, 64).astype(np.float32), "conv2__moving_variance": rs.rand(64).astype(np.float32),
              "as_f64": rs.randn(7, 3), "as_i64": np.arange(-3, 9, dtype=np.int64).reshape(3, 4),
              "as_i32": np.arange(5, dtype=np.int32), "as_u8": np.arange(250, 256, dtype=np.uint8), "global_step": np.array([7.0], np.float32),
              "zero_d": np.array(3.0, np.float32)}
    path = str(tmp_path / "np.npz")
    np.savez(path, **arrays)                                 # ZIP64 local headers, sizes only in the central directory
    got = drs.npz_read(path)
    assert list(got) == list(arrays)
    for k, v in arrays.items():
        assert got[k].shape == v.shape and np.array_equal(got[k], v.astype(np.float32)), k


def test_round_trip_is_bit_exact_for_special_values(drs, tmp_path):
    v = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3.4028235e38, 1.17549435e-38], dtype=np.float32)
    drs.npz_write(str(tmp_path / "s.npz"), {"v": v})
    got = drs.npz_read(str(tmp_path / "s.npz"))["v"]
    assert got.view(np.uint32).tolist() == v.view(np.uint32).tolist()


def test_refusals(drs, tmp_path):
    from drs_b200 import lib
    np.savez_compressed(str(tmp_path / "c.npz"), a=np.zeros(100, np.float32))
    with pytest.raises(lib.DrsError, match="compressed"):
        drs.npz_read(str(tmp_path / "c.npz"))
    with pytest.raises(lib.DrsError, match="cannot open"):
        drs.npz_read(str(tmp_path / "missing.npz"))
    (tmp_path / "junk.npz").write_bytes(b"not a zip archive at all, just some bytes")
    with pytest.raises(lib.DrsError, match="not a zip archive"):
        drs.npz_read(str(tmp_path / "junk.npz"))
    np.savez(str(tmp_path / "f.npz"), a=np.asfortranarray(np.arange(6, dtype=np.float32).reshape(2, 3)))
    with pytest.raises(lib.DrsError, match="Fortran"):
        drs.npz_read(str(tmp_path / "f.npz"))
    np.savez(str(tmp_path / "cplx.npz"), a=np.zeros(3, np.complex64))
    with pytest.raises(lib.DrsError, match="not supported"):
        drs.npz_read(str(tmp_path / "cplx.npz"))
    # a flipped payload byte is caught by the CRC
    np.savez(str(tmp_path / "ok.npz"), a=np.arange(64, dtype=np.float32))
    raw = bytearray((tmp_path / "ok.npz").read_bytes())
    raw[raw.index(b"\x93NUMPY") + 200] ^= 0x40
    (tmp_path / "bad.npz").write_bytes(bytes(raw))
    with pytest.raises(lib.DrsError, match="CRC"):
        drs.npz_read(str(tmp_path / "bad.npz"))
    with pytest.raises(lib.DrsError, match="cannot create"):
        drs.npz_write(str(tmp_path / "no_such_dir" / "x.npz"), {"a": np.zeros(1, np.float32)})


def test_tf_checkpoint_converter_names_and_layout(drs, tmp_path):
    """tools/tf_checkpoint_to_npz.py on a stand-in checkpoint reader (TensorFlow is not installable here): names as the
    reference's graphs create them (isprs:655-723, 1685) map onto the library's keys, foreign variables are skipped, the
    file round-trips through the library's container."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("tfconv", os.path.join(root, "tools", "tf_checkpoint_to_npz.py"))
    tfconv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tfconv)
    assert tfconv.npz_key("conv1/weights:0") == "conv1__weights"
    assert tfconv.npz_key("main_conv3/biases/Momentum") == "main_conv3__biases__Momentum"
    assert tfconv.npz_key("conv6/moving_variance") == "conv6__moving_variance"
    assert tfconv.npz_key("se1_fc2/weights/Momentum") == "se1_fc2__weights__Momentum"
    assert tfconv.npz_key("main_global_step") == "global_step"
    assert tfconv.npz_key("beta1_power") is None and tfconv.npz_key("save/Const") is None
    rs = np.random.RandomState(1)
    fake = {"conv1/weights": rs.randn(5, 5, 4, 64).astype(np.float32), "conv1/weights/Momentum": rs.randn(5, 5, 4, 64).astype(np.float32),
            "conv1/biases": np.full(64, 0.1, np.float32), "conv1/moving_mean": rs.randn(64).astype(np.float32),
            "main_global_step": np.int64(150000), "beta1_power": np.float32(0.9)}
    out = str(tmp_path / "model-150000.npz")
    arrays, skipped = tfconv.convert(fake.keys(), fake.__getitem__, out, drs.npz_write)
    assert skipped == ["beta1_power"] and set(arrays) == {"conv1__weights", "conv1__weights__Momentum", "conv1__biases",
                                                            "conv1__moving_mean", "global_step"}
    with np.load(out) as z:
        assert z["global_step"].tolist() == [150000.0] and z["conv1__weights"].shape == (5, 5, 4, 64)
        assert np.array_equal(z["conv1__weights__Momentum"], fake["conv1/weights/Momentum"])


def test_damaged_files_are_refused_not_crashed_on(drs, tmp_path):
    """drs_load parses files it did not write: every truncation and a few thousand random byte flips of a NumPy-written and
    of a library-written archive either load or raise DrsError -- never a crash, never an allocation driven by a corrupt
    header (the process surviving this loop is the assertion)."""
    from drs_b200 import lib
    a = str(tmp_path / "a.npz")
    np.savez(a, conv1__weights=np.arange(60, dtype=np.float32).reshape(1, 3, 4, 5), b=np.arange(7, dtype=np.float64),
             global_step=np.array([3], dtype=np.int64))
    b = str(tmp_path / "b.npz")
    drs.npz_write(b, {"x": np.arange(10, dtype=np.float32), "y": np.zeros((2, 3), np.float32)})
    q = str(tmp_path / "damaged.npz")
    rs = np.random.RandomState(0)
    refused = loaded = 0
    for src in (a, b):
        raw = open(src, "rb").read()
        cases = [raw[:n] for n in range(0, len(raw), 5)]
        for _ in range(1500):
            d = bytearray(raw)
            for _ in range(rs.randint(1, 4)):
                d[rs.randint(0, len(d))] = rs.randint(0, 256)
            cases.append(bytes(d))
        for c in cases:
            with open(q, "wb") as f:
                f.write(c)
            try:
                drs.npz_read(q)
                loaded += 1
            except lib.DrsError:
                refused += 1
    assert refused > 1000 and loaded > 0          # flips in names / padding / dates are harmless, payload flips hit the CRC
