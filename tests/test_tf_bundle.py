"""TF V2 checkpoint ("tensor bundle") reader without TensorFlow: round trips through the writer,
hand-assembled table blocks (prefix compression, restart points, snappy), CRC failure detection, tf.train.latest_checkpoint
resolution, and the zk / lds variable selection (EMA shadows first / optimizer slots dropped) straight from the files."""
import os
import struct

import numpy as np
import pytest

from kddcup_2020_multimodalitiesrecall_2nd_place_b200 import checkpoints, synth, tf_bundle
from kddcup_2020_multimodalitiesrecall_2nd_place_b200.config import LDS, ZK, ModelConfig


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors
    assert tf_bundle.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert tf_bundle.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tf_bundle.crc32c(bytes(range(32))) == 0x46DD794E
    assert tf_bundle.crc32c(b"123456789") == 0xE3069283
    blob = bytes(np.random.default_rng(0).integers(0, 256, 100003, dtype=np.uint8))
    assert tf_bundle.crc32c(blob) == tf_bundle.crc32c_py(blob)                  # library (hardware) == table reference
    assert tf_bundle.crc32c(blob[50:], tf_bundle.crc32c(blob[:50])) == tf_bundle.crc32c(blob)   # continuation
    # leveldb's masking is an involution partner of Unmask: rotate right 15 + constant
    assert tf_bundle.mask_crc(0) == 0xA282EAD8


def test_round_trip_many_blocks_and_dtypes(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {f"bert/encoder/layer_{i}/attention/self/query/kernel": rng.standard_normal((7, 5)).astype(np.float32)
               for i in range(40)}
    tensors["global_step"] = np.array(251, np.int64)
    tensors["scalar_f"] = np.array(0.5, np.float32)
    tensors["half"] = rng.standard_normal(9).astype(np.float16)
    tensors["dbl"] = rng.standard_normal((2, 3, 4))
    tensors["empty"] = np.zeros((0, 3), np.float32)
    prefix = str(tmp_path / "model.ckpt-251")
    tf_bundle.write_bundle(prefix, tensors, block_entries=7)       # several data blocks + prefix compression
    r = tf_bundle.BundleReader(prefix, verify_data=True)
    assert dict(r.list_variables())["dbl"] == [2, 3, 4] and len(r.list_variables()) == len(tensors)
    for n, a in tensors.items():
        got = r.get_tensor(n)
        assert got.dtype == a.dtype and got.shape == a.shape and np.array_equal(got, a), n
    assert "global_step" not in r.read_all(float_only=True) and "global_step" in r.read_all(float_only=False)
    with pytest.raises(KeyError):
        r.get_tensor("nope")


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "m")
    tf_bundle.write_bundle(prefix, {"a": np.arange(6, dtype=np.float32)})
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[3] ^= 0x40
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(tf_bundle.BundleError, match="crc"):
        tf_bundle.BundleReader(prefix)
    tf_bundle.write_bundle(prefix, {"a": np.arange(6, dtype=np.float32)})
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[5] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(tf_bundle.BundleError, match="data crc"):
        tf_bundle.BundleReader(prefix, verify_data=True).get_tensor("a")
    open(prefix + ".index", "wb").write(b"not a table at all" * 10)
    with pytest.raises(tf_bundle.BundleError, match="magic"):
        tf_bundle.BundleReader(prefix)


def test_hand_assembled_block_with_shared_prefixes_and_snappy(tmp_path):
    """A table written byte by byte from the format description (not by write_bundle): one snappy-compressed data block
    whose second and third keys share a prefix with their predecessor, restart interval 2."""
    def entry(shared, tail, value):
        return bytes([shared, len(tail), len(value)]) + tail + value
    block = entry(0, b"alpha", b"1") + entry(4, b"X", b"22") + entry(0, b"beta", b"333") + entry(2, b"ta", b"")
    block += struct.pack("<III", 0, len(entry(0, b"alpha", b"1") + entry(4, b"X", b"22")), 2)
    # snappy: length, then one literal element holding everything except the last 4 bytes, then a copy of 4 bytes that
    # repeats an earlier span (offset chosen to point at identical bytes is fiddly by hand: use two literals instead)
    comp = bytes([len(block)]) + bytes([(len(block) - 1 - 4) << 2]) + block[:-4] + bytes([(4 - 1) << 2]) + block[-4:]
    assert tf_bundle.snappy_decompress(comp) == block
    body = comp + b"\x01" + struct.pack("<I", tf_bundle.mask_crc(tf_bundle.crc32c(comp + b"\x01")))
    meta_off = len(body)
    meta = struct.pack("<II", 0, 1)
    body += meta + b"\x00" + struct.pack("<I", tf_bundle.mask_crc(tf_bundle.crc32c(meta + b"\x00")))
    idx_off = len(body)
    handle = bytes([0, len(comp)])
    idx = entry(0, b"betta", handle) + struct.pack("<II", 0, 1)
    body += idx + b"\x00" + struct.pack("<I", tf_bundle.mask_crc(tf_bundle.crc32c(idx + b"\x00")))
    footer = bytes([meta_off, len(meta), idx_off, len(idx)])
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", tf_bundle.TABLE_MAGIC)
    path = tmp_path / "hand.index"
    path.write_bytes(body + footer)
    assert tf_bundle.read_table(str(path)) == [(b"alpha", b"1"), (b"alphX", b"22"), (b"beta", b"333"), (b"beta", b"")][:3] + \
        [(b"beta" [:2] + b"ta", b"")]
    # a snappy stream with a real back-reference (copy with 1-byte offset): "abcdabcdabcd"
    assert tf_bundle.snappy_decompress(bytes([12, 3 << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4])) == b"abcdabcdabcd"


@pytest.mark.parametrize("kind", [ZK, LDS])
def test_checkpoint_to_weights_without_tensorflow(kind, tmp_path):
    """What the drivers restore: zk = EMA shadows first (evaluate_normal.py:204-212), lds = raw variables, Adam slots
    ignored (run_pretraining_predict_score.py:347-362) -- from the checkpoint FILES, via a `checkpoint` state file."""
    cfg = ModelConfig(kind, n_layers=2, lq=20, nbox=8, vocab=300)
    w = synth.make_weights(cfg, seed=3)
    rng = np.random.default_rng(1)
    stored = {}
    for name, a in w.items():
        if kind == ZK:
            stored[name] = a + 1.0                                          # the raw variable: NOT what zk restores
            stored[name + "/ExponentialMovingAverage"] = a                  # the shadow: what it restores
        else:
            stored[name] = a
            stored[name + "/adam_m"] = rng.standard_normal(a.shape).astype(np.float32)
            stored[name + "/adam_v"] = rng.standard_normal(a.shape).astype(np.float32)
    stored["global_step"] = np.array(1234, np.int64)
    d = tmp_path / "ckpt"
    d.mkdir()
    tf_bundle.write_bundle(str(d / "model.ckpt-1234"), stored)
    (d / "checkpoint").write_text('model_checkpoint_path: "model.ckpt-1234"\nall_model_checkpoint_paths: "model.ckpt-1234"\n')
    got = checkpoints.tf_variables_from_checkpoint(str(d), prefer_ema=(kind == ZK), wanted=w.keys(), verify_data=True)
    assert list(got) == list(w)
    for name in w:
        assert got[name].dtype == np.float32 and np.array_equal(got[name], w[name]), name
    with pytest.raises(FileNotFoundError):
        checkpoints.tf_variables_from_checkpoint(str(tmp_path))
