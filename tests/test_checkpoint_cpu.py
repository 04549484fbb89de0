"""TensorFlow tensor-bundle checkpoint reader/writer (sap3d_tensorflow_b200/checkpoint.py; the reference's tf.train.Saver
at train.py:180-185,266-267 and gen_pred.py:57-64).  No TensorFlow exists in this image, so the format is pinned by its
published constants and known answers: CRC-32C check values (RFC 3720 B.4), the masked-CRC definition, the table magic, the
BundleHeaderProto bytes, a hand-assembled SSTable block with prefix compression and restarts, plus write->read round trips."""
import os
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.usefixtures("lib_built")


@pytest.fixture()
def ck():
    from sap3d_tensorflow_b200 import checkpoint
    return checkpoint


def test_crc32c_known_answers(ck):
    assert ck.crc32c(b"123456789") == 0xE3069283
    # RFC 3720 B.4 test patterns
    assert ck.crc32c(bytes(32)) == 0x8A9136AA
    assert ck.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E
    assert ck.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    # incremental == one shot, at every split point (covers the unaligned head / 8-byte body / tail paths)
    data = np.random.RandomState(0).randint(0, 256, 257).astype(np.uint8).tobytes()
    whole = ck.crc32c(data)
    for cut in range(0, 257, 7):
        assert ck.crc32c(data[cut:], ck.crc32c(data[:cut])) == whole
    arr = np.frombuffer(data, np.uint8)[1:]                     # unaligned ndarray path
    assert ck.crc32c(arr) == ck.crc32c(data[1:])


def test_crc_mask_roundtrip(ck):
    # leveldb/TF: rotate right by 15, add 0xa282ead8
    assert ck.mask_crc(0) == 0xA282EAD8
    assert ck.mask_crc(0xE3069283) == ((((0xE3069283 >> 15) | (0xE3069283 << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF
    for c in (0, 1, 0xFFFFFFFF, 0xE3069283, 0x12345678):
        assert ck.unmask_crc(ck.mask_crc(c)) == c


def test_varints_and_entry_proto(ck):
    out = bytearray()
    ck._put_varint(out, 300)
    assert bytes(out) == b"\xac\x02" and ck._get_varint(bytes(out), 0) == (300, 2)
    # BundleEntryProto{dtype: DT_FLOAT, shape{dim{size:3} dim{size:4}}, offset: 16, size: 48, crc32c: 0x01020304}
    enc = ck._encode_entry(1, (3, 4), 16, 48, 0x01020304)
    assert enc == b"\x08\x01" + b"\x12\x08" + b"\x12\x02\x08\x03" + b"\x12\x02\x08\x04" + b"\x20\x10" + b"\x28\x30" + b"\x35\x04\x03\x02\x01"
    e = ck._decode_entry(enc)
    assert (e["dtype"], e["shape"], e["shard_id"], e["offset"], e["size"], e["crc32c"]) == (1, (3, 4), 0, 16, 48, 0x01020304)
    # scalar at offset 0: the shape message is present and empty, zero-valued fields are omitted (proto3)
    assert ck._encode_entry(9, (), 0, 8, 0) == b"\x08\x09\x12\x00\x28\x08\x35\x00\x00\x00\x00"
    h = ck._decode_header(ck.HEADER_BYTES)
    assert h == {"num_shards": 1, "endianness": 0, "producer": 1}


def test_hand_assembled_table(ck):
    """an index written by hand from the LevelDB table-format description: one data block whose second key shares a prefix,
    an empty metaindex block, a one-entry index block, the 48-byte footer."""
    def trailer(contents):
        return contents + b"\x00" + struct.pack("<I", ck.mask_crc(ck.crc32c(contents + b"\x00")))
    e1 = b"\x00\x03\x02" + b"abc" + b"v1"            # shared 0, non-shared 3, value 2
    e2 = b"\x02\x02\x02" + b"de" + b"v2"             # key "abde": shares "ab"
    data = e1 + e2 + struct.pack("<II", 0, 1)        # restart array [0], count 1
    meta = struct.pack("<II", 0, 1)
    blob = trailer(data)
    meta_off = len(blob)
    blob += trailer(meta)
    idx_entry = b"\x00\x04\x02" + b"abde" + bytes([0, len(data)])
    idx = idx_entry + struct.pack("<II", 0, 1)
    idx_off = len(blob)
    blob += trailer(idx)
    footer = bytes([meta_off, len(meta), idx_off, len(idx)]).ljust(40, b"\x00") + bytes.fromhex("57fb808b247547db")
    blob += footer
    assert ck.read_table(blob) == [(b"abc", b"v1"), (b"abde", b"v2")]
    # and the writer produces exactly this file for the same two pairs
    assert ck.build_table([(b"abc", b"v1"), (b"abde", b"v2")]) == blob
    # corruption is detected
    bad = bytearray(blob)
    bad[4] ^= 1
    with pytest.raises(ck.CheckpointError, match="checksum"):
        ck.read_table(bytes(bad))
    with pytest.raises(ck.CheckpointError, match="magic"):
        ck.read_table(blob[:-1] + b"\x00")


def test_table_many_keys_multiple_blocks_and_restarts(ck):
    keys = sorted({f"P3D/conv{i % 13}_{i}/kernel".encode() for i in range(500)} | {b""})
    items = [(k, (b"val-" + k) * (1 + len(k) % 3)) for k in keys]
    blob = ck.build_table(items, block_size=512)     # forces ~60 data blocks, >1 restart per block
    assert ck.read_table(blob) == items
    one = ck.build_table(items)                      # default block size: a single data block with 32 restart points
    assert ck.read_table(one) == items


def test_snappy_block(ck):
    # literal "abcd" + copy(offset 4, len 8) -> "abcdabcdabcd" (overlapping back-reference)
    src = bytes([12]) + bytes([3 << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4])
    assert ck._snappy_decompress(src) == b"abcdabcdabcd"


def test_bundle_roundtrip(ck, tmp_path):
    rng = np.random.RandomState(1)
    tensors = {
        "firstconv1": rng.randn(1, 7, 7, 3, 64).astype(np.float32),
        "batch_normalization/gamma": rng.rand(64).astype(np.float32),
        "batch_normalization/moving_mean": rng.randn(64).astype(np.float32),
        "batch_normalization_10/moving_variance": rng.rand(256).astype(np.float32),
        "x_4_0_sa/gamma": np.float32(0.25),
        "x_4_0_sa/conv3d/kernel": rng.randn(1, 1, 1, 1024, 128).astype(np.float32),
        "global_step": np.int64(1234),
        "empty": np.zeros((0, 4), np.float32),
        "beta1_power": np.float32(0.9 ** 5),
        "weights_f64": rng.randn(3, 2),
    }
    prefix = str(tmp_path / "model" / "p3d_1000.ckpt")
    assert ck.save(prefix, tensors) == prefix
    assert sorted(os.listdir(tmp_path / "model")) == ["checkpoint", "p3d_1000.ckpt.data-00000-of-00001", "p3d_1000.ckpt.index"]
    raw = open(prefix + ".index", "rb").read()
    assert raw[-8:] == bytes.fromhex("57fb808b247547db")
    assert os.path.getsize(prefix + ".data-00000-of-00001") == sum(np.asarray(v).nbytes for v in tensors.values())
    back = ck.load(prefix)
    assert set(back) == set(tensors)
    for n, v in tensors.items():
        v = np.asarray(v)
        assert back[n].dtype == v.dtype and back[n].shape == v.shape, n
        np.testing.assert_array_equal(back[n], v)
    # data file is laid out in key order
    names = sorted(tensors, key=lambda s: s.encode())
    assert [n for n, _ in ck.list_variables(prefix)] == names
    first = np.fromfile(prefix + ".data-00000-of-00001", np.float32, 64)
    np.testing.assert_array_equal(first, tensors["batch_normalization/gamma"])
    # subset load / missing name
    sub = ck.load(prefix, ["x_4_0_sa/gamma"])
    assert list(sub) == ["x_4_0_sa/gamma"] and sub["x_4_0_sa/gamma"] == np.float32(0.25)
    with pytest.raises(ck.CheckpointError, match="not in the checkpoint"):
        ck.load(prefix, ["nope"])
    # a flipped data byte fails the tensor checksum
    with open(prefix + ".data-00000-of-00001", "r+b") as f:
        f.seek(100)
        b = f.read(1)
        f.seek(100)
        f.write(bytes([b[0] ^ 0x40]))
    with pytest.raises(ck.CheckpointError, match="tensor checksum"):
        ck.load(prefix)
    assert ck.load(prefix, verify=False)  # readable when asked not to verify


def test_bfloat16_entries_are_widened(ck, tmp_path):
    prefix = str(tmp_path / "m.ckpt")
    vals = np.array([1.0, -2.5, 0.15625, 3.0e38], np.float32)
    bits = (vals.view(np.uint32) >> 16).astype(np.uint16)
    entry = ck._encode_entry(14, (4,), 0, 8, ck.mask_crc(ck.crc32c(bits)))
    open(prefix + ".index", "wb").write(ck.build_table([(b"", ck.HEADER_BYTES), (b"w", entry)]))
    bits.tofile(prefix + ".data-00000-of-00001")
    got = ck.load(prefix)["w"]
    assert got.dtype == np.float32
    np.testing.assert_array_equal(got, (bits.astype(np.uint32) << 16).view(np.float32))


def test_checkpoint_state_file(ck, tmp_path):
    d = str(tmp_path)
    assert ck.get_checkpoint_state(d) is None and ck.latest_checkpoint(d) is None
    t = {"w": np.arange(4, dtype=np.float32)}
    for step in range(1, 5):
        ck.save(os.path.join(d, f"p3d_{step}.ckpt"), t, max_to_keep=3)
    txt = open(os.path.join(d, "checkpoint")).read().splitlines()
    assert txt[0] == 'model_checkpoint_path: "p3d_4.ckpt"'
    assert txt[1:] == [f'all_model_checkpoint_paths: "p3d_{s}.ckpt"' for s in (2, 3, 4)]
    assert not os.path.exists(os.path.join(d, "p3d_1.ckpt.index"))          # rotated out (Saver(max_to_keep))
    assert ck.latest_checkpoint(d) == os.path.join(d, "p3d_4.ckpt")
    st = ck.get_checkpoint_state(d)
    assert st["all_model_checkpoint_paths"] == [os.path.join(d, f"p3d_{s}.ckpt") for s in (2, 3, 4)]
    # a state file written by TF with an absolute path
    open(os.path.join(d, "checkpoint"), "w").write(f'model_checkpoint_path: "{d}/p3d_3.ckpt"\nall_model_checkpoint_paths: "{d}/p3d_3.ckpt"\n')
    assert ck.latest_checkpoint(d) == os.path.join(d, "p3d_3.ckpt")


def test_oracle_variables_roundtrip_by_reference_names(ck, tmp_path):
    """every variable of the p3d_unet graph (reference names, TF layouts) survives a save/load cycle"""
    import torch
    from oracle import p3d_oracle as O

    vs = O.VarStore(seed=3)
    x = torch.zeros(1, 16, 32, 32, 3)
    O.forward("p3d_unet", x, vs, training=False)
    tensors = {n: p.detach().numpy() for n, p in vs.params.items()}
    prefix = str(tmp_path / "unet.ckpt")
    ck.save(prefix, tensors)
    back = ck.load(prefix)
    assert set(back) == set(tensors) and len(back) > 500
    for n in tensors:
        np.testing.assert_array_equal(back[n], tensors[n])
    assert "firstconv1" in back and any(n.endswith("moving_variance") for n in back)


def test_table_and_bundle_roundtrip_properties(ck, tmp_path):
    """hypothesis: any sorted set of keys/values survives build_table -> read_table at any block size; any dict of tensors
    (names with '/', '_', digits; ranks 0-4; the dtypes TF stores for this model) survives save -> load."""
    from hypothesis import given, settings, strategies as st

    keys = st.lists(st.binary(min_size=1, max_size=40), min_size=0, max_size=120, unique=True)

    @settings(max_examples=60, deadline=None)
    @given(keys, st.integers(min_value=32, max_value=4096), st.data())
    def table(ks, block_size, data):
        items = [(k, data.draw(st.binary(max_size=64))) for k in sorted(ks)]
        assert ck.read_table(ck.build_table(items, block_size=block_size)) == items

    table()

    name = st.text(alphabet="abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789_/", min_size=1, max_size=48)
    dtype = st.sampled_from([np.float32, np.float64, np.int32, np.int64, np.uint8, np.float16, np.bool_])
    shape = st.lists(st.integers(min_value=0, max_value=5), min_size=0, max_size=4)
    counter = [0]

    @settings(max_examples=40, deadline=None)
    @given(st.dictionaries(name, st.tuples(dtype, shape), min_size=1, max_size=12), st.integers(0, 2 ** 31 - 1))
    def bundle(spec, seed):
        rng = np.random.RandomState(seed)
        tensors = {n: np.asarray(rng.randn(*s) * 100).astype(d) for n, (d, s) in spec.items()}
        counter[0] += 1
        prefix = str(tmp_path / f"b{counter[0]}.ckpt")
        ck.save(prefix, tensors, update_state=False)
        back = ck.load(prefix)
        assert set(back) == set(tensors)
        for n, v in tensors.items():
            assert back[n].dtype == v.dtype and back[n].shape == v.shape
            np.testing.assert_array_equal(back[n], v)
        assert [n for n, _ in ck.list_variables(prefix)] == sorted(tensors, key=lambda s: s.encode())

    bundle()
