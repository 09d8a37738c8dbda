"""Weight formats on either side of the gather path: `.ot` archives and the export/base65536 strings.
CPU only."""
import hashlib
import json
import os
import re

import numpy as np
import pytest

import synthesis_b200 as s
from synthesis_b200 import weights as W

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
REF = "/root/reference"


def test_read_ot_written_by_libtorch():
    """An archive written by libtorch's own OutputArchive (what tch's VarStore::save calls) reads back to the exact
    tensors that were written."""
    meta = json.load(open(os.path.join(GOLD, "libtorch_model.json")))
    named = W.read_ot(os.path.join(GOLD, "libtorch_model.ot"))
    assert list(named) == [f"l_{l}.{k}" for l in range(1, 6) for k in ("weight", "bias")]
    assert named["l_1.weight"].shape == (128, 63) and named["l_5.bias"].shape == (12,)
    net = s.Connect4Net.load_ot(os.path.join(GOLD, "libtorch_model.ot"))
    blob = net.blob()
    assert blob.size == meta["n_floats"] and hashlib.sha256(blob.tobytes()).hexdigest() == meta["blob_sha256"]
    assert [float(x) for x in blob[:8]] == meta["first8"] and [float(x) for x in blob[-4:]] == meta["last4"]


def test_write_ot_round_trip_and_loads_in_libtorch(tmp_path):
    net = s.Connect4Net.new(3)
    p = str(tmp_path / "model_7.ot")
    net.save_ot(p)
    back = s.Connect4Net.load_ot(p)
    assert back.blob().tobytes() == net.blob().tobytes()
    # libtorch's reader (the one tch's VarStore::load goes through) accepts the archive and sees the same variables
    torch = pytest.importorskip("torch")
    m = torch.jit.load(p)
    got = {k: v.detach().numpy() for k, v in m.named_parameters()}
    assert set(got) == set(net.params)
    for k, v in net.params.items():
        assert got[k].dtype == np.float32 and got[k].shape == v.shape and got[k].tobytes() == v.tobytes()
    # the pickle program is the one libtorch emits for the same tensors
    import zipfile
    ours = zipfile.ZipFile(p)
    theirs = zipfile.ZipFile(os.path.join(GOLD, "libtorch_model.ot"))
    assert ours.read("model_7/data.pkl") == theirs.read("model/data.pkl")  # same names and shapes -> same bytes
    assert ours.read("model_7/code/__torch__.py") == theirs.read("model/code/__torch__.py")
    assert ours.read("model_7/constants.pkl") == theirs.read("model/constants.pkl")


def test_read_ot_rejects_foreign_pickles(tmp_path):
    import pickle
    import zipfile
    p = str(tmp_path / "evil.ot")
    with zipfile.ZipFile(p, "w") as z:
        z.writestr("evil/data.pkl", pickle.dumps(os.getcwd, protocol=2))
    with pytest.raises(pickle.UnpicklingError):
        W.read_ot(p)
    with pytest.raises(KeyError):
        q = str(tmp_path / "partial.ot")
        W.write_ot(q, {"l_1.weight": np.zeros((128, 63), np.float32)})
        s.Connect4Net.load_ot(q)


def test_base65536_block_table_and_round_trip():
    meta = json.load(open(os.path.join(GOLD, "libtorch_model.json")))
    assert hashlib.sha256(",".join(map(str, W.BLOCK_START)).encode()).hexdigest() == meta["base65536_block_start_sha256"]
    if os.path.exists(os.path.join(REF, "base65536/src/lib.rs")):  # in the authoring container: the reference's table itself
        txt = open(os.path.join(REF, "base65536/src/lib.rs")).read()
        tab = [int(x) for x in re.findall(r"\d+", re.search(r"BLOCK_START: \[u32; 256\] = \[(.*?)\];", txt, re.S).group(1))]
        assert tab == W.BLOCK_START
    # base65536/src/lib.rs:62-68 test_encode
    assert W.base65536_decode(W.base65536_encode(b"Hello World")) == b"Hello World"
    # the published example of the base65536 encoding the reference's table comes from (README of qntm/base65536)
    assert W.base65536_encode(b"hello world") == "\u9a68\ua36c\u556f\U00012077\ua372\u1564"
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 255, 1000, 1001):
        raw = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        enc = W.base65536_encode(raw)
        assert len(enc) == (n + 1) // 2 and W.base65536_decode(enc) == raw
    with pytest.raises(ValueError):
        W.base65536_decode("A")


def test_f32_to_bf16_follows_export():
    """export/src/main.rs:8-26 bit for bit, against a scalar restatement and torch's own round-to-nearest-even cast."""
    def scalar(bits):
        if bits & 0x7FFFFFFF > 0x7F800000:
            return ((bits >> 16) | 0x0040) & 0xFFFF
        rb = 0x8000
        if (bits & rb) != 0 and (bits & (3 * rb - 1)) != 0:
            return ((bits >> 16) + 1) & 0xFFFF
        return (bits >> 16) & 0xFFFF
    rng = np.random.default_rng(1)
    bits = np.concatenate([rng.integers(0, 1 << 32, 20000, dtype=np.uint64).astype(np.uint32),
                           np.array([0, 0x80000000, 0x3F808000, 0x3F818000, 0x3F808001, 0x7F800000, 0xFF800000, 0x7FC00000, 0x7F800001,
                                     0x7F7FFFFF, 0x00008000, 0x00018000], np.uint32)])
    got = W.f32_to_bf16(bits.view(np.float32))
    assert [int(x) for x in got] == [scalar(int(b)) for b in bits]
    torch = pytest.importorskip("torch")
    finite = np.isfinite(bits.view(np.float32))
    t = torch.from_numpy(bits.view(np.float32)[finite].copy()).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(got[finite], t)


def test_export_file_and_slimnn_loading():
    net = s.Connect4Net.new(5)
    text = W.export_parameters(net.params)
    lines = text.splitlines()
    assert lines[0] == "load_2d(&mut policy.l_1.weight, String::from(PARAMETERS[0]));"
    assert lines[1] == "load_1d(&mut policy.l_1.bias, String::from(PARAMETERS[1]));"
    assert lines[10] == "const PARAMETERS: [&'static str; 10] = [" and lines[11] == "// l_1 - 0" and lines[-1] == "];"
    strings = W.parse_parameters(text)
    assert len(strings) == 10 and len(strings[0]) == 128 * 63  # two bytes per value = one code point per value
    names = [f"l_{l}.{k}" for l in range(1, 6) for k in ("weight", "bias")]
    for name, st in zip(names, strings):
        v = W.load_nd(st, net.params[name].shape, kind="bf16")
        want = (W.f32_to_bf16(net.params[name]).astype(np.uint32) << 16).view(np.float32).reshape(net.params[name].shape)
        assert v.tobytes() == want.tobytes()
        assert np.allclose(v, net.params[name], rtol=2 ** -8, atol=0)
    # slimnn::loading reads big-endian f32 (loading.rs:3-16): lossless with kind="f32"; a wrong element count is refused
    s32 = W.serialize_tensor(net.params["l_2.weight"], kind="f32")
    assert W.load_nd(s32, (96, 128)).tobytes() == net.params["l_2.weight"].tobytes()
    with pytest.raises(ValueError):
        W.load_nd(s32, (96, 127))


def test_ot_reader_rejects_views_outside_their_storage(tmp_path):
    """A truncated or crafted archive must not make the reader look outside a tensor's storage record (as_strided on sizes,
    strides and offsets taken from the pickle)."""
    import pickle
    from synthesis_b200.weights import _rebuild_tensor_v2
    storage = np.arange(12, dtype=np.float32)
    assert _rebuild_tensor_v2(storage, 0, (3, 4), (4, 1)).shape == (3, 4)
    assert float(_rebuild_tensor_v2(storage, 11, (), ())) == 11.0
    for off, size, stride in ((0, (4, 4), (4, 1)), (1, (3, 4), (4, 1)), (0, (3, 4), (5, 1)), (-1, (3,), (1,)), (0, (3,), (-1,)), (12, (), ()), (0, (3, 4), (4,))):
        with pytest.raises(pickle.UnpicklingError):
            _rebuild_tensor_v2(storage, off, size, stride)
