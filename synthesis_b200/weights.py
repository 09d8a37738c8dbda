"""Weight file formats on either side of the gather path, read and written without libtorch.

* `.ot` — what `VarStore::save` / `VarStore::load` exchange between the learner and the self-play
  workers (synthesis/src/alpha_zero.rs:37, 102, 192-194; evaluator.rs).  tch 0.4.1 routes both
  through libtorch's `torch::serialize::OutputArchive::write(name, tensor)` / `InputArchive`: a zip
  archive `<stem>/data.pkl` (pickle protocol 2 of a `__torch__.Module` whose state is a dict
  name -> `torch._utils._rebuild_tensor_v2(storage, offset, size, stride, requires_grad, hooks)`),
  raw little-endian storages under `<stem>/data/<key>`, the module's generated source
  `<stem>/code/__torch__.py`, `<stem>/constants.pkl` and `<stem>/version`.
  `read_ot` parses that with zipfile + a restricted unpickler; `write_ot` emits the same records.
  Pinned by tests/golden/libtorch_model.ot, an archive written by libtorch's own OutputArchive
  (tests/golden/make_ot_fixture.cpp), and by loading `write_ot` output back through `torch.jit.load`.

* the `export` text format (export/src/main.rs:8-92): every tensor as round-to-nearest-even bf16,
  big-endian bytes, base65536 text (base65536/src/lib.rs:26-56), plus the `load_*d` statement list.
  `slimnn::loading` (slimnn/src/loading.rs:3-39) decodes such strings as big-endian *f32*; both
  element encodings are offered (`serialize_tensor(..., kind="bf16" | "f32")`, `load_nd`).
"""
import io
import os
import pickle
import struct
import zipfile
from collections import OrderedDict

import numpy as np

# ------------------------------------------------------------------------------------------ .ot
_STORAGE_DTYPES = {
    "FloatStorage": np.dtype("<f4"), "DoubleStorage": np.dtype("<f8"), "HalfStorage": np.dtype("<f2"),
    "LongStorage": np.dtype("<i8"), "IntStorage": np.dtype("<i4"), "ShortStorage": np.dtype("<i2"),
    "CharStorage": np.dtype("i1"), "ByteStorage": np.dtype("u1"), "BoolStorage": np.dtype("?"),
}


class _StorageType:
    def __init__(self, name):
        self.dtype = _STORAGE_DTYPES[name]


class _Module:
    """Stand-in for `__torch__.Module`: BUILD hands it the name -> tensor dict."""

    def __init__(self):
        self.state = OrderedDict()

    def __setstate__(self, state):
        self.state = OrderedDict(state)


def _rebuild_tensor_v2(storage, storage_offset, size, stride, requires_grad=False, backward_hooks=None, metadata=None):
    size, stride = tuple(int(x) for x in size), tuple(int(x) for x in stride)
    storage_offset = int(storage_offset)
    item = storage.dtype.itemsize
    # the view must lie inside the storage record: a truncated or crafted archive must not read out of bounds
    if len(size) != len(stride) or storage_offset < 0 or any(n < 0 for n in size) or any(st < 0 for st in stride):
        raise pickle.UnpicklingError("tensor with negative or mismatched size / stride / offset")
    last = storage_offset + sum((n - 1) * st for n, st in zip(size, stride)) if all(n > 0 for n in size) else storage_offset - 1
    if last >= len(storage) or (not size and storage_offset >= len(storage)):
        raise pickle.UnpicklingError(f"tensor view (offset {storage_offset}, size {size}, stride {stride}) exceeds its storage of {len(storage)} elements")
    if not size:
        return np.array(storage[storage_offset], dtype=storage.dtype)
    view = np.lib.stride_tricks.as_strided(storage[storage_offset:], shape=size, strides=tuple(s * item for s in stride))
    return np.ascontiguousarray(view)


class _OtUnpickler(pickle.Unpickler):
    def __init__(self, f, read_record):
        super().__init__(f)
        self._read_record = read_record

    def find_class(self, module, name):
        if module.startswith("__torch__"):
            return _Module
        if module == "torch._utils" and name == "_rebuild_tensor_v2":
            return _rebuild_tensor_v2
        if module == "torch" and name in _STORAGE_DTYPES:
            return _StorageType(name)
        if module == "collections" and name == "OrderedDict":
            return OrderedDict
        raise pickle.UnpicklingError(f".ot archive refers to {module}.{name}, which a tensor archive never needs")

    def persistent_load(self, pid):
        tag, storage_type, key, _location, numel = pid
        if tag != "storage":
            raise pickle.UnpicklingError(f"unknown persistent id {tag!r}")
        raw = self._read_record(f"data/{key}")
        arr = np.frombuffer(raw, dtype=storage_type.dtype)
        if arr.size < int(numel):
            raise pickle.UnpicklingError(f"storage {key}: {arr.size} elements on file, {numel} expected")
        return arr


def read_ot(path) -> "OrderedDict[str, np.ndarray]":
    """`Tensor::load_multi` (what VarStore::load reads): name -> array, in file order."""
    with zipfile.ZipFile(path) as z:
        names = z.namelist()
        pkl = [n for n in names if n.endswith("/data.pkl") or n == "data.pkl"]
        if len(pkl) != 1:
            raise ValueError(f"{path}: not a libtorch archive (no data.pkl)")
        prefix = pkl[0][:-len("data.pkl")]
        if prefix + "byteorder" in names and z.read(prefix + "byteorder").strip() != b"little":
            raise ValueError(f"{path}: big-endian archives are not supported")
        obj = _OtUnpickler(io.BytesIO(z.read(pkl[0])), lambda rec: z.read(prefix + rec)).load()
    state = obj.state if isinstance(obj, _Module) else obj
    if not isinstance(state, dict):
        raise ValueError(f"{path}: unexpected archive root {type(obj).__name__}")
    return OrderedDict((str(k), np.asarray(v)) for k, v in state.items())


class _P2:
    """Just enough of a pickle protocol-2 writer to emit data.pkl the way libtorch's Pickler does (memoised globals)."""

    def __init__(self):
        self.out = bytearray(b"\x80\x02")
        self.memo = {}

    def _put(self, key):
        i = len(self.memo)
        self.memo[key] = i
        self.out += (b"q" + bytes([i])) if i < 256 else (b"r" + struct.pack("<I", i))

    def get(self, key):
        i = self.memo[key]
        self.out += (b"h" + bytes([i])) if i < 256 else (b"j" + struct.pack("<I", i))

    def glob(self, module, name):
        key = ("g", module, name)
        if key in self.memo:
            return self.get(key)
        self.out += b"c" + module.encode() + b"\n" + name.encode() + b"\n"
        self._put(key)

    def text(self, s, memo=True):
        key = ("s", s)
        if memo and key in self.memo:
            return self.get(key)
        b = s.encode("utf-8")
        self.out += b"X" + struct.pack("<I", len(b)) + b
        self._put(key if memo else ("anon", len(self.memo)))

    def integer(self, v):
        if 0 <= v < 256:
            self.out += b"K" + bytes([v])
        elif 0 <= v < 65536:
            self.out += b"M" + struct.pack("<H", v)
        elif -2 ** 31 <= v < 2 ** 31:
            self.out += b"J" + struct.pack("<i", v)
        else:
            self.out += b"\x8a\x08" + struct.pack("<q", v)

    def int_tuple(self, vals):
        self.out += b"("
        for v in vals:
            self.integer(int(v))
        self.out += b"t"


def _zip_write(z, name, data, compress=False):
    info = zipfile.ZipInfo(name, date_time=(1980, 1, 1, 0, 0, 0))
    info.compress_type = zipfile.ZIP_DEFLATED if compress else zipfile.ZIP_STORED
    z.writestr(info, data)


def write_ot(path, named_tensors, stem=None) -> None:
    """`Tensor::save_multi` (what VarStore::save writes): float32 tensors under their VarStore names."""
    named = OrderedDict((str(k), np.ascontiguousarray(v, dtype="<f4")) for k, v in named_tensors.items())
    stem = stem or os.path.splitext(os.path.basename(str(path)))[0] or "archive"
    p = _P2()
    p.glob("__torch__", "Module")
    p.out += b")\x81}("
    for key, (name, arr) in enumerate(named.items()):
        p.text(name, memo=False)
        p.glob("torch._utils", "_rebuild_tensor_v2")
        p.out += b"(("
        p.text("storage")
        p.glob("torch", "FloatStorage")
        p.text(str(key), memo=False)
        p.text("cpu")
        p.integer(arr.size)
        p.out += b"tQ"
        p._put(("pid", key))
        p.integer(0)
        p.int_tuple(arr.shape)
        strides = [int(np.prod(arr.shape[i + 1:], dtype=np.int64)) for i in range(arr.ndim)]
        p.int_tuple(strides)
        p.out += b"\x89"
        p.glob("collections", "OrderedDict")
        p.out += b")RtR"
    p.out += b"ub"
    p._put(("root",))
    p.out += b"."
    code = ["class Module(Module):", "  __parameters__ = [" + "".join(f'"{n}", ' for n in named) + "]", "  __buffers__ = []",
            "  __annotations__ = []"] + [f'  __annotations__["{n}"] = Tensor' for n in named]
    with zipfile.ZipFile(path, "w") as z:
        for key, arr in enumerate(named.values()):
            _zip_write(z, f"{stem}/data/{key}", arr.tobytes())
        _zip_write(z, f"{stem}/data.pkl", bytes(p.out))
        _zip_write(z, f"{stem}/code/__torch__.py", ("\n".join(code) + "\n").encode(), compress=True)
        _zip_write(z, f"{stem}/constants.pkl", b"\x80\x02).")
        _zip_write(z, f"{stem}/version", b"3\n")
        _zip_write(z, f"{stem}/byteorder", b"little")


# ------------------------------------------------------------------------------------------ export
def _block_starts():
    # base65536/src/lib.rs:2-24 BLOCK_START: 256 code-point blocks of 256, as runs (first, last) in steps of 256
    runs = ((13312, 19456), (19968, 40448), (41216, 41728), (42240, 42240), (67072, 67072), (73728, 74240), (77824, 78592),
            (82944, 83200), (92160, 92416), (131072, 165120))
    out = []
    for a, b in runs:
        out.extend(range(a, b + 1, 256))
    assert len(out) == 256
    return out


BLOCK_START = _block_starts()
_BLOCK_INDEX = {v: i for i, v in enumerate(BLOCK_START)}
_PADDING_BLOCK = 5376


def base65536_encode(data: bytes) -> str:
    """base65536/src/lib.rs:26-39: two bytes per code point (low byte + block of the high byte); an odd tail byte
    goes into the padding block."""
    data = bytes(data)
    out = []
    for i in range(0, len(data), 2):
        hi = BLOCK_START[data[i + 1]] if i + 1 < len(data) else _PADDING_BLOCK
        out.append(chr(hi + data[i]))
    return "".join(out)


def base65536_decode(text: str) -> bytes:
    """base65536/src/lib.rs:41-56."""
    out = bytearray()
    for ch in text:
        cp = ord(ch)
        b1 = cp & 0xFF
        out.append(b1)
        if cp - b1 != _PADDING_BLOCK:
            if cp - b1 not in _BLOCK_INDEX:
                raise ValueError(f"code point U+{cp:04X} is not base65536")  # the reference unwraps a None here (panic)
            out.append(_BLOCK_INDEX[cp - b1])
    return bytes(out)


def f32_to_bf16(values) -> np.ndarray:
    """export/src/main.rs:8-26 on an array: NaN keeps its high mantissa half with the quiet bit set; otherwise round to
    nearest, ties to even."""
    x = np.ascontiguousarray(values, dtype=np.float32).reshape(-1).view(np.uint32)
    hi = (x >> np.uint32(16)).astype(np.uint32)
    nan = (x & np.uint32(0x7FFFFFFF)) > np.uint32(0x7F800000)
    round_bit = np.uint32(0x8000)
    up = ((x & round_bit) != 0) & ((x & np.uint32(3 * 0x8000 - 1)) != 0)
    out = np.where(nan, hi | np.uint32(0x0040), hi + up.astype(np.uint32))
    return (out & np.uint32(0xFFFF)).astype(np.uint16)


def serialize_tensor(values, kind: str = "bf16") -> str:
    """export/src/main.rs:28-42 (kind="bf16": what `export` writes) or the big-endian f32 stream slimnn::loading
    decodes (kind="f32", slimnn/src/loading.rs:3-16)."""
    if kind == "bf16":
        raw = f32_to_bf16(values).astype(">u2").tobytes()
    elif kind == "f32":
        raw = np.ascontiguousarray(values, dtype=np.float32).reshape(-1).astype(">f4").tobytes()
    else:
        raise ValueError("kind must be 'bf16' or 'f32'")
    return base65536_encode(raw)


def load_nd(params: str, shape, kind: str = "f32") -> np.ndarray:
    """slimnn::loading::load_1d / load_2d / load_4d (loading.rs:18-39): decode, reinterpret, check the element count
    (the reference asserts), copy row-major.  kind="bf16" reads what `export` actually wrote."""
    raw = base65536_decode(params)
    if kind == "f32":
        if len(raw) % 4:
            raise ValueError("byte count is not a multiple of 4")  # loading.rs:5 assert
        vals = np.frombuffer(raw, dtype=">f4").astype(np.float32)
    elif kind == "bf16":
        if len(raw) % 2:
            raise ValueError("byte count is not a multiple of 2")
        vals = (np.frombuffer(raw, dtype=">u2").astype(np.uint32) << np.uint32(16)).view(np.float32)
    else:
        raise ValueError("kind must be 'bf16' or 'f32'")
    shape = tuple(int(s) for s in np.atleast_1d(shape))
    if vals.size != int(np.prod(shape)):
        raise ValueError(f"expected {int(np.prod(shape))} values, decoded {vals.size}")  # loading.rs:21 assert_eq
    return vals.reshape(shape).copy()


def export_parameters(named_tensors, path=None, kind: str = "bf16") -> str:
    """The file `export <varstore.ot> <dst>` writes (export/src/main.rs:44-92): one `load_{n}d` statement per weight
    and `load_1d` per bias, layer names sorted, then the PARAMETERS string table."""
    named = {str(k): np.asarray(v, dtype=np.float32) for k, v in named_tensors.items()}
    names = sorted({k.split(".")[0] for k in named})
    lines, i = [], 0
    for name in names:
        w = named[f"{name}.weight"]
        lines.append(f"load_{w.ndim}d(&mut policy.{name}.weight, String::from(PARAMETERS[{i}]));\n")
        lines.append(f"load_1d(&mut policy.{name}.bias, String::from(PARAMETERS[{i + 1}]));\n")
        i += 2
    lines.append(f"const PARAMETERS: [&'static str; {len(named)}] = [\n")
    i = 0
    for name in names:
        sw = serialize_tensor(named[f"{name}.weight"], kind)
        sb = serialize_tensor(named[f"{name}.bias"], kind)
        lines.append(f"// {name} - {i}\n\"{sw}\",\n\"{sb}\",\n")
        i += 2
    lines.append("];\n")
    text = "".join(lines)
    if path is not None:
        with open(path, "w", encoding="utf-8") as f:
            f.write(text)
    return text


def parse_parameters(text: str):
    """The PARAMETERS strings of an exported file, in order."""
    body = text[text.index("const PARAMETERS"):]
    return [ln.strip()[1:-2] for ln in body.splitlines() if ln.strip().startswith('"')]
