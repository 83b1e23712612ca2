"""Minimal reader/writer for the Keras-2.x HDF5 checkpoints the reference ships.

The reference loads its models with ``tf.keras.models.load_model(path)``
(``BlazePoser/blazeFaceDetectorH5.py:102``, ``Model-96/test.py:21``,
``JoinModels.py:29-31``).  Neither TensorFlow nor h5py exist in this environment,
so this module understands exactly the subset of HDF5 those files use:

* superblock version 0, 8-byte offsets / lengths,
* version-1 object headers with continuation blocks,
* "old style" groups (symbol table message -> v1 B-tree -> SNOD -> local heap),
* contiguous (layout v3 class 1) or compact little-endian datasets,
* attributes stored as object-header messages (used for ``weight_names`` etc.);
  the large ``model_config`` / ``training_config`` JSON attributes are located by
  content because Keras writes them as one contiguous byte string.

It also writes the same subset (``write_h5``) so that ``Model.save(path)`` /
``ModelCheckpoint`` outputs can be produced without h5py.

Nothing in here touches the GPU.
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


class _Dataset:
    __slots__ = ("shape", "dtype", "offset", "nbytes", "compact")

    def __init__(self):
        self.shape: Tuple[int, ...] = ()
        self.dtype = None
        self.offset = None
        self.nbytes = 0
        self.compact: Optional[bytes] = None


class H5File:
    """Read-only view of one Keras ``.h5`` file held in memory."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.buf = f.read()
        self.path = path
        b = self.buf
        if b[:8] != _SIG:
            raise H5FormatError(f"{path}: not an HDF5 file")
        if b[8] != 0:
            raise H5FormatError(f"{path}: superblock version {b[8]} unsupported (need 0)")
        if b[13] != 8 or b[14] != 8:
            raise H5FormatError(f"{path}: only 8-byte offsets/lengths supported")
        # root group symbol-table entry sits at byte 56 (8 sig + 16 header + 4 addresses)
        self.root_header = struct.unpack_from("<Q", b, 56 + 8)[0]
        self._datasets: Optional[Dict[str, _Dataset]] = None
        self._attrs: Dict[str, Dict[str, object]] = {}

    # ------------------------------------------------------------------ object headers
    def _messages(self, addr: int) -> Iterator[Tuple[int, int, int]]:
        """Yield (type, data_offset, size) for every message of a v1 object header."""
        b = self.buf
        if b[addr] != 1:
            raise H5FormatError(f"object header v{b[addr]} at {addr} unsupported")
        nmsg = struct.unpack_from("<H", b, addr + 2)[0]
        hsize = struct.unpack_from("<I", b, addr + 8)[0]
        blocks = [(addr + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and seen < nmsg:
                mtype, msize = struct.unpack_from("<HH", b, pos)
                data = pos + 8
                seen += 1
                if mtype == 0x10:  # continuation
                    off, ln = struct.unpack_from("<QQ", b, data)
                    blocks.append((off, ln))
                else:
                    yield mtype, data, msize
                pos = data + msize

    # ------------------------------------------------------------------ groups
    def _heap_string(self, heap_addr: int, off: int) -> str:
        b = self.buf
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5FormatError("bad local heap")
        data_addr = struct.unpack_from("<Q", b, heap_addr + 24)[0]
        start = data_addr + off
        end = b.index(b"\x00", start)
        return b[start:end].decode("utf-8")

    def _btree_leaves(self, addr: int) -> Iterator[int]:
        b = self.buf
        if b[addr:addr + 4] != b"TREE":
            raise H5FormatError("bad B-tree node")
        level = b[addr + 5]
        used = struct.unpack_from("<H", b, addr + 6)[0]
        pos = addr + 24  # signature(4) type(1) level(1) used(2) left(8) right(8)
        for i in range(used):
            child = struct.unpack_from("<Q", b, pos + 8 + i * 16)[0]
            if level == 0:
                yield child
            else:
                yield from self._btree_leaves(child)

    def _children(self, header_addr: int) -> List[Tuple[str, int]]:
        out: List[Tuple[str, int]] = []
        for mtype, data, _ in self._messages(header_addr):
            if mtype != 0x11:
                continue
            btree, heap = struct.unpack_from("<QQ", self.buf, data)
            for snod in self._btree_leaves(btree):
                b = self.buf
                if b[snod:snod + 4] != b"SNOD":
                    raise H5FormatError("bad symbol node")
                n = struct.unpack_from("<H", b, snod + 6)[0]
                for i in range(n):
                    e = snod + 8 + 40 * i
                    name_off, obj = struct.unpack_from("<QQ", b, e)
                    out.append((self._heap_string(heap, name_off), obj))
        return out

    # ------------------------------------------------------------------ datatypes
    @staticmethod
    def _dtype(b: bytes, pos: int):
        cls = b[pos] & 0x0F
        bits0 = b[pos + 1]
        size = struct.unpack_from("<I", b, pos + 4)[0]
        if bits0 & 1:
            raise H5FormatError("big-endian data unsupported")
        if cls == 1:
            return np.dtype(f"<f{size}")
        if cls == 0:
            signed = (bits0 >> 3) & 1
            return np.dtype(f"<{'i' if signed else 'u'}{size}")
        if cls == 3:  # fixed-length string
            return np.dtype(f"S{size}")
        if cls == 9:  # variable length (strings) - caller handles
            return "vlen"
        raise H5FormatError(f"datatype class {cls} unsupported")

    @staticmethod
    def _dataspace(b: bytes, pos: int) -> Tuple[int, ...]:
        ver, rank, flags = b[pos], b[pos + 1], b[pos + 2]
        if ver == 1:
            dpos = pos + 8
        elif ver == 2:
            dpos = pos + 4
        else:
            raise H5FormatError(f"dataspace v{ver} unsupported")
        return tuple(struct.unpack_from("<Q", b, dpos + 8 * i)[0] for i in range(rank))

    def _parse_object(self, addr: int, path: str, out: Dict[str, _Dataset]):
        b = self.buf
        ds = _Dataset()
        is_group = False
        has_layout = False
        attrs: Dict[str, object] = {}
        for mtype, data, size in self._messages(addr):
            if mtype == 0x11:
                is_group = True
            elif mtype == 0x01:
                ds.shape = self._dataspace(b, data)
            elif mtype == 0x03:
                ds.dtype = self._dtype(b, data)
            elif mtype == 0x08:
                ver = b[data]
                if ver != 3:
                    raise H5FormatError(f"layout v{ver} unsupported")
                lclass = b[data + 1]
                if lclass == 1:
                    ds.offset, ds.nbytes = struct.unpack_from("<QQ", b, data + 2)
                elif lclass == 0:
                    n = struct.unpack_from("<H", b, data + 2)[0]
                    ds.compact = b[data + 4:data + 4 + n]
                    ds.nbytes = n
                else:
                    raise H5FormatError("chunked datasets unsupported")
                has_layout = True
            elif mtype == 0x0C:
                try:
                    k, v = self._attribute(data)
                    attrs[k] = v
                except H5FormatError:
                    pass
        if attrs:
            self._attrs[path or "/"] = attrs
        if is_group:
            for name, child in self._children(addr):
                self._parse_object(child, f"{path}/{name}", out)
        elif has_layout:
            out[path] = ds

    def _attribute(self, pos: int) -> Tuple[str, object]:
        b = self.buf
        ver = b[pos]
        if ver not in (1, 2, 3):
            raise H5FormatError("attribute version")
        name_sz, dt_sz, ds_sz = struct.unpack_from("<HHH", b, pos + 2)
        p = pos + 8
        if ver == 3:
            p += 1

        def pad(n):
            return (n + 7) & ~7 if ver == 1 else n

        name = b[p:p + name_sz].split(b"\x00")[0].decode()
        p += pad(name_sz)
        dt = self._dtype(b, p)
        dtpos = p
        p += pad(dt_sz)
        shape = self._dataspace(b, p) if ds_sz >= 4 else ()
        p += pad(ds_sz)
        if isinstance(dt, str):  # vlen: not needed by callers (global-heap strings)
            raise H5FormatError("vlen attribute")
        count = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(b, dtype=dt, count=count, offset=p)
        if dt.kind == "S":
            vals = [x.split(b"\x00")[0].decode() for x in arr.tolist()]
            return name, (vals if shape else vals[0])
        return name, (arr.reshape(shape).copy() if shape else arr[0])

    # ------------------------------------------------------------------ public API
    def datasets(self) -> Dict[str, _Dataset]:
        if self._datasets is None:
            out: Dict[str, _Dataset] = {}
            self._parse_object(self.root_header, "", out)
            self._datasets = out
        return self._datasets

    def attrs(self, path: str = "/") -> Dict[str, object]:
        self.datasets()
        return self._attrs.get(path, {})

    def read(self, path: str) -> np.ndarray:
        ds = self.datasets()[path]
        if ds.compact is not None:
            arr = np.frombuffer(ds.compact, dtype=ds.dtype)
        elif ds.offset == _UNDEF or ds.nbytes == 0:
            arr = np.zeros(int(np.prod(ds.shape)), dtype=ds.dtype)
        else:
            arr = np.frombuffer(self.buf, dtype=ds.dtype,
                                count=int(np.prod(ds.shape)) if ds.shape else 1,
                                offset=ds.offset)
        return arr.reshape(ds.shape).copy()

    def weights(self) -> Dict[str, np.ndarray]:
        """All arrays under ``/model_weights`` keyed by the path below it."""
        pre = "/model_weights/"
        return {k[len(pre):]: self.read(k) for k in self.datasets() if k.startswith(pre)}

    def _json_blob(self, marker: bytes) -> Optional[dict]:
        i = self.buf.find(marker)
        if i < 0:
            return None
        end = self.buf.index(b"\x00", i)
        obj, _ = json.JSONDecoder().raw_decode(self.buf[i:end].decode("utf-8"))
        return obj

    def model_config(self) -> dict:
        cfg = self._json_blob(b'{"class_name"')
        if cfg is None:
            raise H5FormatError(f"{self.path}: no model_config JSON found")
        return cfg

    def training_config(self) -> Optional[dict]:
        return self._json_blob(b'{"loss"')


# ====================================================================== writer
class _Writer:
    """Emits the same HDF5 subset the reader accepts (superblock v0, v1 headers)."""

    def __init__(self):
        self.buf = bytearray()

    def tell(self):
        return len(self.buf)

    def align(self, n=8):
        while len(self.buf) % n:
            self.buf.append(0)

    def put(self, data: bytes) -> int:
        self.align()
        off = len(self.buf)
        self.buf += data
        return off

    def patch(self, off: int, data: bytes):
        self.buf[off:off + len(data)] = data


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    pad = (-len(body)) % 8
    body = body + b"\x00" * pad
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _dataspace_msg(shape) -> bytes:
    rank = len(shape)
    body = struct.pack("<BBB5x", 1, rank, 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)
    return body


def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f":
        size = dt.itemsize
        if size == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits = bytes([0x20, 31, 0])
        elif size == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits = bytes([0x20, 63, 0])
        else:
            raise H5FormatError("float size")
        return bytes([0x11]) + bits + struct.pack("<I", size) + props
    if dt.kind in "iu":
        bits = bytes([0x08 if dt.kind == "i" else 0x00, 0, 0])
        return bytes([0x10]) + bits + struct.pack("<I", dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        return bytes([0x13]) + bytes([0, 0, 0]) + struct.pack("<I", dt.itemsize)
    raise H5FormatError(f"dtype {dt} unsupported by writer")


def _attr_msg(name: str, value) -> bytes:
    if isinstance(value, (bytes, str)):
        raw = value.encode() if isinstance(value, str) else value
        arr = np.array(raw + b"\x00", dtype=f"S{len(raw) + 1}")
        shape = ()
    elif isinstance(value, (list, tuple)) and value and isinstance(value[0], (str, bytes)):
        enc = [v.encode() if isinstance(v, str) else v for v in value]
        width = max(len(e) for e in enc) + 1
        arr = np.array(enc, dtype=f"S{width}")
        shape = arr.shape
    else:
        arr = np.asarray(value)
        shape = arr.shape
    nm = name.encode() + b"\x00"
    dt = _dtype_msg(arr.dtype)
    ds = _dataspace_msg(shape)

    def pad(x):
        return x + b"\x00" * ((-len(x)) % 8)

    body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + pad(nm) + pad(dt) + pad(ds) + arr.tobytes()
    return body


def _object_header(w: _Writer, msgs: List[bytes]) -> int:
    body = b"".join(msgs)
    hdr = struct.pack("<BxHII4x", 1, len(msgs), 1, len(body))
    return w.put(hdr + body)


def _write_group(w: _Writer, tree: dict, attrs_of: Dict[int, dict]) -> int:
    """tree: name -> (np.ndarray | dict). Returns object-header address."""
    entries: List[Tuple[str, int]] = []
    for name in sorted(tree):
        node = tree[name]
        if isinstance(node, dict):
            addr = _write_group(w, node, attrs_of)
        else:
            arr = np.ascontiguousarray(node)
            if arr.dtype.byteorder == ">":
                arr = arr.astype(arr.dtype.newbyteorder("<"))
            data_off = w.put(arr.tobytes()) if arr.nbytes else _UNDEF
            msgs = [
                _msg(0x01, _dataspace_msg(arr.shape)),
                _msg(0x03, _dtype_msg(arr.dtype), 1),
                _msg(0x08, struct.pack("<BBQQ", 3, 1, data_off, arr.nbytes)),
            ]
            addr = _object_header(w, msgs)
        entries.append((name, addr))
    # local heap: offset 0 holds the empty string
    heap_data = bytearray(b"\x00" * 8)
    name_offs = []
    for name, _ in entries:
        name_offs.append(len(heap_data))
        heap_data += name.encode() + b"\x00"
        while len(heap_data) % 8:
            heap_data.append(0)
    heap_data_addr = w.put(bytes(heap_data))
    heap_addr = w.put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), _UNDEF, heap_data_addr))
    # one SNOD per <=8 entries keeps leaf K small; a single level-0 B-tree node lists them
    K = 8
    snods = []
    for i in range(0, max(len(entries), 1), K):
        chunk = list(zip(name_offs[i:i + K], entries[i:i + K]))
        body = b"SNOD" + struct.pack("<BxH", 1, len(chunk))
        for off, (_, addr) in chunk:
            body += struct.pack("<QQII16x", off, addr, 0, 0)
        body += b"\x00" * (40 * (2 * K - len(chunk)))
        snods.append((w.put(body), chunk))
    node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), _UNDEF, _UNDEF)
    node += struct.pack("<Q", 0)
    for addr, chunk in snods:
        last_key = chunk[-1][0] if chunk else 0
        node += struct.pack("<QQ", addr, last_key)
    btree_addr = w.put(node)
    msgs = [_msg(0x11, struct.pack("<QQ", btree_addr, heap_addr))]
    for k, v in attrs_of.get(id(tree), {}).items():
        msgs.append(_msg(0x0C, _attr_msg(k, v)))
    return _object_header(w, msgs)


def write_h5(path: str, weights: Dict[str, np.ndarray], model_config: dict,
             training_config: Optional[dict] = None, extra_attrs: Optional[dict] = None):
    """Write ``/model_weights/<key>`` datasets plus Keras' root JSON attributes."""
    root: dict = {"model_weights": {}}
    for key, arr in weights.items():
        node = root["model_weights"]
        parts = key.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = np.asarray(arr)
    attrs_of: Dict[int, dict] = {}
    rattrs = {"backend": "tensorflow", "keras_version": "2.13.1",
              "model_config": json.dumps(model_config)}
    if training_config is not None:
        rattrs["training_config"] = json.dumps(training_config)
    if extra_attrs:
        rattrs.update(extra_attrs)
    attrs_of[id(root)] = rattrs
    layer_names = list(root["model_weights"].keys())
    attrs_of[id(root["model_weights"])] = {"layer_names": layer_names or [""],
                                           "backend": "tensorflow", "keras_version": "2.13.1"}
    w = _Writer()
    w.buf += _SIG + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0)
    w.buf += struct.pack("<QQQQ", 0, _UNDEF, 0, _UNDEF)  # base, free-space, EOF (patched), driver
    w.buf += struct.pack("<QQII16x", 0, 0, 0, 0)          # root symbol-table entry (patched)
    root_addr = _write_group(w, root, attrs_of)
    w.align()
    w.patch(56 + 8, struct.pack("<Q", root_addr))
    w.patch(24 + 16, struct.pack("<Q", len(w.buf)))
    tmp = f"{path}.tmp.{os.getpid()}"                     # a reader (or a crash) never sees a half-written checkpoint
    with open(tmp, "wb") as f:
        f.write(bytes(w.buf))
    os.replace(tmp, path)
