"""A minimal pure-Python HDF5 reader and writer: just enough of the format to exchange Keras-3 `.weights.h5`
checkpoints (Brats.py:732,933,1195) without h5py / libhdf5, neither of which exists in this image.

Scope (HDF5 File Format Specification, version 3.0), i.e. what `h5py.File(path, "w")` with default library bounds
writes for plain numeric arrays, which is what Keras' H5IOStore does for `save_weights`:
  * superblock version 0 or 1 (8-byte offsets and lengths);
  * old-style groups: object header v1 with a Symbol Table message (0x0011) -> B-tree v1 ("TREE", node type 0) ->
    symbol table nodes ("SNOD") -> names in a local heap ("HEAP");
  * datasets: object header v1 (continuation blocks followed) with Dataspace v1/v2 (0x0001), Datatype (0x0003:
    fixed-point and IEEE floating-point, either byte order) and Data Layout v3 (0x0008) contiguous or compact.
Chunked / compressed datasets, new-style (link-message / fractal-heap) groups, object header v2 and superblocks
2 / 3 are recognised and rejected with a clear error.  Attributes and every other message type are skipped.

Honesty note: there is no HDF5 library in this image to cross-check against, so the reader is validated against
files made by the writer below, and the writer follows the same specification sections (III.A superblock, III.B
B-trees, III.C symbol table nodes, III.D local heaps, IV.A.1.a object header v1, IV.A.2.b/d/i message layouts).
The committed fixture `tests/golden/keras3_hippocampus_n8.weights.h5` is such a file.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF

MSG_NIL, MSG_DATASPACE, MSG_LINKINFO, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL = 0x0, 0x1, 0x2, 0x3, 0x4, 0x5
MSG_LINK, MSG_LAYOUT, MSG_CONTINUATION, MSG_SYMBOL_TABLE = 0x6, 0x8, 0x10, 0x11


class H5FormatError(RuntimeError):
    pass


# ---------------------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------------------
class H5Reader:
    """Reads every numeric dataset of a file: `datasets()` -> {"/group/.../name": ndarray}."""

    def __init__(self, data: bytes):
        self.b = data
        if data[:8] != SIGNATURE:
            raise H5FormatError("not an HDF5 file (signature missing at offset 0)")
        ver = data[8]
        if ver not in (0, 1):
            raise H5FormatError(f"superblock version {ver}: only versions 0 / 1 (h5py default) are supported")
        if data[13] != 8 or data[14] != 8:
            raise H5FormatError("only 8-byte offsets / lengths are supported")
        p = 24 + (4 if ver == 1 else 0)          # after sizes, group K values and consistency flags
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", data, p)
        p += 32
        # root group symbol table entry
        _name_off, self.root_header, cache_type, _ = struct.unpack_from("<QQII", data, p)
        self.root_scratch = struct.unpack_from("<QQ", data, p + 24) if cache_type == 1 else None

    @classmethod
    def open(cls, path: str) -> "H5Reader":
        with open(path, "rb") as f:
            return cls(f.read())

    # ---- object headers ---------------------------------------------------------------------------------
    def _messages(self, addr: int) -> List[Tuple[int, bytes]]:
        b = self.b
        a = self.base + addr
        if b[a:a + 4] == b"OHDR":
            raise H5FormatError("object header version 2 (file written with libver='latest'): not supported")
        version, _, nmsg, _refcount, hsize = struct.unpack_from("<BBHII", b, a)
        if version != 1:
            raise H5FormatError(f"object header version {version} at {addr:#x}")
        out: List[Tuple[int, bytes]] = []
        blocks = [(a + 16, hsize)]                # the first chunk starts after the 16-byte (8-aligned) prefix
        while blocks and len(out) < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                body = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == MSG_CONTINUATION:
                    off, length = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self.base + off, length))
                out.append((mtype, body))
        return out

    # ---- groups -----------------------------------------------------------------------------------------
    def _heap_name(self, heap_addr: int, offset: int) -> str:
        b = self.b
        a = self.base + heap_addr
        if b[a:a + 4] != b"HEAP":
            raise H5FormatError(f"local heap signature missing at {heap_addr:#x}")
        data_addr = struct.unpack_from("<Q", b, a + 24)[0]
        s = self.base + data_addr + offset
        e = b.index(b"\x00", s)
        return b[s:e].decode("utf-8")

    def _btree_entries(self, addr: int, heap_addr: int) -> Iterator[Tuple[str, int]]:
        b = self.b
        a = self.base + addr
        if b[a:a + 4] != b"TREE":
            raise H5FormatError(f"B-tree signature missing at {addr:#x}")
        node_type, level, used = struct.unpack_from("<BBH", b, a + 4)
        if node_type != 0:
            raise H5FormatError("expected a group B-tree node")
        p = a + 24                                   # signature, type, level, entries, two sibling addresses
        children = []
        for i in range(used):
            p += 8                                   # key i (heap offset of a name)
            children.append(struct.unpack_from("<Q", b, p)[0])
            p += 8
        for child in children:
            if level > 0:
                yield from self._btree_entries(child, heap_addr)
            else:
                yield from self._snod_entries(child, heap_addr)

    def _snod_entries(self, addr: int, heap_addr: int) -> Iterator[Tuple[str, int]]:
        b = self.b
        a = self.base + addr
        if b[a:a + 4] != b"SNOD":
            raise H5FormatError(f"symbol table node signature missing at {addr:#x}")
        nsym = struct.unpack_from("<H", b, a + 6)[0]
        for i in range(nsym):
            name_off, header = struct.unpack_from("<QQ", b, a + 8 + 40 * i)
            yield self._heap_name(heap_addr, name_off), header

    # ---- datasets ---------------------------------------------------------------------------------------
    @staticmethod
    def _dtype(body: bytes) -> Optional[np.dtype]:
        cls_ver, bits0 = body[0], body[1]
        size = struct.unpack_from("<I", body, 4)[0]
        klass = cls_ver & 0x0F
        order = ">" if (bits0 & 1) else "<"
        if klass == 0:                               # fixed point
            signed = bool(bits0 & 0x08)
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}")
        if klass == 1 and size in (2, 4, 8):         # IEEE floating point
            return np.dtype(f"{order}f{size}")
        return None                                  # strings, compounds, ...: not numeric arrays

    @staticmethod
    def _dims(body: bytes) -> Tuple[int, ...]:
        version, rank = body[0], body[1]
        if version == 1:
            p = 8
        elif version == 2:
            p = 4
        else:
            raise H5FormatError(f"dataspace message version {version}")
        return tuple(struct.unpack_from("<" + "Q" * rank, body, p)) if rank else ()

    def _dataset(self, msgs: Dict[int, bytes], path: str) -> Optional[np.ndarray]:
        dt = self._dtype(msgs[MSG_DATATYPE])
        if dt is None:
            return None
        dims = self._dims(msgs[MSG_DATASPACE])
        lay = msgs[MSG_LAYOUT]
        if lay[0] != 3:
            raise H5FormatError(f"{path}: data layout message version {lay[0]} (only version 3 is supported)")
        count = int(np.prod(dims)) if dims else 1
        if lay[1] == 1:                              # contiguous
            addr, size = struct.unpack_from("<QQ", lay, 2)
            if addr == UNDEF:
                return np.zeros(dims, dtype=dt.newbyteorder("="))
            raw = self.b[self.base + addr:self.base + addr + size]
        elif lay[1] == 0:                            # compact
            size = struct.unpack_from("<H", lay, 2)[0]
            raw = lay[4:4 + size]
        else:
            raise H5FormatError(f"{path}: chunked / virtual layout is not supported (Keras writes weights contiguous)")
        if len(raw) < count * dt.itemsize:
            raise H5FormatError(f"{path}: truncated data ({len(raw)} of {count * dt.itemsize} bytes)")
        return np.frombuffer(raw, dtype=dt, count=count).reshape(dims).astype(dt.newbyteorder("="))

    # ---- walk -------------------------------------------------------------------------------------------
    def _walk(self, header: int, path: str, out: Dict[str, np.ndarray], seen: set) -> None:
        if header in seen:
            return
        seen.add(header)
        msgs: Dict[int, bytes] = {}
        for t, body in self._messages(header):
            msgs.setdefault(t, body)
        if MSG_SYMBOL_TABLE in msgs:
            btree, heap = struct.unpack_from("<QQ", msgs[MSG_SYMBOL_TABLE], 0)
            for name, child in self._btree_entries(btree, heap):
                self._walk(child, f"{path}/{name}" if path != "/" else f"/{name}", out, seen)
        elif MSG_LINKINFO in msgs or MSG_LINK in msgs:
            raise H5FormatError(f"{path}: new-style group (link messages); re-save with h5py's default libver")
        elif MSG_LAYOUT in msgs and MSG_DATASPACE in msgs and MSG_DATATYPE in msgs:
            arr = self._dataset(msgs, path)
            if arr is not None:
                out[path] = arr

    def datasets(self) -> Dict[str, np.ndarray]:
        out: Dict[str, np.ndarray] = {}
        self._walk(self.root_header, "/", out, set())
        return out


def read_h5(path: str) -> Dict[str, np.ndarray]:
    """{"/group/.../dataset": ndarray} of every numeric dataset in the file."""
    return H5Reader.open(path).datasets()


# ---------------------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------------------
def _pad8(n: int) -> int:
    return (n + 7) & ~7


class _Writer:
    GROUP_LEAF_K, GROUP_INTERNAL_K = 4, 16

    def __init__(self):
        self.buf = bytearray(96)                    # superblock v0 is filled in at the end

    def alloc(self, data: bytes) -> int:
        addr = _pad8(len(self.buf))
        self.buf.extend(b"\x00" * (addr - len(self.buf)))
        self.buf.extend(data)
        return addr

    @staticmethod
    def _message(mtype: int, body: bytes) -> bytes:
        body = body + b"\x00" * (_pad8(len(body)) - len(body))
        return struct.pack("<HHB3x", mtype, len(body), 0) + body

    def _object_header(self, messages: List[bytes]) -> int:
        data = b"".join(messages)
        return self.alloc(struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(data)) + data)

    def dataset(self, arr: np.ndarray) -> int:
        arr = np.asarray(arr, order="C")             # (ascontiguousarray would turn a scalar into shape (1,))
        if arr.dtype.kind == "f" and arr.dtype.itemsize in (4, 8):
            size = arr.dtype.itemsize
            exp_bits, man_bits, bias = (8, 23, 127) if size == 4 else (11, 52, 1023)
            # class 1 (floating point) version 1; little-endian, mantissa normalisation "msb implied", sign at the top bit
            dtype = struct.pack("<BBBBI", 0x11, 0x20, 8 * size - 1, 0, size) + \
                struct.pack("<HHBBBBI", 0, 8 * size, man_bits, exp_bits, 0, man_bits, bias)
        elif arr.dtype.kind in "iu":
            size = arr.dtype.itemsize
            dtype = struct.pack("<BBBBI", 0x10, 0x08 if arr.dtype.kind == "i" else 0x00, 0, 0, size) + \
                struct.pack("<HH", 0, 8 * size)
        else:
            raise H5FormatError(f"cannot write dtype {arr.dtype}")
        raw = arr.astype(arr.dtype.newbyteorder("<")).tobytes()
        data_addr = self.alloc(raw) if raw else UNDEF
        rank = arr.ndim
        space = struct.pack("<BBB5x", 1, rank, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        fill = struct.pack("<BBBB", 2, 2, 0, 0)       # v2: allocate late, write fill at allocation, no value defined
        layout = struct.pack("<BBQQ", 3, 1, data_addr, len(raw))
        return self._object_header([self._message(MSG_DATASPACE, space), self._message(MSG_DATATYPE, dtype),
                                    self._message(MSG_FILL, fill), self._message(MSG_LAYOUT, layout)])

    def group(self, children: Dict[str, int]) -> Tuple[int, int, int]:
        """children: name -> object header address.  Returns (header, btree, heap) addresses."""
        names = sorted(children, key=lambda s: s.encode("utf-8"))
        if len(names) > 2 * self.GROUP_LEAF_K * 2 * self.GROUP_INTERNAL_K:
            raise H5FormatError("too many links in one group for this writer")
        # local heap data: the empty string at offset 0, then the names, each padded to 8 bytes
        heap_data = bytearray(8)
        offsets = {}
        for n in names:
            offsets[n] = len(heap_data)
            e = n.encode("utf-8") + b"\x00"
            heap_data.extend(e + b"\x00" * (_pad8(len(e)) - len(e)))
        heap_data_addr = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, heap_data_addr))
        # symbol table nodes of up to 2K entries each
        cap = 2 * self.GROUP_LEAF_K
        snods, keys = [], [0]
        for i in range(0, max(len(names), 1), cap):
            part = names[i:i + cap]
            body = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(part)))
            for n in part:
                body.extend(struct.pack("<QQII16x", offsets[n], children[n], 0, 0))
            body.extend(b"\x00" * (8 + 40 * cap - len(body)))
            snods.append(self.alloc(bytes(body)))
            keys.append(offsets[part[-1]] if part else 0)
        node = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF))
        for i, s in enumerate(snods):
            node.extend(struct.pack("<QQ", keys[i], s))
        node.extend(struct.pack("<Q", keys[len(snods)]))
        full = 24 + (2 * self.GROUP_INTERNAL_K + 1) * 8 + 2 * self.GROUP_INTERNAL_K * 8
        node.extend(b"\x00" * (full - len(node)))
        btree = self.alloc(bytes(node))
        header = self._object_header([self._message(MSG_SYMBOL_TABLE, struct.pack("<QQ", btree, heap))])
        return header, btree, heap

    def finish(self, root: Tuple[int, int, int]) -> bytes:
        header, btree, heap = root
        eof = _pad8(len(self.buf))
        self.buf.extend(b"\x00" * (eof - len(self.buf)))
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.GROUP_LEAF_K, self.GROUP_INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, header, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_h5(path: str, datasets: Dict[str, np.ndarray]) -> None:
    """Write {"/a/b/name": ndarray} as an HDF5 file of old-style groups and contiguous datasets."""
    tree: dict = {}
    for p, arr in datasets.items():
        parts = [s for s in p.split("/") if s]
        if not parts:
            raise H5FormatError("empty dataset path")
        node = tree
        for s in parts[:-1]:
            node = node.setdefault(s, {})
            if not isinstance(node, dict):
                raise H5FormatError(f"{p}: a dataset is used as a group")
        node[parts[-1]] = np.asarray(arr)
    w = _Writer()

    def emit(node) -> int:
        if isinstance(node, dict):
            return w.group({k: emit(v) for k, v in node.items()})[0]
        return w.dataset(node)

    children = {k: emit(v) for k, v in tree.items()}
    data = w.finish(w.group(children))
    with open(path, "wb") as f:
        f.write(data)
