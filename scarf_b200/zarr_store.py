"""A minimal Zarr v2 directory store: exactly what the make_graph / run_mapping contract needs.

The reference persists every stage through `zarr` (<= 2.16) + numcodecs Blosc (`scarf/writers.py:58-89`); neither is
installed where this runs.  This module writes and reads the same on-disk format -- `.zgroup` / `.zarray` /
`.zattrs` JSON plus one file per chunk, C order, little endian.  New arrays are written with `compressor: null`,
which real zarr opens unchanged (readers take the codec from `.zarray`, `graph_datastore.py:474-511`).  Arrays the
reference wrote (`numcodecs.Blosc(cname='lz4', clevel=5, shuffle=...)`, `scarf/writers.py:79-89`) are read through
the native frame decoder of the C-ABI (`scf_host_blosc_decode`, scarf_b200/csrc/host_blosc.cu); a partial write into
such an array stores the chunk as a Blosc "memcpy" frame, which c-blosc reads like any other.  Chunk shapes follow
the reference's `create_zarr_dataset` calls so that chunk-wise consumers see the same blocking.
"""
from __future__ import annotations

import ctypes
import json
import os
import shutil
import struct

import numpy as np

from . import lib


def blosc_decode(frame: bytes) -> np.ndarray:
    """One Blosc-1 frame -> its bytes (uint8 array), decoded by the library's host routine."""
    src = np.frombuffer(frame, dtype=np.uint8)
    nbytes = ctypes.c_int64()
    lib.call("scf_host_blosc_info", src.ctypes.data, src.size, ctypes.addressof(nbytes), None, None, launches=0)
    out = np.empty(nbytes.value, dtype=np.uint8)
    lib.call("scf_host_blosc_decode", src.ctypes.data, src.size, out.ctypes.data, out.size, launches=0)
    return out


def blosc_store(raw: bytes, typesize: int) -> bytes:
    """A Blosc-1 frame that carries `raw` uncompressed (flag 0x02, "memcpyed"): version 2, LZ4 format id."""
    return struct.pack("<BBBBIII", 2, 1, 0x02 | (1 << 5), min(typesize, 255), len(raw), len(raw), len(raw) + 16) + raw


def _dtype_str(dt: np.dtype) -> str:
    dt = np.dtype(dt)
    if dt.kind == "U":
        return f"<U{dt.itemsize // 4}"
    if dt.kind == "b":
        return "|b1"
    if dt.itemsize == 1:
        return f"|{dt.kind}1"
    return f"<{dt.kind}{dt.itemsize}"


class Attrs:
    """`.zattrs`: a dict persisted on every assignment (zarr's Attributes behave the same way)."""

    def __init__(self, path):
        self._file = os.path.join(path, ".zattrs")

    def _load(self):
        if os.path.exists(self._file):
            with open(self._file) as f:
                return json.load(f)
        return {}

    def __getitem__(self, k):
        return self._load()[k]

    def __contains__(self, k):
        return k in self._load()

    def get(self, k, default=None):
        return self._load().get(k, default)

    def __setitem__(self, k, v):
        d = self._load()
        d[k] = v
        with open(self._file, "w") as f:
            json.dump(d, f, indent=4, sort_keys=True)

    def asdict(self):
        return self._load()


class Array:
    def __init__(self, path):
        self.path = path
        with open(os.path.join(path, ".zarray")) as f:
            meta = json.load(f)
        comp = meta.get("compressor")
        if comp is not None and comp.get("id") != "blosc":
            raise NotImplementedError(f"{path}: compressor {comp.get('id')!r}; Scarf stores use Blosc (or none)")
        if meta.get("filters"):
            raise NotImplementedError(f"{path}: Zarr filters are not supported")
        if meta.get("order", "C") != "C":
            raise NotImplementedError(f"{path}: only C-ordered chunks are supported")
        self.blosc = comp is not None
        self.shape = tuple(meta["shape"])
        self.chunks = tuple(meta["chunks"])
        self.dtype = np.dtype(meta["dtype"])
        self.fill_value = meta.get("fill_value", 0)
        self.attrs = Attrs(path)

    @property
    def ndim(self):
        return len(self.shape)

    def _grid(self):
        return tuple((s + c - 1) // c for s, c in zip(self.shape, self.chunks))

    def _chunk_file(self, idx):
        return os.path.join(self.path, ".".join(str(i) for i in idx))

    def _fill(self):
        return np.full(self.chunks, self.fill_value if self.dtype.kind != "U" else "", dtype=self.dtype)

    def read_chunk(self, idx):
        """The whole chunk `idx` (a tuple) in chunk shape; a missing file reads as the fill value (zarr semantics)."""
        f = self._chunk_file(idx)
        if not os.path.exists(f):
            return self._fill()
        if not self.blosc:
            return np.fromfile(f, dtype=self.dtype).reshape(self.chunks)
        with open(f, "rb") as fh:
            raw = blosc_decode(fh.read())
        return raw.view(self.dtype).reshape(self.chunks)

    def _write_chunk(self, idx, chunk):
        if not self.blosc:
            chunk.tofile(self._chunk_file(idx))
            return
        with open(self._chunk_file(idx), "wb") as fh:
            fh.write(blosc_store(np.ascontiguousarray(chunk).tobytes(), self.dtype.itemsize))

    def __setitem__(self, key, value):
        """Whole-array (`a[:] = x`) or leading-axis row range (`a[lo:hi] = x`, `a[lo:hi, :] = x`) assignment."""
        lo, hi = self._rows(key)
        value = np.asarray(value, dtype=self.dtype)
        if value.ndim == 0:
            value = np.full((hi - lo,) + self.shape[1:], value, dtype=self.dtype)
        if value.shape != (hi - lo,) + self.shape[1:]:
            raise ValueError(f"shape mismatch: {value.shape} into rows [{lo}, {hi}) of {self.shape}")
        c0 = self.chunks[0]
        grid = self._grid()
        for ci in range(lo // c0, (hi + c0 - 1) // c0 if hi > lo else lo // c0):
            r0, r1 = ci * c0, min((ci + 1) * c0, self.shape[0])
            a, b = max(lo, r0), min(hi, r1)
            for cj in range(grid[1] if self.ndim == 2 else 1):
                idx = (ci, cj) if self.ndim == 2 else (ci,)
                full = (a == r0 and b == r1)
                chunk = self._fill() if full else np.array(self.read_chunk(idx))
                if self.ndim == 2:
                    k0, k1 = cj * self.chunks[1], min((cj + 1) * self.chunks[1], self.shape[1])
                    chunk[a - r0:b - r0, : k1 - k0] = value[a - lo:b - lo, k0:k1]
                else:
                    chunk[a - r0:b - r0] = value[a - lo:b - lo]
                self._write_chunk(idx, chunk)

    def _rows(self, key):
        if isinstance(key, tuple):
            if len(key) > 1 and any(k != slice(None) for k in key[1:]):
                raise NotImplementedError("only leading-axis ranges are supported")
            key = key[0]
        if key is Ellipsis:
            key = slice(None)
        if not isinstance(key, slice) or key.step not in (None, 1):
            raise NotImplementedError("only contiguous leading-axis ranges are supported")
        lo, hi, _ = key.indices(self.shape[0])
        return lo, hi

    def __getitem__(self, key):
        lo, hi = self._rows(key)
        out = np.empty((hi - lo,) + self.shape[1:], dtype=self.dtype)
        c0 = self.chunks[0]
        grid = self._grid()
        for ci in range(lo // c0, (hi + c0 - 1) // c0 if hi > lo else lo // c0):
            r0, r1 = ci * c0, min((ci + 1) * c0, self.shape[0])
            a, b = max(lo, r0), min(hi, r1)
            for cj in range(grid[1] if self.ndim == 2 else 1):
                idx = (ci, cj) if self.ndim == 2 else (ci,)
                chunk = self.read_chunk(idx)
                if self.ndim == 2:
                    k0, k1 = cj * self.chunks[1], min((cj + 1) * self.chunks[1], self.shape[1])
                    out[a - lo:b - lo, k0:k1] = chunk[a - r0:b - r0, : k1 - k0]
                else:
                    out[a - lo:b - lo] = chunk[a - r0:b - r0]
        return out


class Group:
    def __init__(self, path, create=False):
        self.path = path
        if create:
            os.makedirs(path, exist_ok=True)
            zg = os.path.join(path, ".zgroup")
            if not os.path.exists(zg):
                with open(zg, "w") as f:
                    json.dump({"zarr_format": 2}, f)
        elif not os.path.exists(os.path.join(path, ".zgroup")):
            raise KeyError(path)
        self.attrs = Attrs(path)

    def _p(self, name):
        return os.path.join(self.path, *[x for x in name.split("/") if x])

    def __contains__(self, name):
        p = self._p(name)
        return os.path.exists(os.path.join(p, ".zgroup")) or os.path.exists(os.path.join(p, ".zarray"))

    def __getitem__(self, name):
        p = self._p(name)
        if os.path.exists(os.path.join(p, ".zarray")):
            return Array(p)
        if os.path.exists(os.path.join(p, ".zgroup")):
            return Group(p)
        raise KeyError(name)

    def keys(self):
        return sorted(x for x in os.listdir(self.path) if not x.startswith(".") and x in self)

    def create_group(self, name, overwrite=False):
        p = self._p(name)
        if overwrite and os.path.exists(p):
            shutil.rmtree(p)
        parts = [x for x in name.split("/") if x]
        for i in range(1, len(parts) + 1):  # every intermediate level is a group, like zarr's require_group
            Group(os.path.join(self.path, *parts[:i]), create=True)
        return Group(p)

    def create_dataset(self, name, shape, dtype, chunks, overwrite=True):
        """`create_zarr_dataset` (scarf/writers.py:58-89): a 1-tuple `chunks` on a 2-D shape is completed to full
        width, exactly as zarr does."""
        shape = tuple(int(s) for s in shape)
        chunks = tuple(int(c) for c in chunks)
        if len(chunks) < len(shape):
            chunks = chunks + shape[len(chunks):]
        chunks = tuple(max(1, c) for c in chunks)
        p = self._p(name)
        parent = os.path.dirname(p)
        if parent != self.path and not os.path.exists(os.path.join(parent, ".zgroup")):
            self.create_group(os.path.relpath(parent, self.path))
        if os.path.exists(p):
            if not overwrite:
                raise ValueError(f"{name} exists")
            shutil.rmtree(p)
        os.makedirs(p)
        dt = np.dtype(dtype)
        meta = {"chunks": list(chunks), "compressor": None, "dtype": _dtype_str(dt),
                "fill_value": "" if dt.kind == "U" else (False if dt.kind == "b" else 0), "filters": None,
                "order": "C", "shape": list(shape), "zarr_format": 2}
        with open(os.path.join(p, ".zarray"), "w") as f:
            json.dump(meta, f, indent=4, sort_keys=True)
        return Array(p)


def open_group(path, mode="a"):
    """`zarr.open(path, mode)` for a directory store: 'r' / 'r+' need an existing store, 'w' starts empty."""
    if mode == "w" and os.path.exists(path):
        shutil.rmtree(path)
    if mode in ("r", "r+") and not os.path.exists(os.path.join(path, ".zgroup")):
        raise FileNotFoundError(path)
    return Group(path, create=mode in ("a", "w"))
