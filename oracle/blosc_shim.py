"""Minimal Blosc-1 / Zarr-v2 chunk reader (test infrastructure).

The reference stores every array through ``numcodecs.Blosc(cname='lz4',
clevel=5, shuffle=BITSHUFFLE)`` (scarf/writers.py:79-89); numcodecs/zarr are
not installed here, so the golden PBMC fixture is decoded with this shim:
Blosc-1 frame = 16-byte header ``<BBBBIII`` (version, versionlz, flags,
typesize, nbytes, blocksize, cbytes), ``int32 bstarts[nblocks]``, then per
block ``nsplits`` streams of ``[int32 csize][lz4 raw block]``.
"""
import json
import os
import struct

import numpy as np
import pyarrow as pa

_BYTE_SHUFFLE, _MEMCPYED, _BIT_SHUFFLE, _DONT_SPLIT = 0x01, 0x02, 0x04, 0x10


def _lz4_raw(buf: bytes, out_size: int) -> bytes:
    return pa.decompress(buf, decompressed_size=out_size, codec="lz4_raw", asbytes=True)


def _unshuffle_bytes(block: np.ndarray, typesize: int) -> np.ndarray:
    n = block.size // typesize
    head = block[: n * typesize].reshape(typesize, n).T.reshape(-1)
    return np.concatenate([head, block[n * typesize:]])


def _unshuffle_bits(block: np.ndarray, typesize: int) -> np.ndarray:
    nelem = block.size // typesize
    nelem8 = nelem - nelem % 8
    body = block[: nelem8 * typesize]
    if nelem8:
        bits = np.unpackbits(body.reshape(typesize * 8, nelem8 // 8), axis=1, bitorder="little")
        body = np.packbits(bits.T.copy(), axis=1, bitorder="little").reshape(-1)
    return np.concatenate([body, block[nelem8 * typesize:]])


def blosc_decompress(frame: bytes) -> bytes:
    _, _, flags, typesize, nbytes, blocksize, cbytes = struct.unpack("<BBBBIII", frame[:16])
    if flags & _MEMCPYED:
        return frame[16:16 + nbytes]
    nblocks = (nbytes + blocksize - 1) // blocksize
    bstarts = struct.unpack(f"<{nblocks}i", frame[16:16 + 4 * nblocks])
    out = np.empty(nbytes, dtype=np.uint8)
    for b in range(nblocks):
        bsize = min(blocksize, nbytes - b * blocksize)
        leftover = bsize != blocksize
        nsplits = typesize if (not (flags & _DONT_SPLIT) and not leftover and typesize <= 16
                               and bsize // typesize >= 128) else 1
        neblock = bsize // nsplits
        pos = bstarts[b]
        parts = []
        for _ in range(nsplits):
            (csize,) = struct.unpack("<i", frame[pos:pos + 4])
            pos += 4
            chunk = frame[pos:pos + csize]
            pos += csize
            parts.append(chunk if csize == neblock else _lz4_raw(chunk, neblock))
        block = np.frombuffer(b"".join(parts), dtype=np.uint8)
        if flags & _BIT_SHUFFLE and typesize > 0:
            block = _unshuffle_bits(block, typesize)
        elif flags & _BYTE_SHUFFLE and typesize > 1:
            block = _unshuffle_bytes(block, typesize)
        out[b * blocksize: b * blocksize + bsize] = block
    return out.tobytes()


def read_zarr_array(path: str) -> np.ndarray:
    """Read a whole Zarr-v2 array directory (C order, Blosc or no compressor)."""
    with open(os.path.join(path, ".zarray")) as fh:
        meta = json.load(fh)
    shape, chunks = meta["shape"], meta["chunks"]
    dtype = np.dtype(meta["dtype"])
    out = np.full(shape, meta["fill_value"] if meta["fill_value"] is not None else 0, dtype=dtype)
    grid = [range((s + c - 1) // c) for s, c in zip(shape, chunks)]
    import itertools
    for idx in itertools.product(*grid):
        fn = os.path.join(path, ".".join(map(str, idx)))
        if not os.path.exists(fn):
            continue
        with open(fn, "rb") as fh:
            raw = fh.read()
        if meta["compressor"] is not None:
            raw = blosc_decompress(raw)
        block = np.frombuffer(raw, dtype=dtype).reshape(chunks)
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, shape))
        out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
    return out
