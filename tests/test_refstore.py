"""Opening a store the REFERENCE wrote: Blosc chunk decoding (host routine of the C-ABI), the dense ->
CSR kernels and the first-open statistics, against tests/golden/pbmc_ref_store -- chunk files byte for byte as
numcodecs wrote them for the reference's own test fixture (tests/golden/make_pbmc_ref_store.py)."""
import ctypes
import glob
import os
import shutil

import numpy as np
import pytest

from conftest import GOLDEN

STORE = os.path.join(GOLDEN, "pbmc_ref_store")
N_GENES = 3000


def _chunk_files():
    fs = [f for f in glob.glob(os.path.join(STORE, "**", "*"), recursive=True)
          if os.path.isfile(f) and not os.path.basename(f).startswith(".")]
    assert len(fs) == 9
    return sorted(fs)


def test_blosc_decoder_equals_oracle_on_reference_chunks():
    """bit-shuffled uint32 counts (flags 0x24), byte-shuffled <U columns (0x31) and the 1-byte bool column (0x21)."""
    from oracle.blosc_shim import blosc_decompress
    from scarf_b200.zarr_store import blosc_decode

    seen = set()
    for f in _chunk_files():
        frame = open(f, "rb").read()
        seen.add(frame[2])
        assert blosc_decode(frame).tobytes() == blosc_decompress(frame), f
    assert {0x24, 0x31, 0x21} <= seen


def test_blosc_decoder_rejects_what_it_cannot_decode():
    from scarf_b200 import lib
    from scarf_b200.zarr_store import blosc_decode, blosc_store

    frame = open(os.path.join(STORE, "RNA", "counts", "0.1"), "rb").read()
    for bad in (frame[:10], frame[:200], frame[:-7]):  # short header, cut block table / body
        with pytest.raises(ValueError):
            blosc_decode(bad)
    garbled = bytearray(frame)
    garbled[400:420] = b"\xff" * 20  # inside an LZ4 stream: lengths run past the block
    with pytest.raises(ValueError):
        blosc_decode(bytes(garbled))
    zstd = bytearray(frame)
    zstd[2] = (zstd[2] & 0x1F) | (4 << 5)  # inner codec id 4 (zstd)
    with pytest.raises(ValueError, match="lz4"):
        blosc_decode(bytes(zstd))
    out = np.empty(8, dtype=np.uint8)  # destination must have the frame's size
    src = np.frombuffer(frame, dtype=np.uint8)
    assert lib.raw("scf_host_blosc_decode")(src.ctypes.data, src.size, out.ctypes.data, out.size) > 0
    assert b"nbytes" in lib.raw("scf_last_error")()
    raw = np.arange(1000, dtype=np.uint32).tobytes()  # "memcpy" frames: what a partial write into a Blosc array stores
    assert blosc_decode(blosc_store(raw, 4)).tobytes() == raw
    assert blosc_decode(blosc_store(b"", 4)).size == 0


def _blosc_encode(raw: bytes, typesize: int, shuffle: int, blocksize: int, dont_split: bool = False) -> bytes:
    """Test-side Blosc-1 encoder (frame layout as in oracle/blosc_shim.py; LZ4 blocks from pyarrow): shuffle 0 none,
    1 bytes, 2 bits.  A stream that does not shrink is stored raw (csize == its decoded size), as c-blosc does."""
    import struct

    import pyarrow as pa

    nbytes = len(raw)
    blocksize = min(blocksize, nbytes)  # c-blosc: a block is never larger than the buffer and holds whole elements
    if blocksize > typesize:
        blocksize -= blocksize % typesize
    flags = (1 << 5) | {0: 0, 1: 0x01, 2: 0x04}[shuffle] | (0x10 if dont_split else 0)
    nblocks = (nbytes + blocksize - 1) // blocksize
    body, bstarts = b"", []
    pos0 = 16 + 4 * nblocks
    for b in range(nblocks):
        blk = np.frombuffer(raw[b * blocksize:(b + 1) * blocksize], dtype=np.uint8)
        bsize = blk.size
        n = bsize // typesize
        if shuffle == 1 and typesize > 1:
            blk = np.concatenate([blk[: n * typesize].reshape(n, typesize).T.reshape(-1), blk[n * typesize:]])
        elif shuffle == 2:
            n8 = n - n % 8
            head = blk[: n8 * typesize]
            if n8:
                bits = np.unpackbits(head.reshape(n8, typesize), axis=1, bitorder="little")  # [elem, bit row]
                head = np.packbits(bits.T.copy(), axis=1, bitorder="little").reshape(-1)
            blk = np.concatenate([head, blk[n8 * typesize:]])
        leftover = bsize != blocksize
        nsplits = typesize if (not dont_split and not leftover and typesize <= 16 and bsize // typesize >= 128) else 1
        ne = bsize // nsplits
        bstarts.append(pos0 + len(body))
        for s_ in range(nsplits):
            part = blk[s_ * ne:(s_ + 1) * ne].tobytes()
            comp = pa.compress(part, codec="lz4_raw", asbytes=True)
            if len(comp) >= ne:
                comp = part
            body += struct.pack("<i", len(comp)) + comp
    head = struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, blocksize, pos0 + len(body))
    return head + struct.pack(f"<{nblocks}i", *bstarts) + body


@pytest.mark.parametrize("typesize,shuffle,blocksize,dont_split", [
    (1, 0, 4096, True), (4, 2, 8192, False), (4, 2, 5000, False), (4, 1, 8192, False), (8, 1, 4096, True),
    (2, 2, 2048, False), (68, 1, 6800, False), (4, 0, 1 << 20, False), (3, 2, 3000, False)])
def test_blosc_decoder_round_trips(typesize, shuffle, blocksize, dont_split):
    """Independent of the reference's fixture: data with every kind of LZ4 match (runs, short periods, long
    repeats, incompressible stretches), several blocks with a ragged last one, split and unsplit streams."""
    from oracle.blosc_shim import blosc_decompress
    from scarf_b200.zarr_store import blosc_decode

    rng = np.random.default_rng(typesize * 100 + shuffle)
    parts = [np.zeros(3000, np.uint8), rng.integers(0, 256, 2500, dtype=np.uint8)]
    for period in (1, 2, 3, 5, 7, 8, 9, 16, 33, 200):
        parts.append(np.tile(rng.integers(0, 256, period, dtype=np.uint8), 4000 // period + 1))
    counts = (rng.random(6000) < 0.07) * rng.integers(1, 40, 6000)
    parts.append(counts.astype("<u4").view(np.uint8))
    raw = np.concatenate(parts).tobytes()
    raw = raw[: len(raw) - len(raw) % typesize + (typesize // 2)]  # a few bytes past the last whole element
    frame = _blosc_encode(raw, typesize, shuffle, blocksize, dont_split)
    assert blosc_decompress(frame) == raw  # the encoder speaks the format the oracle's reader pins on the fixture
    assert blosc_decode(frame).tobytes() == raw


def test_blosc_decoder_survives_corrupted_frames():
    """Byte flips in the streams, in the header / block table, truncations: the decoder returns a status (or decodes
    different bytes), it never reads or writes outside its buffers -- every copy is bounds-checked against both ends."""
    from scarf_b200 import lib

    fn = lib.raw("scf_host_blosc_decode")
    frames = [open(f, "rb").read() for f in _chunk_files()]
    rng = np.random.default_rng(0)
    rejected = 0
    for it in range(1500):
        fr = bytearray(frames[it % len(frames)])
        nbytes = int.from_bytes(fr[4:8], "little")
        mode = it % 4
        if mode == 0:
            for _ in range(rng.integers(1, 6)):
                fr[rng.integers(0, len(fr))] = rng.integers(0, 256)
        elif mode == 1:
            for _ in range(rng.integers(1, 4)):
                fr[rng.integers(0, min(len(fr), 200))] = rng.integers(0, 256)
        elif mode == 2:
            fr = fr[: rng.integers(0, len(fr))]
        else:
            a = rng.integers(16, len(fr) - 8)
            fr[a:a + 8] = rng.integers(0, 256, 8, dtype=np.uint8).tobytes()
        src = np.frombuffer(bytes(fr), dtype=np.uint8)
        out = np.empty(nbytes, dtype=np.uint8)
        rejected += fn(src.ctypes.data if src.size else None, src.size, out.ctypes.data, out.size) != 0
    assert rejected > 1000


def test_blosc_decoder_threads():
    """the decoder is a pure function: chunks are decoded from a thread pool when a store is read"""
    from concurrent.futures import ThreadPoolExecutor

    from scarf_b200.zarr_store import blosc_decode

    frames = [open(f, "rb").read() for f in _chunk_files()]
    want = [blosc_decode(f).tobytes() for f in frames]
    with ThreadPoolExecutor(8) as pool:
        got = list(pool.map(lambda f: blosc_decode(f).tobytes(), frames * 4))
    assert got == want * 4


def test_reference_store_reads_equal_golden(pbmc, tmp_path):
    from oracle.blosc_shim import read_zarr_array
    from scarf_b200.zarr_store import open_group

    z = open_group(STORE, "r")
    counts = z["RNA/counts"]
    assert counts.shape == (892, N_GENES) and counts.chunks == (1000, 1000) and counts.dtype == np.uint32
    want = pbmc["counts"][:, :N_GENES].toarray()
    assert np.array_equal(counts[:], want) and np.array_equal(counts[100:250], want[100:250])
    assert np.array_equal(z["RNA/featureData/names"][:], pbmc["names"][:N_GENES])
    assert z["cellData/I"][:].all() and z["cellData/ids"][:][0] == "AATCACGAGCAGCCCT-1"
    # a write into a Blosc array keeps it readable by any Blosc reader (stored frame), other chunks stay as they were
    dst = str(tmp_path / "copy.zarr")
    shutil.copytree(STORE, dst)
    zc = open_group(dst, "r+")
    keep = np.arange(892) % 3 == 0
    zc["cellData/I"][:] = keep
    part = np.arange(50 * N_GENES, dtype=np.uint32).reshape(50, N_GENES)
    zc["RNA/counts"][10:60] = part
    assert np.array_equal(open_group(dst, "r")["cellData/I"][:], keep)
    want[10:60] = part
    assert np.array_equal(read_zarr_array(os.path.join(dst, "RNA", "counts")), want)  # the oracle's reader agrees
    assert np.array_equal(read_zarr_array(os.path.join(dst, "cellData", "I")), keep)


def test_unsupported_codecs_raise(tmp_path):
    import json

    from scarf_b200.zarr_store import open_group

    dst = str(tmp_path / "copy.zarr")
    shutil.copytree(STORE, dst)
    fn = os.path.join(dst, "cellData", "I", ".zarray")
    meta = json.load(open(fn))
    meta["compressor"] = {"id": "zlib", "level": 1}
    json.dump(meta, open(fn, "w"))
    with pytest.raises(NotImplementedError, match="zlib"):
        open_group(dst, "r")["cellData/I"]


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,density", [(1, 1, 1.0), (7, 5, 0.5), (300, 1001, 0.07), (1000, 4096, 0.1),
                                               (64, 130, 0.0), (513, 36601, 0.06), (40, 515, 1.0)])
def test_dense_to_csr_cases(rows, cols, density):
    import scipy.sparse as sp
    import torch

    from scarf_b200 import ops

    rng = np.random.default_rng(rows * 131 + cols)
    dense = (rng.random((rows, cols)) < density) * rng.integers(1, 2 ** 32, (rows, cols), dtype=np.uint64)
    dense = dense.astype(np.uint32)
    if rows > 3:
        dense[2] = 0  # an empty row
    ld = ops.round_up(cols, 4)
    stage = torch.full((rows, ld), -1, dtype=torch.int32, device="cuda")  # pad columns hold garbage: must be ignored
    stage[:, :cols] = torch.from_numpy(dense.view(np.int32)).cuda()
    cnt, idx, val = ops.dense_to_csr(stage, cols)
    m = sp.csr_matrix(dense)
    m.sort_indices()
    assert np.array_equal(cnt.cpu().numpy(), np.diff(m.indptr))
    assert np.array_equal(idx.cpu().numpy(), m.indices)
    assert np.array_equal(val.cpu().numpy().view(np.uint32), m.data)
    with pytest.raises(ValueError):
        ops.dense_to_csr(torch.zeros((4, 6), dtype=torch.int32, device="cuda"), 6)  # ld % 4 != 0


@pytest.mark.gpu
def test_open_reference_written_store(pbmc, tmp_path):
    """DataStore(path) on the reference's own fixture store: CSR identical to the golden counts, first-open columns
    as `_ini_cell_props` / `_ini_feature_props` define them, and the graph path gives the same neighbours as the
    same counts ingested through from_csr."""
    import torch

    from scarf_b200.datastore import DataStore
    from scarf_b200.zarr_store import open_group

    dst = str(tmp_path / "ref.zarr")
    shutil.copytree(STORE, dst)
    ds = DataStore(dst)
    want = pbmc["counts"][:, :N_GENES].tocsr()
    want.sort_indices()
    csr = ds.RNA.csr
    assert (csr.n_rows, csr.n_cols) == want.shape
    assert np.array_equal(csr.indptr.cpu().numpy(), want.indptr)
    assert np.array_equal(csr.indices.cpu().numpy(), want.indices)
    assert np.array_equal(csr.data.cpu().numpy().view(np.uint32), want.data)
    from scarf_b200 import ops

    part = ops.csr_from_dense_zarr(open_group(dst, "r")["RNA/counts"], "cuda", row_range=(0, 500), n_threads=2)
    assert part.n_rows == 500 and np.array_equal(part.indptr.cpu().numpy(), want.indptr[:501])
    assert np.array_equal(part.indices.cpu().numpy(), want.indices[: want.indptr[500]])
    with pytest.raises(ValueError, match="chunk row boundary"):
        ops.csr_from_dense_zarr(open_group(dst, "r")["RNA/counts"], "cuda", row_range=(100, 500))
    dense = want.toarray().astype(np.int64)
    n_feats = (dense > 0).sum(1)
    assert np.array_equal(ds.cells.fetch_all("RNA_nCounts"), dense.sum(1).astype(np.float64))
    assert np.array_equal(ds.cells.fetch_all("RNA_nFeatures"), n_feats.astype(np.float64))
    assert np.array_equal(ds.cells.fetch_all("I"), n_feats > 10)
    n_cells = (dense > 0).sum(0)
    assert np.array_equal(ds.RNA.feats.fetch_all("nCells"), n_cells.astype(np.float64))
    assert np.array_equal(ds.RNA.feats.fetch_all("dropOuts"), np.abs(892 - n_cells).astype(np.float64))
    assert np.array_equal(ds.RNA.feats.fetch_all("I"), n_cells > 20)
    # reopening finds the columns and changes nothing; real zarr would read the rewritten `I` (stored Blosc frame)
    before = {c: ds.cells.fetch_all(c) for c in ds.cells.columns}
    ds2 = DataStore(dst)
    assert all(np.array_equal(ds2.cells.fetch_all(c), v) for c, v in before.items())
    assert open_group(dst, "r")["cellData/I"].blosc
    # the graph path on it == the same counts through from_csr
    ref = DataStore.from_csr(str(tmp_path / "csr.zarr"), want, feature_ids=pbmc["names"][:N_GENES])
    outs = []
    for d in (ds, ref):
        d.mark_hvgs(top_n=200, min_cells=10)
        d.make_graph(feat_key="hvgs", dims=10, k=11)
        g = d._get_latest_graph_loc("RNA", "I", "hvgs")
        knn = os.path.dirname(g)
        outs.append((d.RNA.feats.fetch_all("I__hvgs"), d.zw[knn]["indices"][:], d.zw[knn]["distances"][:],
                     d.zw[g]["weights"][:]))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    assert outs[0][0].sum() == 200
    torch.cuda.synchronize()
