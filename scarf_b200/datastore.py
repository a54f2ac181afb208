"""Scarf-compatible front end of the path: ``DataStore.mark_hvgs`` / ``make_graph`` / ``run_mapping`` /
``load_graph`` with the reference's signatures, defaults, parameter resolution, error behaviour and Zarr layout
(SURVEY.md 8(b), App. B), computing on the GPU through the C-ABI.

What mirrors what (paths relative to the reference checkout):
  DataStore.__init__      scarf/datastore/base_datastore.py:77,324-401 + scarf/assay.py:201-225 (nCounts, nFeatures,
                          nCells, the cell / feature ``I`` columns)
  mark_hvgs               scarf/datastore/datastore.py:223-314 -> scarf/assay.py:945-1074
  make_graph              scarf/datastore/graph_datastore.py:513-1020 (+ _set_graph_params :63-363)
  run_mapping             scarf/datastore/mapping_datastore.py:31-209 (+ mapping_utils.align_features :148-214)
  load_graph              scarf/datastore/graph_datastore.py:474-511,1022-1075

Differences that are deliberate (and documented in DESIGN.md): raw counts enter as CSR (``from_csr``) and are then
kept as three 1-D arrays under ``<assay>/counts_csr``; a store the reference wrote (dense chunked ``<assay>/counts``,
Blosc) is opened as it is and converted to CSR on the GPU (``ops.csr_from_dense_zarr``); the
``ann__...`` group holds the float32 embedding (``embedding``) in place of hnswlib's ``ann_idx`` file; arrays are
written uncompressed; options the GPU path does not implement raise ``NotImplementedError`` (no CPU fallback).
"""
from __future__ import annotations

import os
import time
from typing import Optional

import numpy as np
import torch

from . import graph
from . import hvg as hvg_host
from . import ops
from .dist import Comm, ShardPlan
from .ops import CsrDevice
from .zarr_store import Group, open_group

__all__ = ["DataStore", "AnnStream", "MetaData", "RNAassay"]


class MetaData:
    """Column store for cell / feature attributes (scarf/metadata.py): boolean key columns select subsets."""

    def __init__(self, zgrp: Group):
        self.z = zgrp
        self.N = int(self.z["I"].shape[0])

    @property
    def columns(self):
        return self.z.keys()

    def get_dtype(self, column):
        return self.z[column].dtype.type if self.z[column].dtype.kind != "b" else bool

    def fetch_all(self, column):
        if column not in self.z:
            raise KeyError(f"ERROR: {column} not found in MetaData")
        return self.z[column][:]

    def active_index(self, key):
        col = self.fetch_all(key)
        if col.dtype != bool:
            raise ValueError(f"ERROR: {key} is not of boolean type. Cannot perform fetch operation")
        return np.where(col)[0]

    def fetch(self, column, key="I"):
        return self.fetch_all(column)[self.active_index(key)]

    def insert(self, column_name, values, fill_value=np.nan, key="I", overwrite=False):
        """metadata.py:395-433: ``values`` covers the rows where ``key`` is True; the rest get ``fill_value``."""
        if column_name in ("I", "ids", "names"):
            raise ValueError(f"ERROR: {column_name} is a protected column name in MetaData class.")
        if column_name in self.z and not overwrite:
            raise ValueError(f"ERROR: {column_name} already exists. Please set `overwrite` to True to overwrite.")
        values = np.asarray(values)
        idx = np.arange(self.N) if len(values) == self.N else self.active_index(key)  # metadata.py:321-323
        if len(values) != len(idx):
            raise ValueError(f"ERROR: `values` are of incorrect length: {len(values)} for {len(idx)} active rows")
        # metadata.py:308-318 (_fill_to_index): bool columns are always filled with False, integer columns with 0 and
        # keep their dtype, everything else with `fill_value`
        if values.dtype.kind == "b":
            fill_value = False
        elif values.dtype.kind in "iu":
            try:
                if np.isnan(fill_value):
                    if len(values) and values.min() <= -1:
                        raise ValueError("`fill_value should be an integer value. ")
                    fill_value = 0
            except TypeError:
                raise ValueError("`fill_value should be an interger value. ")
        full = np.full(self.N, fill_value, dtype=values.dtype)
        full[idx] = values
        a = self.z.create_dataset(column_name, full.shape, full.dtype, (100000,))
        a[:] = full


    def sift(self, column, min_v=-np.inf, max_v=np.inf, keep_bounds=False):
        """metadata.py:483-505."""
        v = self.fetch_all(column)
        return (v >= min_v) & (v <= max_v) if keep_bounds else (v > min_v) & (v < max_v)

    def update_key(self, values, key):
        """metadata.py:437-450: the key column becomes ``key & values`` (values cover all rows)."""
        values = np.asarray(values, dtype=bool)
        if len(values) != self.N:
            raise ValueError(f"ERROR: `values` must cover all {self.N} rows")
        a = self.z[key]
        new = values & self.fetch_all(key)
        a[:] = new


def create_subset_hash(cell_idx, feat_idx) -> int:
    """Assay._create_subset_hash (scarf/assay.py:317-329): Python's hash of the two index tuples -- deterministic for
    integers, so a store written by the reference and one written here recognise each other's cached groups."""
    return hash(tuple([hash(tuple(np.asarray(cell_idx).tolist())), hash(tuple(np.asarray(feat_idx).tolist()))]))


def _slice_csr_rows(c: CsrDevice, a: int, b: int) -> CsrDevice:
    ip = c.indptr[a:b + 1]
    lo, hi = int(ip[0].item()), int(ip[-1].item())
    return CsrDevice((ip - lo).contiguous(), c.indices[lo:hi].contiguous(), c.data[lo:hi].contiguous(), b - a, c.n_cols)


class RNAassay:
    """The raw counts of one assay on the device plus its feature table (scarf/assay.py: RNAassay).  ``csr`` holds all
    rows (loaded on first use); ``csr_rows(lo, hi)`` loads a contiguous span of rows only -- what one rank of a sharded
    run needs (the last span is cached)."""

    def __init__(self, zroot: Group, name: str, cells: MetaData, device, lazy: bool = False):
        self.name, self.z, self.cells, self.device = name, zroot[name], cells, device
        self.feats = MetaData(self.z["featureData"])
        self.sf = 1000  # scarf/assay.py:776
        self._span, self._span_csr = None, None
        if "counts_csr" in self.z:
            self._indptr = self.z["counts_csr"]["indptr"][:].astype(np.int64)
            shape = tuple(self.z["counts_csr"].attrs["shape"])
            self.chunk_rows = 1000
        else:  # a store the reference wrote: dense chunked `counts` (scarf/writers.py:164-204, assay.py:134)
            self._indptr = None
            shape = tuple(self.z["counts"].shape)
            self.chunk_rows = int(self.z["counts"].chunks[0])
        self.n_rows, self.n_cols = int(shape[0]), int(shape[1])
        if self.n_rows != cells.N or self.n_cols != self.feats.N:
            raise ValueError(f"ERROR: counts of assay {name} are {self.n_rows} x {self.n_cols} but the store "
                             f"lists {cells.N} cells and {self.feats.N} features")
        if not lazy:
            self.csr  # noqa: B018  (loads and validates)

    def csr_rows(self, lo: int, hi: int) -> CsrDevice:
        lo, hi = int(lo), int(hi)
        if self._span == (lo, hi):
            return self._span_csr
        if self._span is not None and self._span[0] <= lo and hi <= self._span[1]:
            c = _slice_csr_rows(self._span_csr, lo - self._span[0], hi - self._span[0])
        elif self._indptr is not None:
            g = self.z["counts_csr"]
            ip = self._indptr[lo:hi + 1]
            a, b = (int(ip[0]), int(ip[-1])) if hi > lo else (0, 0)
            c = CsrDevice.from_host(ip - a if hi > lo else np.zeros(1, np.int64), g["indices"][a:b], g["data"][a:b],
                                    (hi - lo, self.n_cols), self.device, validate=True)
        else:
            al = lo - lo % self.chunk_rows  # the dense reader starts on a chunk-row boundary
            c = ops.csr_from_dense_zarr(self.z["counts"], self.device, row_range=(al, hi))
            if al != lo:
                c = _slice_csr_rows(c, lo - al, hi - al)
        if self._span is None or (hi - lo) >= (self._span[1] - self._span[0]):
            self._span, self._span_csr = (lo, hi), c
        return c

    @property
    def csr(self) -> CsrDevice:
        return self.csr_rows(0, self.n_rows)

    @property
    def nCounts(self):
        return self.cells.fetch_all(f"{self.name}_nCounts")


class _PcaSummary:
    """Stands where sklearn's IncrementalPCA object stood for the attributes callers read."""

    def __init__(self, explained_variance, explained_variance_ratio):
        self.explained_variance_, self.explained_variance_ratio_ = explained_variance, explained_variance_ratio


class _KMeans:
    def __init__(self, centers):
        self.cluster_centers_ = centers


class _ExactIndex:
    """Stands where hnswlib's Index stood (scarf/ann.py:14-28): exact search on the stored embedding."""

    def __init__(self, embedding: torch.Tensor, dims: int):
        self.embedding, self.dims = embedding, dims

    def knn_query(self, a, k: int = 1):
        a = np.atleast_2d(np.asarray(a, dtype=np.float32))
        q = torch.zeros((a.shape[0], self.embedding.stride(0)), dtype=torch.float32, device=self.embedding.device)
        q[:, : a.shape[1]] = torch.from_numpy(a).to(q.device)
        idx, dist = ops.knn_l2(q, self.embedding, self.dims, k, self_offset=-1, method=1)
        return idx.cpu().numpy().astype(np.uint64), dist.cpu().numpy()

    def save_index(self, path):
        np.save(path + ".npy", self.embedding[:, : self.dims].cpu().numpy())


class AnnStream:
    """What ``make_graph(return_ann_object=True)`` hands back (scarf/ann.py:55): same attribute names.  Built either
    from a fresh :class:`graph.GraphResult` or from the arrays a previous run left in the store (cache hit)."""

    def __init__(self, k: int, dims: int, mu: torch.Tensor, sigma: torch.Tensor, loadings: torch.Tensor, n_cells: int,
                 embedding_all: torch.Tensor, kmeans=None, labels=None, eigenvalues=None):
        self.k, self.method, self.dims = k, "pca", dims
        self.mu_d, self.sigma_d, self.loadings_d = mu, sigma, loadings  # float64 device tensors (run_mapping)
        self.mu, self.sigma = mu.cpu().numpy(), sigma.cpu().numpy()
        self.loadings = loadings.cpu().numpy()
        self.nCells, self.nFeats = n_cells, int(mu.numel())
        self.harmonizedData, self.data = None, None
        self.annIdx = _ExactIndex(embedding_all, dims)
        self.kmeans, self.clusterLabels = kmeans, labels
        self._pca = None
        if eigenvalues is not None:
            # what the reference's elbow plot reads (graph_datastore.py:1013-1016): variance explained by every kept
            # component.  Exact-PCA values lambda_i / trace(cov); the z-scaled columns have unit variance, so the trace
            # is nFeats * n / (n - 1) (constant columns, kept at sigma = 1, make this a slight over-estimate).
            ev = eigenvalues.detach().cpu().numpy().astype(np.float64)
            total = float(self.nFeats) * n_cells / max(n_cells - 1, 1)
            self._pca = _PcaSummary(ev, ev / total)

    @classmethod
    def from_result(cls, res: graph.GraphResult, kmeans=None, labels=None):
        return cls(res.k, res.dims, res.mu, res.sigma, res.loadings, res.n_cells, res.embedding_all, kmeans, labels,
                   res.eigenvalues)

    def reducer(self, x):
        """``transform_z(x).dot(loadings)`` (scarf/ann.py:138,191-192); accepts a block or one row."""
        x = np.asarray(x, dtype=np.float64)
        return ((x - self.mu) / self.sigma).dot(self.loadings)

    def transform_ann(self, a, k=None, self_indices=None):
        """scarf/ann.py:194-205: query ``k`` (+1 and drop self when ``self_indices`` is given)."""
        k = self.k if k is None else k
        if self_indices is None:
            return self.annIdx.knn_query(a, k)
        i, d = self.annIdx.knn_query(a, k + 1)
        from .graph import fix_knn_query

        return fix_knn_query(i, d, np.asarray(self_indices))


def write_graph_arrays(zw: Group, knn_loc: str, graph_loc: str, res: graph.GraphResult, comm: Optional[Comm] = None,
                       batch_size: int = 1000) -> int:
    """The row-sharded arrays of the graph (scarf/knn_utils.py:54-59,108-117: ``indices`` u8 / ``distances`` f8 in
    chunks of ``batch_size`` rows, ``edges`` u8 / ``weights`` f8 in chunks of ``batch_size * k``): rank 0 creates the
    datasets, then every rank writes the chunks of its own rows [row_offset, row_offset + n_local) -- shards are aligned
    to ``batch_size`` rows (dist.ShardPlan), so no chunk file has two writers.  Returns the bytes this rank wrote."""
    comm = comm or Comm()
    n_local, k = (int(x) for x in res.indices.shape)
    n = res.n_cells
    specs = (("indices", knn_loc, (n, k), "u8", (batch_size,)), ("distances", knn_loc, (n, k), "f8", (batch_size,)),
             ("edges", graph_loc, (n * k, 2), "u8", (batch_size * k,)),
             ("weights", graph_loc, (n * k,), "f8", (batch_size * k,)))
    if comm.rank == 0:
        for name, loc, shape, dt, chunks in specs:
            if loc not in zw:
                zw.create_group(loc)
            zw[loc].create_dataset(name, shape, dt, chunks)
    comm.barrier()
    if comm.world > 1 and res.row_offset % batch_size:
        raise ValueError("sharded graph write: shard boundaries must fall on chunk rows")
    nbytes = 0
    r0 = res.row_offset
    for name, loc, shape, dt, chunks in specs:
        t = getattr(res, name)
        mul = k if name in ("edges", "weights") else 1
        if n_local:
            a = t.cpu().numpy().astype(dt)
            zw[loc][name][r0 * mul:(r0 + n_local) * mul] = a
            nbytes += a.nbytes
    comm.barrier()
    return nbytes


class DataStore:
    """A Scarf-style datastore whose graph path runs on B200.  Create with :meth:`from_csr`, reopen by path."""

    def __init__(self, zarr_loc: str, assay_types: Optional[dict] = None, default_assay: Optional[str] = None,
                 min_features_per_cell: int = 10, min_cells_per_feature: int = 20, mito_pattern: Optional[str] = None,
                 ribo_pattern: Optional[str] = None, nthreads: int = 2, zarr_mode: str = "r+",
                 workspace: Optional[str] = None, synchronizer=None, *, device="cuda", comm: Optional[Comm] = None):
        """scarf/datastore/datastore.py:46-90 + base_datastore.py:77-186 (same parameters, order and defaults; ``device``
        and ``comm`` are keyword-only additions).  ``mito_pattern`` / ``ribo_pattern`` (percentMito / percentRibo
        annotations), ``nthreads`` and ``synchronizer`` (dask / zarr plumbing) have nothing to act on here and are
        accepted for compatibility; ``workspace`` and non-RNA ``assay_types`` raise."""
        if not torch.cuda.is_available():
            raise RuntimeError("scarf_b200.DataStore needs a CUDA device: the path has no CPU fallback")
        if zarr_mode not in ["r", "r+"]:
            raise ValueError("ERROR: Zarr file can only be accessed using either 'r' or 'r+' mode")
        if workspace is not None:
            raise NotImplementedError("scarf_b200.DataStore: `workspace` sub-hierarchies are not implemented")
        self.zw = open_group(zarr_loc, zarr_mode)
        self.z = self.zw
        self._zarr_mode = zarr_mode
        self.nthreads = nthreads
        self.device = torch.device(device)
        self.comm = comm
        self._world = comm.world if comm is not None else 1
        self._rank = comm.rank if comm is not None else 0
        self.last_make_graph_timing = {}
        if "cellData" not in self.zw:
            raise KeyError(f"cellData not found in zarr file at {self.zw.path}")
        self.cells = MetaData(self.zw["cellData"])
        self._defaultAssay = self._load_default_assay(default_assay)
        if assay_types is not None and str(assay_types.get(self._defaultAssay, "RNA")).upper() != "RNA":
            raise NotImplementedError("scarf_b200.DataStore implements the RNAassay path only")
        self._assays = {}
        setattr(self, self._defaultAssay, self._get_assay(self._defaultAssay))
        self._ini_props(self._defaultAssay, min_features_per_cell, min_cells_per_feature)

    @property
    def assay_names(self):
        """Groups that carry the writers' ``is_assay`` attribute (base_datastore.py:127-141)."""
        return [k for k in self.zw.keys() if isinstance(self.zw[k], Group) and "is_assay" in self.zw[k].attrs]

    def _load_default_assay(self, assay_name: Optional[str] = None) -> str:
        """base_datastore.py:143-182: explicit name (must exist) -> the store's ``defaultAssay`` attribute -> the only
        assay; the choice is remembered in the root attributes."""
        names = self.assay_names
        if assay_name is None:
            if "defaultAssay" in self.zw.attrs:
                return self.zw.attrs["defaultAssay"]
            if len(names) != 1:
                raise ValueError("ERROR: You have more than one assay data. "
                                 f"Choose one from: {' '.join(names)}\n using 'default_assay' parameter. "
                                 "Please note that names are case-sensitive.")
            assay_name = names[0]
        elif assay_name not in names:
            raise ValueError(f"ERROR: The provided default assay name: {assay_name} was not found. "
                             f"Please Choose one from: {' '.join(names)}\n"
                             "Please note that the names are case-sensitive.")
        if self._zarr_mode != "r" and self._rank == 0:  # a read-only store keeps its attributes
            self.zw.attrs["defaultAssay"] = assay_name
        return assay_name

    def _barrier(self):
        if self.comm is not None:
            self.comm.barrier()

    def _shard(self, cell_idx: np.ndarray, align: int):
        """This rank's part of the selected cells: rows [s, e) of ``cell_idx`` (contiguous, aligned to ``align`` selected
        rows: the Zarr chunk rows, so that every chunk file has one writer) and the span [lo, hi) of raw rows that
        holds them.  One rank: everything."""
        if self._world == 1:
            return 0, int(cell_idx.size), 0, self.cells.N
        s0, e0 = ShardPlan.make(int(cell_idx.size), self._world, align).rows(self._rank)
        lo, hi = (int(cell_idx[s0]), int(cell_idx[e0 - 1]) + 1) if e0 > s0 else (0, 0)
        return s0, e0, lo, hi

    def _ini_props(self, from_assay, min_features, min_cells):
        """First open of a store (base_datastore.py:324-401 `_ini_cell_props`, assay.py:201-225 `_ini_feature_props`):
        ``<assay>_nCounts`` / ``<assay>_nFeatures`` per cell, ``nCells`` / ``dropOuts`` per feature and the ``I``
        filters, computed on the GPU from the CSR when the columns are missing (every rank scans its own block of raw
        rows; rank 0 writes).  The reference's percentMito / percentRibo columns are cell annotations outside the graph
        path and are not written."""
        assay = self._get_assay(from_assay)
        have = self.cells.columns
        need_cells = f"{from_assay}_nCounts" not in have or f"{from_assay}_nFeatures" not in have
        need_feats = "nCells" not in assay.feats.columns or "dropOuts" not in assay.feats.columns
        self._barrier()  # every rank has looked before rank 0 starts writing
        if need_cells or need_feats:
            a, b = ShardPlan.make(self.cells.N, self._world, assay.chunk_rows).rows(self._rank)
            csr = assay.csr_rows(a, b)
            if need_cells:
                n_counts, n_feats = graph.cell_totals(csr)
                if self.comm is not None and self._world > 1:
                    n_counts, n_feats = self.comm.allgather_rows(n_counts), self.comm.allgather_rows(n_feats)
                if self._rank == 0:
                    self.cells.insert(f"{from_assay}_nCounts", n_counts.cpu().numpy(), overwrite=True)
                    self.cells.insert(f"{from_assay}_nFeatures", n_feats.cpu().numpy().astype(np.float64),
                                      overwrite=True)
            if need_feats:
                nc = graph.gene_ncells(csr, self.comm).cpu().numpy().astype(np.float64)
                if self._rank == 0:
                    assay.feats.insert("nCells", nc, overwrite=True)
                    assay.feats.insert("dropOuts", np.abs(self.cells.N - nc), overwrite=True)
                    assay.feats.update_key(nc > min_cells, "I")  # assay.py:225
            self._barrier()
        v = self.cells.fetch(f"{from_assay}_nFeatures", key="I")
        if len(v) and min_features <= np.median(v):  # base_datastore.py:384-399 (every open; a no-op once applied)
            keep = self.cells.sift(f"{from_assay}_nFeatures", min_features, np.inf)
            cur = self.cells.fetch_all("I")
            changed = not np.array_equal(keep & cur, cur)
            self._barrier()
            if changed and self._rank == 0 and self._zarr_mode != "r":
                self.cells.update_key(keep, "I")
            self._barrier()

    # ---------------------------------------------------------------------------------------------------------
    @classmethod
    def from_csr(cls, zarr_loc: str, counts, feature_ids, feature_names=None, cell_ids=None, assay_name="RNA",
                 min_features_per_cell: int = 10, min_cells_per_feature: int = 20, device="cuda", **kw):
        """Ingest raw counts given as CSR (scipy) -- the role of ``SparseToZarr`` + the first ``DataStore(...)`` open
        (scarf/writers.py:645-772, base_datastore.py:324-401, assay.py:201-225): writes cellData / featureData with
        ids, names, ``I``, ``<assay>_nCounts``, ``<assay>_nFeatures``, ``nCells`` computed on the GPU."""
        counts = counts.tocsr()
        counts.sort_indices()
        n, g = counts.shape
        root = open_group(zarr_loc, "w")
        dev = torch.device(device)
        csr = CsrDevice.from_scipy(counts, dev)
        n_counts, n_feats = graph.cell_totals(csr)
        n_cells = graph.gene_ncells(csr)
        cg = root.create_group("cellData")
        cell_ids = np.asarray(cell_ids if cell_ids is not None else [f"c{i}" for i in range(n)]).astype("U")
        feature_ids = np.asarray(feature_ids).astype("U")
        feature_names = np.asarray(feature_names if feature_names is not None else feature_ids).astype("U")
        nf = n_feats.cpu().numpy().astype(np.float64)

        def put(grp, name, arr):
            a = grp.create_dataset(name, arr.shape, arr.dtype, (100000,))
            a[:] = arr

        put(cg, "ids", cell_ids), put(cg, "names", cell_ids)
        put(cg, "I", nf > min_features_per_cell)  # base_datastore.py:389-399
        put(cg, f"{assay_name}_nCounts", n_counts.cpu().numpy())
        put(cg, f"{assay_name}_nFeatures", nf)
        ag = root.create_group(assay_name)
        ag.attrs["is_assay"] = True
        ag.attrs["misc"] = {"sf": 1000}
        fg = ag.create_group("featureData")
        nc = n_cells.cpu().numpy().astype(np.float64)
        put(fg, "ids", feature_ids), put(fg, "names", feature_names)
        put(fg, "I", nc > min_cells_per_feature)  # assay.py:225
        put(fg, "nCells", nc)
        rg = ag.create_group("counts_csr")
        rg.attrs["shape"] = [int(n), int(g)]
        put(rg, "indptr", counts.indptr.astype(np.int64))
        put(rg, "indices", counts.indices.astype(np.int32))
        put(rg, "data", counts.data.astype(np.uint32))
        return cls(zarr_loc, default_assay=assay_name, device=device, min_features_per_cell=min_features_per_cell,
                   min_cells_per_feature=min_cells_per_feature, **kw)

    # ---------------------------------------------------------------------------------------------------------
    def _get_assay(self, from_assay):
        if from_assay not in self._assays:
            if from_assay not in self.zw or not ("counts_csr" in self.zw[from_assay] or "counts" in self.zw[from_assay]):
                raise ValueError(f"ERROR: Assay {from_assay} was not found.")
            self._assays[from_assay] = RNAassay(self.zw, from_assay, self.cells, self.device, lazy=self._world > 1)
        return self._assays[from_assay]

    def _get_latest_keys(self, from_assay, cell_key, feat_key):
        """graph_datastore.py / base: fall back on the keys the latest make_graph stored on the assay group."""
        if from_assay is None:
            from_assay = self._defaultAssay
        a = self.zw[from_assay].attrs
        if cell_key is None:
            cell_key = a.get("latest_cell_key", "I")
        if feat_key is None:
            feat_key = a.get("latest_feat_key")
            if feat_key is None:
                raise ValueError("ERROR: No graph has been made yet: `feat_key` is unknown")
        return from_assay, cell_key, feat_key

    # ---------------------------------------------------------------------------------------------------------
    def mark_hvgs(self, from_assay: Optional[str] = None, cell_key: Optional[str] = None,
                  min_cells: Optional[int] = None, top_n: int = 500, min_var: float = -np.inf, max_var: float = np.inf,
                  min_mean: float = -np.inf, max_mean: float = np.inf, n_bins: int = 200, lowess_frac: float = 0.1,
                  blacklist: str = hvg_host.DEFAULT_BLACKLIST, keep_bounds: bool = False, show_plot: bool = True,
                  hvg_key_name: str = "hvgs", max_cells: Optional[float] = np.inf, **plot_kwargs) -> None:
        """scarf/datastore/datastore.py:223-314 (same parameters and defaults).  Stores ``<cell_key>__<hvg_key_name>``
        (bool) and the statistics columns of ``set_summary_stats`` in the feature table.  ``show_plot`` / plot
        arguments are accepted and ignored: plotting is not part of this path."""
        if cell_key is None:
            cell_key = "I"
        if max_cells is None:
            max_cells = np.inf
        if cell_key not in self.cells.columns:
            raise ValueError(f"ERROR: cell_key {cell_key} not found in cell metadata")
        if from_assay is None:
            from_assay = self._defaultAssay
        assay = self._get_assay(from_assay)
        if not isinstance(assay, RNAassay):
            raise TypeError(f"ERROR: This method of feature selection can only be applied to RNAassay type of assay.")
        if min_cells is None:
            min_cells = int(0.01 * self.cells.N)  # datastore.py:291
        cell_idx = self.cells.active_index(cell_key)
        s0, e0, lo, hi = self._shard(cell_idx, assay.chunk_rows)
        cells = torch.from_numpy(cell_idx[s0:e0] - lo).to(self.device)
        n_counts = torch.from_numpy(assay.nCounts[lo:hi]).to(self.device)
        feat_I = assay.feats.fetch_all("I")
        mask, st = graph.mark_hvgs_csr(assay.csr_rows(lo, hi), cells, feat_I, n_counts, self.cells.N,
                                       gene_names=assay.feats.fetch_all("names"), top_n=top_n, min_cells=min_cells,
                                       min_mean=min_mean, max_mean=max_mean, n_bins=n_bins, lowess_frac=lowess_frac,
                                       blacklist=blacklist, comm=self.comm, return_stats=True, min_var=min_var,
                                       max_var=max_var, max_cells=max_cells, keep_bounds=keep_bounds)
        self._barrier()
        if self._rank != 0:  # the statistics are identical on every rank after the all-reduce: one writer
            self._barrier()
            return None
        ident = f"{cell_key}__"
        for name, col in (("normed_tot", "normed_tot"), ("avg", "avg"), ("nz_mean", "nz_mean"),
                          ("sigmas", "sigmas"), ("normed_n", "normed_n"), ("c_var", f"c_var__{n_bins}__{lowess_frac}")):
            assay.feats.insert(ident + col, st[name][feat_I], overwrite=True)  # assay.py:884-897, metadata.py:612
        assay.feats.insert(ident + hvg_key_name, mask[feat_I], fill_value=False, overwrite=True)
        self._barrier()
        return None

    # ---------------------------------------------------------------------------------------------------------
    def _set_graph_params(self, from_assay, cell_key, feat_key, log_transform=None, renormalize_subset=None,
                          reduction_method="auto", dims=None, pca_cell_key=None, ann_metric=None, ann_efc=None,
                          ann_ef=None, ann_m=None, rand_state=None, k=None, n_centroids=None, local_connectivity=None,
                          bandwidth=None) -> tuple:
        """graph_datastore.py:63-363: every ``None`` resolves as explicit -> value cached by the latest run in the
        same branch of the Zarr tree -> default."""
        dv = {"log_transform": True, "renormalize_subset": True, "dims": 11, "ann_metric": "l2", "rand_state": 4466,
              "k": 11, "n_centroids": 1000, "local_connectivity": 1.0, "bandwidth": 1.5}
        zw = self.zw
        normed_loc = f"{from_assay}/normed__{cell_key}__{feat_key}"
        cached = zw[normed_loc].attrs.get("subset_params") if normed_loc in zw else None
        if log_transform is None:
            log_transform = cached["log_transform"] if cached else dv["log_transform"]
        if renormalize_subset is None:
            renormalize_subset = cached["renormalize_subset"] if cached else dv["renormalize_subset"]
        log_transform, renormalize_subset = bool(log_transform), bool(renormalize_subset)
        c_dims = c_pca = None
        if normed_loc in zw and "latest_reduction" in zw[normed_loc].attrs:
            c_dims, c_pca = zw[normed_loc].attrs["latest_reduction"].rsplit("__", 2)[1:]
        dims_given = dims is not None
        if dims is None:
            dims = int(c_dims) if c_dims is not None else dv["dims"]
        if pca_cell_key is None:
            pca_cell_key = c_pca if c_pca is not None else cell_key
        elif not dims_given:  # as executed, the reference checks the column only when `dims` is not given (:188-208)
            if pca_cell_key not in self.cells.columns:
                raise ValueError(f"ERROR: `pca_use_cell_key` {pca_cell_key} does not exist in cell metadata")
            if self.cells.get_dtype(pca_cell_key) != bool:
                raise TypeError("ERROR: Type of `pca_use_cell_key` column in cell metadata should be `bool`")
        dims = int(dims)
        reduction_method = reduction_method.lower()
        if reduction_method not in ["pca", "lsi", "auto", "custom"]:
            raise ValueError("ERROR: Please choose either 'pca' or 'lsi' as reduction method")
        if reduction_method == "auto":
            reduction_method = "pca"  # RNAassay (graph_datastore.py:49-61)
        reduction_loc = f"{normed_loc}/reduction__{reduction_method}__{dims}__{pca_cell_key}"
        c = [None] * 5
        if reduction_loc in zw and "latest_ann" in zw[reduction_loc].attrs:
            c = zw[reduction_loc].attrs["latest_ann"].rsplit("/", 1)[1].split("__")[1:]
        if ann_metric is None:
            ann_metric = c[0] if c[0] is not None else dv["ann_metric"]
        if ann_efc is None and c[1] is not None:
            ann_efc = int(c[1])
        if ann_ef is None and c[2] is not None:
            ann_ef = int(c[2])
        if ann_m is None:
            ann_m = int(c[3]) if c[3] is not None else min(max(48, int(dims * 1.5)), 64)
        if rand_state is None:
            rand_state = int(c[4]) if c[4] is not None else dv["rand_state"]
        ann_metric, ann_m, rand_state = str(ann_metric), int(ann_m), int(rand_state)
        if k is None:
            k = dv["k"]
            if reduction_loc in zw and "latest_ann" in zw[reduction_loc].attrs:
                ann_loc = zw[reduction_loc].attrs["latest_ann"]
                if ann_loc in zw and "latest_knn" in zw[ann_loc].attrs:
                    k = int(zw[ann_loc].attrs["latest_knn"].rsplit("__", 1)[1])
        k = int(k)
        ann_ef = int(min(100, max(k * 3, 50)) if ann_ef is None else ann_ef)
        ann_efc = int(min(100, max(k * 3, 50)) if ann_efc is None else ann_efc)
        ann_loc = f"{reduction_loc}/ann__{ann_metric}__{ann_efc}__{ann_ef}__{ann_m}__{rand_state}"
        knn_loc = f"{ann_loc}/knn__{k}"
        if n_centroids is None:
            n_centroids = dv["n_centroids"]
            if reduction_loc in zw and "latest_kmeans" in zw[reduction_loc].attrs:
                n_centroids = int(zw[reduction_loc].attrs["latest_kmeans"].split("/")[-1].split("__")[1])
        n_centroids = int(n_centroids)
        c_lc = c_bw = None
        if knn_loc in zw and "latest_graph" in zw[knn_loc].attrs:
            c_lc, c_bw = map(float, zw[knn_loc].attrs["latest_graph"].rsplit("/", 1)[1].split("__")[1:])
        if local_connectivity is None:
            local_connectivity = c_lc if c_lc is not None else dv["local_connectivity"]
        if bandwidth is None:
            bandwidth = c_bw if c_bw is not None else dv["bandwidth"]
        return (log_transform, renormalize_subset, reduction_method, dims, pca_cell_key, ann_metric, ann_efc, ann_ef,
                ann_m, rand_state, k, n_centroids, float(local_connectivity), float(bandwidth))

    def make_graph(self, from_assay: Optional[str] = None, cell_key: Optional[str] = None,
                   feat_key: Optional[str] = None, pca_cell_key: Optional[str] = None, reduction_method: str = "auto",
                   dims: Optional[int] = None, k: Optional[int] = None, ann_metric: Optional[str] = None,
                   ann_efc: Optional[int] = None, ann_ef: Optional[int] = None, ann_m: Optional[int] = None,
                   ann_parallel: bool = False, rand_state: Optional[int] = None, n_centroids: Optional[int] = None,
                   batch_size: Optional[int] = None, log_transform: Optional[bool] = None,
                   renormalize_subset: Optional[bool] = None, local_connectivity: Optional[float] = None,
                   bandwidth: Optional[float] = None, update_keys: bool = True, return_ann_object: bool = False,
                   custom_loadings: Optional[np.ndarray] = None, feat_scaling: bool = True,
                   lsi_skip_first: bool = True, harmonize: bool = False, batch_columns=None,
                   show_elbow_plot: bool = False, ann_index_fetcher=None, ann_index_saver=None):
        """graph_datastore.py:513-1020 on the GPU.  Same groups, array names, dtypes, chunking and ``latest_*``
        attributes; a re-run with the same parameters finds its groups and only reloads (the Zarr tree is the
        checkpoint)."""
        if from_assay is None:
            from_assay = self._defaultAssay
        assay = self._get_assay(from_assay)
        if batch_size is None:
            batch_size = assay.chunk_rows  # assay.rawData.chunksize[0] (graph_datastore.py:738-745; writers.py:164-204)
        if cell_key is None:
            cell_key = "I"
        if feat_key is None:
            bool_cols = [x.split("__", 1) for x in assay.feats.columns
                         if assay.feats.get_dtype(x) == bool and x != "I"]
            bool_cols = " ".join(f"{x[1]}({x[0]})" for x in bool_cols)
            raise ValueError(
                "ERROR: You have to choose which features that should be used for graph construction. "
                "Ideally you should have performed a feature selection step before making this graph. "
                "Feature selection step adds a column to your feature table. \n"
                f"You have following boolean columns in the feature metadata of assay {from_assay} which you can "
                f"choose from: {bool_cols}\n The values in brackets indicate the cell_key for which the feat_key is "
                "available. Choosing 'I' as `feat_key` means that you will use all the genes for graph creation.")
        if custom_loadings is not None or reduction_method.lower() in ("lsi", "custom") or harmonize or \
                not feat_scaling or ann_index_fetcher is not None or ann_index_saver is not None:
            raise NotImplementedError("scarf_b200.make_graph implements reduction_method='pca' with feature scaling; "
                                      "custom loadings, LSI, Harmony and custom index stores are not on the GPU path")
        (log_transform, renormalize_subset, reduction_method, dims, pca_cell_key, ann_metric, ann_efc, ann_ef, ann_m,
         rand_state, k, n_centroids, local_connectivity, bandwidth) = self._set_graph_params(
            from_assay, cell_key, feat_key, log_transform, renormalize_subset, reduction_method, dims, pca_cell_key,
            ann_metric, ann_efc, ann_ef, ann_m, rand_state, k, n_centroids, local_connectivity, bandwidth)
        if ann_metric != "l2":
            raise NotImplementedError("scarf_b200.make_graph implements ann_metric='l2' only")
        zw = self.zw
        normed_loc = f"{from_assay}/normed__{cell_key}__{feat_key}"
        reduction_loc = f"{normed_loc}/reduction__{reduction_method}__{dims}__{pca_cell_key}"
        ann_loc = f"{reduction_loc}/ann__{ann_metric}__{ann_efc}__{ann_ef}__{ann_m}__{rand_state}"
        knn_loc = f"{ann_loc}/knn__{k}"
        kmeans_loc = f"{reduction_loc}/kmeans__{n_centroids}__{rand_state}"
        graph_loc = f"{knn_loc}/graph__{local_connectivity}__{bandwidth}"

        # ---- Assay.save_normalized_data bookkeeping (assay.py:400-478): subset hash / params, latest keys ----
        cell_idx = self.cells.active_index(cell_key)
        feat_col = cell_key + "__" + feat_key if feat_key != "I" else "I"
        feat_mask = assay.feats.fetch_all(feat_col)
        if feat_mask.dtype != bool:
            raise ValueError(f"ERROR: {feat_col} is not of boolean type. Cannot perform fetch operation")
        feat_idx = np.where(feat_mask)[0]
        subset_hash = create_subset_hash(cell_idx, feat_idx)
        subset_params = {"log_transform": log_transform, "renormalize_subset": renormalize_subset}
        self._barrier()  # every rank has read the store's state before rank 0 changes it
        stale = False
        if normed_loc in zw:
            at = zw[normed_loc].attrs
            # a cached group is stale when it records a different subset or different parameters (assay.py:446-460);
            # a group without a record (foreign or legacy content) is left alone
            stale = ("subset_hash" in at and at["subset_hash"] != subset_hash) or \
                    ("subset_params" in at and at["subset_params"] != subset_params)
        cached = (not stale) and all(x in zw for x in (reduction_loc, ann_loc, knn_loc, graph_loc, kmeans_loc)) and \
            "embedding" in zw[ann_loc] and "mu" in zw[normed_loc] and "reduction" in zw[reduction_loc]
        self._barrier()
        if self._rank == 0:
            if stale:
                zw.create_group(normed_loc, overwrite=True)  # everything below it is recomputed
            if normed_loc not in zw:
                zw.create_group(normed_loc)
            zw[normed_loc].attrs["subset_hash"] = subset_hash
            zw[normed_loc].attrs["subset_params"] = subset_params
            if update_keys:
                zw[from_assay].attrs["latest_cell_key"] = cell_key
                zw[from_assay].attrs["latest_feat_key"] = feat_key
        self._barrier()

        if cached:
            # the Zarr tree is the checkpoint (graph_datastore.py:797-881, 916-1001): nothing is recomputed; the object
            # run_mapping needs is rebuilt from the stored mu / sigma / loadings / embedding
            if self._rank == 0:
                self._set_latest(normed_loc, reduction_loc, ann_loc, kmeans_loc, knn_loc, graph_loc)
            self._barrier()
            self.last_make_graph_timing = {"cache_hit": True}
            return self._load_ann_object(normed_loc, reduction_loc, ann_loc, kmeans_loc, dims, k) \
                if return_ann_object else None

        t_start = time.time()
        s0, e0, lo, hi = self._shard(cell_idx, batch_size)
        cells_t = torch.from_numpy(cell_idx[s0:e0] - lo).to(self.device)
        n_counts = torch.from_numpy(assay.nCounts[lo:hi]).to(self.device)
        pca_rows = None
        if pca_cell_key != cell_key:  # use_for_pca = cells.fetch(pca_cell_key, key=cell_key)  (graph_datastore.py:764)
            use_for_pca = self.cells.fetch_all(pca_cell_key)[cell_idx]
            if use_for_pca.dtype != bool:
                raise ValueError(f"ERROR: `pca_use_cell_key` {pca_cell_key} is not of boolean type")
            if not use_for_pca.all():
                pca_rows = torch.from_numpy(np.where(use_for_pca[s0:e0])[0]).to(self.device)
        res = graph.make_graph_csr(assay.csr_rows(lo, hi), cells_t, feat_mask, dims=dims, k=k, lc=local_connectivity,
                                   bw=bandwidth, batch_size=batch_size, log_transform=log_transform,
                                   renormalize_subset=renormalize_subset, n_counts=n_counts, comm=self.comm,
                                   gram_mode=3, knn_method=1, pca_rows=pca_rows)
        torch.cuda.synchronize()
        t_graph = time.time()
        centers, labels = graph.fit_kmeans(res.embedding_all, res.dims, max(n_centroids, 2), rand_state)
        ann_obj = AnnStream.from_result(res, _KMeans(centers.cpu().numpy().astype(np.float64)),
                                        labels.cpu().numpy().astype(np.float64))
        t_kmeans = time.time()
        self._write_graph(res, ann_obj, normed_loc, reduction_loc, ann_loc, kmeans_loc, knn_loc, graph_loc, batch_size)
        if self._rank == 0:
            self._set_latest(normed_loc, reduction_loc, ann_loc, kmeans_loc, knn_loc, graph_loc)
        self._barrier()
        t_end = time.time()
        self.last_make_graph_timing = {"graph_s": round(t_graph - t_start, 4), "kmeans_s": round(t_kmeans - t_graph, 4),
                                       "zarr_write_s": round(t_end - t_kmeans, 4)}
        return ann_obj if return_ann_object else None

    def _load_ann_object(self, normed_loc, reduction_loc, ann_loc, kmeans_loc, dims, k) -> AnnStream:
        """AnnStream from the arrays of a finished run (mapping_datastore.py:143 reaches make_graph only to get this
        object: loadings, mu / sigma and the index are loaded, not recomputed -- graph_datastore.py:797-881)."""
        zw, dev = self.zw, self.device
        mu = torch.from_numpy(zw[normed_loc]["mu"][:]).to(dev)
        sigma = torch.from_numpy(zw[normed_loc]["sigma"][:]).to(dev)
        load = torch.from_numpy(np.ascontiguousarray(zw[reduction_loc]["reduction"][:])).to(dev)
        emb = zw[ann_loc]["embedding"][:]
        n, d = emb.shape
        y = torch.zeros((n, ops.round_up(d, 32)), dtype=torch.float32, device=dev)
        y[:, :d] = torch.from_numpy(np.ascontiguousarray(emb, dtype=np.float32)).to(dev)
        km = _KMeans(zw[kmeans_loc]["cluster_centers"][:]) if "cluster_centers" in zw[kmeans_loc] else None
        lb = zw[kmeans_loc]["cluster_labels"][:] if "cluster_labels" in zw[kmeans_loc] else None
        k_eff = int(zw[ann_loc.rstrip("/") + f"/knn__{k}"]["indices"].shape[1])
        return AnnStream(k_eff, int(d), mu, sigma, load, int(n), y, km, lb)

    def save_normalized_data(self, from_assay: Optional[str] = None, cell_key: str = "I", feat_key: str = "hvgs",
                             batch_size: int = 1000, log_transform: bool = True, renormalize_subset: bool = True):
        """Assay.save_normalized_data (scarf/assay.py:400-478 -> writers.dask_to_zarr :915-935): materialises the
        normalised feature matrix ``x = log1p(sf * c / scalar)`` as ``<assay>/normed__<ck>__<fk>/data`` (f8, shape
        (N, H), chunks (batch, H)).  ``make_graph`` never needs it here (Z stays in HBM), but the reference's other
        readers of that array (CORAL, ``integrate_assays``, ``metric_silhouette``) do, so the drop-in store can hold it:
        the values come from the same fused normalise kernel, one row batch at a time.  Returns the array's location."""
        if from_assay is None:
            from_assay = self._defaultAssay
        assay = self._get_assay(from_assay)
        zw = self.zw
        cell_idx = self.cells.active_index(cell_key)
        feat_col = cell_key + "__" + feat_key if feat_key != "I" else "I"
        feat_mask = assay.feats.fetch_all(feat_col)
        if feat_mask.dtype != bool:
            raise ValueError(f"ERROR: {feat_col} is not of boolean type. Cannot perform fetch operation")
        feat_idx = np.where(feat_mask)[0]
        n_feat = int(feat_idx.size)
        normed_loc = f"{from_assay}/normed__{cell_key}__{feat_key}"
        subset_hash = create_subset_hash(cell_idx, feat_idx)
        subset_params = {"log_transform": bool(log_transform), "renormalize_subset": bool(renormalize_subset)}
        if normed_loc in zw:
            at = zw[normed_loc].attrs
            if ("subset_hash" in at and at["subset_hash"] != subset_hash) or \
                    ("subset_params" in at and at["subset_params"] != subset_params):
                zw.create_group(normed_loc, overwrite=True)  # assay.py:459-460: a different subset owns the group
        if normed_loc not in zw:
            zw.create_group(normed_loc)
        dev = self.device
        cmap = np.full(assay.csr.n_cols, -1, dtype=np.int32)
        cmap[feat_idx] = np.arange(n_feat, dtype=np.int32)
        col_map = torch.from_numpy(cmap).to(dev)
        cells_t = torch.from_numpy(cell_idx).to(dev)
        if renormalize_subset:
            row_sum, _ = ops.csr_row_sums(assay.csr, cells_t, col_map)  # assay.py:814-823
        else:
            row_sum = torch.from_numpy(assay.nCounts).to(dev)[cells_t].contiguous()
        out = zw[normed_loc].create_dataset("data", (int(cell_idx.size), n_feat), "f8", (batch_size, n_feat))
        ldz = ops.round_up(n_feat, 4)
        buf = torch.empty((batch_size, ldz), dtype=torch.float32, device=dev)
        for lo in range(0, int(cell_idx.size), batch_size):
            hi = min(lo + batch_size, int(cell_idx.size))
            ops.csr_norm_scale(assay.csr, cells_t[lo:hi].contiguous(), col_map, n_feat, row_sum[lo:hi].contiguous(), buf,
                               graph.SF, log_transform)
            out[lo:hi] = buf[: hi - lo, :n_feat].cpu().numpy().astype(np.float64)
        zw[normed_loc].attrs["subset_hash"] = subset_hash      # assay.py:471-472
        zw[normed_loc].attrs["subset_params"] = subset_params
        return f"{normed_loc}/data"

    def _set_latest(self, normed_loc, reduction_loc, ann_loc, kmeans_loc, knn_loc, graph_loc):
        zw = self.zw  # graph_datastore.py:1003-1008
        zw[normed_loc].attrs["latest_reduction"] = reduction_loc
        zw[reduction_loc].attrs["latest_ann"] = ann_loc
        zw[reduction_loc].attrs["latest_kmeans"] = kmeans_loc
        zw[ann_loc].attrs["isHarmonized"] = False
        zw[ann_loc].attrs["latest_knn"] = knn_loc
        zw[knn_loc].attrs["latest_graph"] = graph_loc

    def _write_graph(self, res, ann_obj, normed_loc, reduction_loc, ann_loc, kmeans_loc, knn_loc, graph_loc,
                     batch_size):
        """Array names / dtypes / chunks of SURVEY.md App. B.  The replicated arrays (mu, sigma, loadings, k-means, the
        embedding -- every rank holds all of it after the all-gather) are written by rank 0, the row-sharded ones by the
        rank that owns the rows (:func:`write_graph_arrays`)."""
        zw = self.zw

        def put(loc, name, arr, chunks, dtype):
            if loc not in zw:
                zw.create_group(loc)
            a = zw[loc].create_dataset(name, arr.shape, dtype, chunks)
            a[:] = arr.astype(dtype)

        if self._rank == 0:
            put(normed_loc, "mu", ann_obj.mu, (100000,), "f8")            # graph_datastore.py:778-796
            put(normed_loc, "sigma", ann_obj.sigma, (100000,), "f8")
            put(reduction_loc, "reduction", ann_obj.loadings, (batch_size, ann_obj.loadings.shape[0]), "f8")  # :921-928
            put(ann_loc, "embedding", res.embedding_all[:, : res.dims].cpu().numpy(), (batch_size,), "f4")
            put(kmeans_loc, "cluster_centers", ann_obj.kmeans.cluster_centers_, (1000, 1000), "f8")       # :958-976
            put(kmeans_loc, "cluster_labels", ann_obj.clusterLabels, (100000,), "f8")
        write_graph_arrays(zw, knn_loc, graph_loc, res, self.comm, batch_size)

    # ---------------------------------------------------------------------------------------------------------
    def _get_latest_graph_loc(self, from_assay, cell_key, feat_key):
        """graph_datastore.py:379-397."""
        normed_loc = f"{from_assay}/normed__{cell_key}__{feat_key}"
        reduction_loc = self.zw[normed_loc].attrs["latest_reduction"]
        ann_loc = self.zw[reduction_loc].attrs["latest_ann"]
        knn_loc = self.zw[ann_loc].attrs["latest_knn"]
        return self.zw[knn_loc].attrs["latest_graph"]

    def load_graph(self, from_assay: Optional[str] = None, cell_key: Optional[str] = None,
                   feat_key: Optional[str] = None, symmetric: Optional[bool] = None,
                   upper_only: Optional[bool] = None, use_k: Optional[int] = None, graph_loc: Optional[str] = None):
        """graph_datastore.py:1022-1075 + _store_to_sparse :474-511 -> scipy sparse matrix.  As in the reference the
        defaults (None) return the stored, directed graph; ``symmetric=True`` gives ``g + g.T - g * g.T`` and
        ``upper_only=True`` its upper triangle (what run_leiden / run_umap ask for)."""
        from_assay, cell_key, feat_key = self._get_latest_keys(from_assay, cell_key, feat_key)
        if graph_loc is None:
            graph_loc = self._get_latest_graph_loc(from_assay, cell_key, feat_key)
        if graph_loc not in self.zw:
            raise ValueError(f"{graph_loc} not found in zarr location {self.zw.path}. Run `make_graph` for assay "
                             f"{from_assay}")
        knn_loc = graph_loc.rsplit("/", 1)[0]
        n_cells, k = self.zw[knn_loc]["indices"].shape
        store = self.zw[graph_loc]
        return graph.graph_to_sparse(store["edges"][:], store["weights"][:], n_cells, k, use_k, symmetric, upper_only,
                                     device=getattr(self, "device", None))

    # ---------------------------------------------------------------------------------------------------------
    def run_mapping(self, target_assay: RNAassay, target_name: str, target_feat_key: str, target_cell_key: str = "I",
                    from_assay: Optional[str] = None, cell_key: str = "I", feat_key: Optional[str] = None,
                    save_k: int = 3, batch_size: int = 1000, ref_mu: bool = True, ref_sigma: bool = True,
                    run_coral: bool = False, exclude_missing: bool = False, filter_null: bool = False,
                    feat_scaling: bool = True, ann_index_fetcher=None, ann_index_saver=None) -> None:
        """mapping_datastore.py:31-209: project the cells of ``target_assay`` onto this store's graph and store
        ``projections/<target_name>/{indices (u8), distances (f8)}``."""
        from_assay, cell_key, feat_key = self._get_latest_keys(from_assay, cell_key, feat_key)
        source = self._get_assay(from_assay)
        if type(target_assay) != type(source):
            raise TypeError(f"ERROR: Source assay ({type(source)}) and target assay ({type(target_assay)}) are of "
                            "different types. Mapping can only be performed between same assay types")
        if target_feat_key == feat_key:
            raise ValueError(f"ERROR: `target_feat_key` cannot be sample as `feat_key`: {feat_key}")
        if run_coral or exclude_missing or filter_null or not feat_scaling:
            raise NotImplementedError("scarf_b200.run_mapping implements the default path (no CORAL, "
                                      "exclude_missing=False, filter_null=False, feat_scaling=True)")
        target_assay.sf = source.sf
        feat_col = cell_key + "__" + feat_key if feat_key != "I" else "I"
        s_feat_idx = np.where(source.feats.fetch_all(feat_col))[0]
        t_col = graph.order_features(source.feats.fetch_all("ids"), target_assay.feats.fetch_all("ids"), s_feat_idx)
        ann_obj = self.make_graph(from_assay=from_assay, cell_key=cell_key, feat_key=feat_key,
                                  return_ann_object=True, update_keys=False)
        if save_k > ann_obj.k:
            save_k = ann_obj.k
        normed_loc = f"{from_assay}/normed__{cell_key}__{feat_key}"
        params = self.zw[normed_loc].attrs["subset_params"]
        t_idx = target_assay.cells.active_index(target_cell_key)
        if self._world == 1:
            q0, q1, lo, hi = 0, int(t_idx.size), 0, target_assay.n_rows
        else:  # every rank maps a block of the target cells (aligned to the projection arrays' chunk rows)
            q0, q1 = ShardPlan.make(int(t_idx.size), self._world, batch_size).rows(self._rank)
            lo, hi = (int(t_idx[q0]), int(t_idx[q1 - 1]) + 1) if q1 > q0 else (0, 0)
        t_cells = torch.from_numpy(t_idx[q0:q1] - lo).to(self.device)
        t_counts = torch.from_numpy(target_assay.nCounts[lo:hi]).to(self.device)
        m = graph.run_mapping_csr(target_assay.csr_rows(lo, hi), t_cells, t_col, ann_obj.mu_d, ann_obj.sigma_d,
                                  ann_obj.loadings_d, ann_obj.annIdx.embedding, ann_obj.dims, save_k=save_k,
                                  use_ref_mu=ref_mu, use_ref_sigma=ref_sigma, log_transform=params["log_transform"],
                                  renormalize_subset=params["renormalize_subset"], n_counts=t_counts, comm=self.comm)
        nc = int(t_idx.size)
        self._barrier()
        if self._rank == 0:
            if "projections" not in self.zw[from_assay]:
                self.zw[from_assay].create_group("projections")
            store = self.zw[from_assay]["projections"].create_group(target_name, overwrite=True)
            store.create_dataset("indices", (nc, save_k), "u8", (batch_size,))
            store.create_dataset("distances", (nc, save_k), "f8", (batch_size,))
        self._barrier()
        store = self.zw[from_assay]["projections"][target_name]
        if q1 > q0:
            store["indices"][q0:q1] = m.indices.cpu().numpy().astype(np.uint64)
            store["distances"][q0:q1] = m.distances.cpu().numpy().astype(np.float64)
        self._barrier()
        return None
