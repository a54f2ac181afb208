"""Generates tests/golden/pbmc_ref_store/: a column-trimmed copy of the Zarr store the reference ships as its test
fixture, chunk files byte for byte as the reference wrote them (numcodecs Blosc lz4, bit / byte shuffle).

Run in the build container (needs /root/reference; the GPU box never runs this):
    python tests/golden/make_pbmc_ref_store.py

Input (reference, read-only): scarf/tests/datasets/1K_pbmc_citeseq.zarr.tar.gz -- 892 cells x 36,601 genes, `RNA/counts`
dense uint32 in (1000, 1000) chunks, `cellData` / `RNA/featureData` columns {I, ids, names} in (100000,) chunks; the
store has never been opened by a DataStore (no nCounts / nFeatures / nCells columns yet).
Output: the same tree with `RNA/counts` cut to its first N_GENES columns: only the `shape` entries of the `.zarray`
files change (a Zarr chunk always holds a full chunk shape, so the chunk files stay valid verbatim).
"""
import json
import os
import shutil
import tarfile
import tempfile

REF = "/root/reference/scarf/tests/datasets/1K_pbmc_citeseq.zarr.tar.gz"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pbmc_ref_store")
N_GENES = 3000  # three column chunks

KEEP = ([".zgroup", "cellData/.zgroup", "RNA/.zgroup", "RNA/.zattrs", "RNA/featureData/.zgroup"]
        + [f"cellData/{c}/{f}" for c in ("I", "ids", "names") for f in (".zarray", "0")]
        + [f"RNA/featureData/{c}/{f}" for c in ("I", "ids", "names") for f in (".zarray", "0")]
        + ["RNA/counts/.zarray"] + [f"RNA/counts/0.{j}" for j in range(N_GENES // 1000)])

if os.path.exists(OUT):
    shutil.rmtree(OUT)
with tempfile.TemporaryDirectory() as tmp:
    tarfile.open(REF, "r:gz").extractall(tmp)
    for rel in KEEP:
        src, dst = os.path.join(tmp, rel), os.path.join(OUT, rel)
        if not os.path.exists(src):  # cellData has no .zattrs etc.
            raise FileNotFoundError(rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
for rel, shape in [("RNA/counts", None)] + [(f"RNA/featureData/{c}", [N_GENES]) for c in ("I", "ids", "names")]:
    fn = os.path.join(OUT, rel, ".zarray")
    with open(fn) as f:
        meta = json.load(f)
    meta["shape"] = [meta["shape"][0], N_GENES] if shape is None else shape
    with open(fn, "w") as f:
        json.dump(meta, f, indent=4, sort_keys=True)
print("ok", OUT, sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(OUT) for f in fs), "bytes")
