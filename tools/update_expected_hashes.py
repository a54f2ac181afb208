"""Rewrites profiles/expected_hashes.json from the `parity` blocks of a 1-GPU bench.py line (headline + legs).
usage: python tools/update_expected_hashes.py gpurun_out/bench_full.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    txt = open(sys.argv[1]).read()
    line = json.loads(txt[txt.index('{"metric'):].splitlines()[0])
    assert line["n_gpus"] == 1, "record the expectation from a 1-GPU run"
    path = os.path.join(ROOT, "profiles", "expected_hashes.json")
    cur = json.load(open(path)) if os.path.exists(path) else {}
    blocks = [line.get("parity")] + [leg.get("parity") for leg in (line.get("legs") or {}).values() if isinstance(leg, dict)]
    for p in blocks:
        if not p:
            continue
        assert p["fp64_spot_equal"], "the run's own FP64 spot check failed"
        changed = cur.get(p["hash_key"]) != p["hash"]
        cur[p["hash_key"]] = p["hash"]
        print(p["hash_key"], p["hash"], "(changed)" if changed else "(same)")
    json.dump(cur, open(path, "w"), indent=1)
    open(path, "a").write("\n")


if __name__ == "__main__":
    main()
