"""
Generates tests/golden/*.  Run from the repo root:  python tests/golden/make_golden.py

The reference (C#) cannot run in this image and ships no golden files, so the
vectors here are:
  * dotnet_random.json  -- published first outputs of .NET `new Random(seed).Next()`
                           for seeds 0 / 1 / 42 (legacy seeded generator), the values
                           the reference's level draw and test fixtures depend on;
  * metric_float64.npz  -- float64 numpy distances computed the way the reference's
                           own bindings/__tests__/metric_test.py:7-31 does, for random
                           pairs at dims 96/100/127/128/960 (tolerance 1e-5 there);
  * reference_windows.json -- the pass windows the reference's tests assert
                           (GraphTests.cs:29, ParametersTests.cs:29,47,65,
                           GraphResizeTests.cs) with the oracle's measured value
                           recorded beside each as a regression pin.
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))


def main():
    json.dump({"next": {"0": [1559595546], "1": [534011718], "42": [1434747710]},
               "source": "well-known outputs of System.Random(seed).Next() on .NET (legacy seeded algorithm)"},
              open(HERE / "dotnet_random.json", "w"), indent=1)

    rng = np.random.default_rng(20261018)
    out = {}
    for dim in (96, 100, 127, 128, 960):
        a = rng.random((64, dim), dtype=np.float32)
        b = rng.random((64, dim), dtype=np.float32)
        a64, b64 = a.astype(np.float64), b.astype(np.float64)
        diff = a64 - b64
        out[f"a_{dim}"], out[f"b_{dim}"] = a, b
        out[f"sq_{dim}"] = np.einsum("ij,ij->i", diff, diff)
        na, nb = np.linalg.norm(a64, axis=1), np.linalg.norm(b64, axis=1)
        out[f"cos_{dim}"] = 1.0 - np.einsum("ij,ij->i", a64, b64) / (na * nb)
        an, bn = a64 / na[:, None], b64 / nb[:, None]
        out[f"an_{dim}"], out[f"bn_{dim}"] = an.astype(np.float32), bn.astype(np.float32)
        out[f"ucos_{dim}"] = 1.0 - np.einsum("ij,ij->i", out[f"an_{dim}"].astype(np.float64), out[f"bn_{dim}"].astype(np.float64))
    np.savez_compressed(HERE / "metric_float64.npz", **out)

    from oracle import oracle as O
    v2000 = O.normalize_rows_like_reference(O.random_vectors(128, 2000))
    v1000 = O.normalize_rows_like_reference(O.random_vectors(128, 1000))
    v5000 = O.normalize_rows_like_reference(O.random_vectors(128, 5000))
    cases = [
        ("BuildGraphSingleThread", "GraphTests.cs:17-37", "ucosine", {}, "v2000", 0.85, 1.01),
        ("TestParameterMinNN", "ParametersTests.cs:15-30", "cosine", {"min_nn": 1}, "v1000", 0.70, 0.90),
        ("TestParameterMaxCandidates", "ParametersTests.cs:33-48", "cosine", {"max_candidates": 32}, "v1000", 0.90, 1.01),
        ("TestParameterLowRecall", "ParametersTests.cs:51-66", "cosine", {"max_edges": 8, "min_nn": 1, "max_candidates": 16}, "v1000", -0.01, 0.50),
        ("SingleThreadGraphResize", "GraphResizeTests.cs:17-39", "sq_euclid", {"collection_size": 10}, "v5000", 0.85, 1.01),
    ]
    data = {"v2000": v2000, "v1000": v1000, "v5000": v5000}
    rec = []
    for name, cite, metric, kw, fx, lo, hi in cases:
        ix = O.OracleIndex(metric, **kw)
        ids = ix.add(data[fx])
        r = O.self_recall(ix, data[fx], ids)
        assert lo < r < hi, (name, r)
        rec.append({"name": name, "cite": cite, "metric": metric, "params": kw, "fixture": fx, "lo": lo, "hi": hi, "oracle_value": r})
        print(name, r)
    json.dump(rec, open(HERE / "reference_windows.json", "w"), indent=1)


if __name__ == "__main__":
    main()
