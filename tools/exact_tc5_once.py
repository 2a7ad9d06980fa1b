"""One tcgen05 exact top-10 of 10k queries over 1M x 128 (for an ncu launch list of exact_tc5_scan_kernel)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import hnswindex  # noqa: E402

n, nq = 1_000_000, 10_000
rng = np.random.default_rng(1234)
A = rng.standard_normal((16, 128)).astype(np.float32)
x = (rng.standard_normal((n + nq, 16)).astype(np.float32) @ A + 0.25 * rng.standard_normal((n + nq, 128)).astype(np.float32)).astype(np.float32)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_max_edges(2)
ix.set_max_candidates(4)
ix.set_collection_size(n)
ix.set_allow_removals(False)
ix.add(x[:n])
ix.exact_knn(x[n:], 10, tensor_cores=True)
print("fallbacks", ix.last_exact_fallbacks)
