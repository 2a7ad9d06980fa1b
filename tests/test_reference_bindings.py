"""
The reference's OWN Python package and its OWN tests, unchanged, against HNSWIndex.Native.so.

SURVEY §7 step 6 / §8(b): the acceptance gate of the drop-in boundary.  `__graft_entry__.build()` installs the
reference's pure-Python package byte for byte into baseline/_ref/ (git-ignored, travels to the GPU box):
    bindings/__init__.py, bindings/bindings.py   -> baseline/_ref/hnswindex/
    bindings/__tests__/*_test.py                 -> baseline/_ref/hnswindex_tests/
Here that package is staged in a temp dir with our library at the place its loader looks
(bindings.py:15-41: <package>/artifacts/native/linux-x64/HNSWIndex.Native.so) and run in a SEPARATE interpreter whose
sys.path holds neither this repo's own `hnswindex` mirror nor the oracle, so every call goes
reference bindings.py -> ctypes -> C ABI -> CUDA.
  * CPU (not gpu): the package imports (every argtypes/restype assignment of bindings.py:46-119 resolves against the
    library), the copies are byte-identical to /root/reference, and Index.add fails loudly without a GPU.
  * GPU: every test_* of metric_test.py, parameters_test.py and recall_test.py passes under pytest.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF_INSTALL = ROOT / "baseline" / "_ref"
REFERENCE = Path("/root/reference")


def stage_reference_package(tmp: Path) -> Path:
    """tmp/site/hnswindex = the reference package + our .so where bindings.py:35-41 expects it; tmp/ref_tests = its tests."""
    import __graft_entry__ as ge
    ge.build()
    if not (REF_INSTALL / "hnswindex" / "bindings.py").exists():
        pytest.skip("baseline/_ref is absent: build() installs it where /root/reference exists and it travels with the snapshot")
    import hnswindex as ours
    site = tmp / "site"
    shutil.copytree(REF_INSTALL / "hnswindex", site / "hnswindex")
    native = site / "hnswindex" / "artifacts" / "native" / "linux-x64"
    native.mkdir(parents=True)
    os.symlink(ours.LIB_PATH, native / "HNSWIndex.Native.so")
    shutil.copytree(REF_INSTALL / "hnswindex_tests", tmp / "ref_tests")
    return site


def clean_env(site: Path) -> dict:
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    env["PYTHONPATH"] = str(site)
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    return env


def run_py(site: Path, code: str, timeout=300):
    return subprocess.run([sys.executable, "-c", code], env=clean_env(site), cwd=str(site), capture_output=True, text=True, timeout=timeout)


def test_installed_copy_is_the_unmodified_reference(tmp_path):
    stage_reference_package(tmp_path)
    if not REFERENCE.is_dir():
        pytest.skip("/root/reference only exists in the build container")
    pairs = [(REFERENCE / "bindings" / n, REF_INSTALL / "hnswindex" / n) for n in ("__init__.py", "bindings.py")]
    pairs += [(f, REF_INSTALL / "hnswindex_tests" / f.name) for f in sorted((REFERENCE / "bindings" / "__tests__").glob("*_test.py"))]
    assert len(pairs) == 5
    for a, b in pairs:
        assert hashlib.sha256(a.read_bytes()).hexdigest() == hashlib.sha256(b.read_bytes()).hexdigest(), b


def test_reference_package_loads_our_library(tmp_path):
    site = stage_reference_package(tmp_path)
    r = run_py(site, """
import os, sys
import hnswindex, hnswindex.bindings as b
assert os.path.realpath(hnswindex.__file__).startswith(os.path.realpath(os.getcwd())), hnswindex.__file__
assert b.lib._name.endswith('artifacts/native/linux-x64/HNSWIndex.Native.so'), b.lib._name
print(os.path.realpath(b.lib._name))
ix = hnswindex.Index(8)
# the setters of bindings.py:189-354 only touch the pending parameters (HNSWIndexExports.cs:219-273): no device needed
ix.set_collection_size(100); ix.set_max_edges(8); ix.set_max_candidates(50); ix.set_remove_max_candidates(50)
ix.set_distribution_rate(0.5); ix.set_random_seed(7); ix.set_min_nn(3); ix.set_allow_removals(False)
assert b._last_error() == '' or isinstance(b._last_error(), str)
""")
    assert r.returncode == 0, r.stdout + r.stderr
    import hnswindex as ours
    assert os.path.realpath(r.stdout.strip().splitlines()[-1]) == os.path.realpath(ours.LIB_PATH)


def test_reference_package_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    site = stage_reference_package(tmp_path)
    r = run_py(site, """
import numpy as np, hnswindex
ix = hnswindex.Index(8)
try:
    ix.add(np.zeros((4, 8), dtype=np.float32))
except RuntimeError as e:
    print('RAISED', e)
""")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "RAISED hnsw_create failed" in r.stdout, r.stdout


@pytest.mark.gpu
def test_reference_python_tests_pass_unchanged(tmp_path):
    """bindings/__tests__/{metric,parameters,recall}_test.py, collected and run by pytest exactly as the files are."""
    site = stage_reference_package(tmp_path)
    r = subprocess.run([sys.executable, "-m", "pytest", str(tmp_path / "ref_tests"), "-q", "-p", "no:cacheprovider", "--rootdir", str(tmp_path)],
                       env=clean_env(site), cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    log = ROOT / "gpurun_out"
    if log.is_dir():
        (log / "reference_python_tests.log").write_text(out)
    assert r.returncode == 0, out[-4000:]
    assert "12 passed" in out, out[-2000:]
    # and the library those tests ran against is ours (loaded through the reference's own loader)
    r2 = run_py(site, "import os, hnswindex.bindings as b; print(os.path.realpath(b.lib._name))")
    import hnswindex as ours
    assert os.path.realpath(r2.stdout.strip()) == os.path.realpath(ours.LIB_PATH)


@pytest.mark.gpu
def test_reference_package_answers_equal_the_mirror(tmp_path):
    """Same data through the reference's Index class and through this repo's mirror of it: identical ids and distance bits.
    Items are added one call at a time -- the deterministic, serial order of parameters_test.py:60-81 (a batched Add links
    its items concurrently, like the reference's Parallel.For, and two such builds need not be the same graph)."""
    site = stage_reference_package(tmp_path)
    out_npz = tmp_path / "ref_out.npz"
    r = run_py(site, f"""
import numpy as np, hnswindex
rng = np.random.default_rng(11)
data = rng.random((700, 96), dtype=np.float32); q = rng.random((200, 96), dtype=np.float32)
ix = hnswindex.Index(96, metric='cosine'); ix.set_max_candidates(64); ix.set_min_nn(20)
ids = np.array([ix.add([v])[0] for v in data])
ki, kd = ix.knn_query(q, 10)
ri, rd = ix.range_query(q[:20], 0.21)
ix.remove(ids[:1])
ki2, kd2 = ix.knn_query(q, 10)
np.savez({str(out_npz)!r}, ids=ids, ki=ki, kd=kd, ki2=ki2, kd2=kd2, rn=np.array([len(x) for x in ri]),
         ri=np.concatenate(ri) if len(ri) else np.zeros(0), rd=np.concatenate(rd) if len(rd) else np.zeros(0))
""")
    assert r.returncode == 0, r.stdout + r.stderr
    import numpy as np
    import hnswindex
    ref = np.load(out_npz)
    rng = np.random.default_rng(11)
    data = rng.random((700, 96), dtype=np.float32); q = rng.random((200, 96), dtype=np.float32)
    ix = hnswindex.Index(96, metric="cosine"); ix.set_max_candidates(64); ix.set_min_nn(20)
    ids = np.array([ix.add([v])[0] for v in data])
    ki, kd = ix.knn_query(q, 10)
    ri, rd = ix.range_query(q[:20], 0.21)
    ix.remove(ids[:1])
    ki2, kd2 = ix.knn_query(q, 10)
    assert ref["ki"].dtype == np.int32 and ref["kd"].dtype == np.float32 and ref["ki"].shape == (200, 10)   # bindings.py:506-521
    assert np.array_equal(ref["ids"], ids)
    assert np.array_equal(ref["ki"], ki) and np.array_equal(ref["kd"].view(np.uint32), kd.view(np.uint32))
    assert np.array_equal(ref["ki2"], ki2) and np.array_equal(ref["kd2"].view(np.uint32), kd2.view(np.uint32))
    assert np.array_equal(ref["rn"], np.array([len(x) for x in ri])) and ref["rn"].sum() > 0
    assert np.array_equal(ref["ri"], np.concatenate(ri)) and np.array_equal(ref["rd"], np.concatenate(rd))


@pytest.mark.gpu
def test_reference_readme_example_runs_verbatim(tmp_path):
    """README.md:80-96 of the reference (BASELINE config[0]): the example as printed there, through the reference's package."""
    site = stage_reference_package(tmp_path)
    example = REF_INSTALL / "readme_example.py"
    if not example.exists():
        pytest.skip("baseline/_ref/readme_example.py is absent")
    code = example.read_text(encoding="utf-8")
    assert "index.knn_query(vectors, k=1)" in code and "Index(dim=128" in code
    r = run_py(site, code + """
import numpy as np
assert ids.shape == (2000, 1) and distances.shape == (2000, 1) and ids.dtype == np.int32 and distances.dtype == np.float32
hit = ids[:, 0] == np.arange(2000)
print('SELF_RECALL', hit.mean(), float(np.abs(distances[hit, 0]).max()))
""")
    assert r.returncode == 0, r.stdout + r.stderr
    rec, dmax = [float(x) for x in r.stdout.strip().splitlines()[-1].split()[1:]]
    assert rec > 0.85 and dmax < 1e-5            # recall_test.py:15's bar; a vector is at distance 0 from itself
