"""
Builds the B200 engine as ONE shared library with nvcc for sm_100a, in-tree:

    hnswindex.net_b200/hnswindex/artifacts/native/linux-x64/HNSWIndex.Native.so

That is the path/name the reference's Python package expects (bindings/bindings.py:15-41), so the
library is a drop-in for it.  Usage:  python hnswindex.net_b200/build.py [--force] [--verbose]
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "build"
OUT_DIR = HERE / "hnswindex" / "artifacts" / "native" / "linux-x64"
OUT = OUT_DIR / "HNSWIndex.Native.so"

SOURCES = ["search_kernels.cu", "wide_kernels.cu", "build_kernels.cu", "exact_kernels.cu", "exact_tf32_kernels.cu", "range_kernels.cu", "remove_kernels.cu", "partition_kernels.cu", "engine.cu", "capi.cu", "snapshot.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-fmad=false",            # metric bit-exactness: no implicit FMA contraction
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2,-Wall,-Wno-unknown-pragmas",
    "-Xptxas", "-v",
]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stamp():
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*")) + [HERE.parent / "include" / "hnsw_b200.h", Path(__file__)]):
        if p.is_file():
            h.update(p.name.encode())
            h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    OBJ.mkdir(exist_ok=True)
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    stamp_file = OBJ / "stamp"
    stamp = _stamp()
    if not force and OUT.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return OUT
    srcs = [s for s in SOURCES if (CSRC / s).exists()]

    def compile_one(src):
        obj = OBJ / (src.rsplit(".", 1)[0] + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / (src + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return str(obj)

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, srcs))
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(OUT)] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp_file.write_text(stamp)
    return OUT


C2_SEARCH_KERNEL = "_ZN4hnsw17knn_search_kernelILi0ELb0ELi7EEEvNS_9GraphViewEPKfiiiiiiiPiPfPjPNS_10QueryStatsES6_i"


def kernel_sass_stamp(mangled=C2_SEARCH_KERNEL):
    """sha256 of one kernel's SASS inside the built library (addresses and encodings stripped): identifies the exact machine
    code an ncu capture was taken from, independent of edits elsewhere in csrc/.  None when cuobjdump or the kernel is missing."""
    import re
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not OUT.exists() or not Path(tool).exists():
        return None
    r = subprocess.run([tool, "-sass", "-fun", mangled, str(OUT)], capture_output=True, text=True)
    if r.returncode != 0:
        return None
    ins = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", r.stdout)
    if len(ins) < 100:
        return None
    return hashlib.sha256("\n".join(ins).encode()).hexdigest()


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
