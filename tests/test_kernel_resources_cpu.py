"""Compile-time canaries for the hot kernels (CPU: reads what nvcc / ptxas reported while the library was built).

Round 2 lost a factor of two on the construction benchmark to a change that had nothing to do with the kernel it hit: ptxas,
held to 96 registers by a launch bound, compiled `build_select_kernel` in a mode that reported no spills but ran 2.4x slower
(DESIGN.md 3, K4-K6).  The kernel now has register headroom; these checks keep it -- and the headline search kernel -- from
drifting back to such an edge unnoticed: register counts inside the residency the launchers assume, no spills, no stack frame,
and the headline kernel's machine code still the one the committed ncu capture measured (or the bench line says so itself).
"""
import json
import re
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "hnswindex.net_b200"
sys.path.insert(0, str(ROOT))


def _ptxas_report(source):
    import __graft_entry__ as ge
    ge.build()
    log = PKG / "build" / (source + ".log")
    if not log.exists():
        sys.path.insert(0, str(PKG))
        import build
        build.build(force=True)
    text = log.read_text()
    out = {}
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                         r"ptxas info\s+: Used (\d+) registers", text):
        out[m.group(1)] = {"stack": int(m.group(2)), "spill_st": int(m.group(3)), "spill_ld": int(m.group(4)), "regs": int(m.group(5))}
    assert out, f"no ptxas -v output in {log}"
    return out


def _one(report, needle):
    hits = [k for k in report if needle in k]
    assert len(hits) == 1, (needle, hits)
    return report[hits[0]]


def test_headline_search_kernel_fits_five_blocks_without_spills():
    # knn_search_kernel<sq_euclid, base layer, short rows | small list | 16-bit tags>: 5 blocks x 128 threads per SM need <= 102
    r = _one(_ptxas_report("search_kernels.cu"), "knn_search_kernelILi0ELb0ELi7E")
    assert r["regs"] <= 96 and r["stack"] == 0 and r["spill_st"] == 0 and r["spill_ld"] == 0, r


@pytest.mark.parametrize("needle,max_regs", [
    ("build_select_kernelILi0ELi141E", 120),   # C5 shape, 128-float vectors: 4 blocks x 128 threads allow 128; keep headroom
    ("build_select_kernelILi0ELi13E", 128),    # C5 shape, other dimensions
    ("build_link_kernelILi0ELi129E", 128),
])
def test_construction_kernels_have_register_headroom_and_no_spills(needle, max_regs):
    r = _one(_ptxas_report("build_kernels.cu"), needle)
    assert r["regs"] <= max_regs and r["stack"] == 0 and r["spill_st"] == 0 and r["spill_ld"] == 0, r


def test_no_kernel_of_the_search_and_build_units_spills_more_than_a_few_words():
    # a few cosine / unit-cosine instantiations at the 96-register bound of 5 blocks per SM spill one to three 4-byte values;
    # the CTA-per-query kernels run at 64 registers with up to 68 bytes spilled outside their hot loops (measured against
    # larger budgets, wide_kernels.cu) -- anything beyond that is a change worth looking at
    for unit, bound in (("search_kernels.cu", 32), ("build_kernels.cu", 0), ("wide_kernels.cu", 96)):
        bad = {k: v for k, v in _ptxas_report(unit).items() if v["spill_st"] > bound or v["spill_ld"] > bound}
        assert not bad, bad


def test_traffic_figures_name_the_machine_code_they_were_measured_on():
    # profiles/search_kernel_traffic*.json carry the sha256 of the captured kernel's SASS; bench.py prints the DRAM figure only when
    # the library it runs hashes the same.  Here: the records are well-formed and the hash function still finds both kernels.
    sys.path.insert(0, str(PKG))
    import build
    for name in ("search_kernel_traffic.json", "search_kernel_traffic_c3.json"):
        rec = json.loads((ROOT / "profiles" / name).read_text())
        assert re.fullmatch(r"[0-9a-f]{64}", rec["kernel_sass_sha256"]) and rec["dram_bytes_per_launch"] > 0
        now = build.kernel_sass_stamp(rec["kernel_mangled"])
        if now is None:
            pytest.skip("cuobjdump not available")
        assert re.fullmatch(r"[0-9a-f]{64}", now)
