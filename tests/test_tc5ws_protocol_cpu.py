"""The mbarrier protocol of the (experimental) warp-specialised tcgen05 kernel, model-checked under random schedules:
no deadlock, no wait satisfied by an aliased phase, no buffer overwritten while in use (tools/tc5ws_protocol_model.py)."""
import importlib.util
from pathlib import Path


def test_warp_specialised_protocol_under_random_schedules():
    spec = importlib.util.spec_from_file_location("tc5ws_model", Path(__file__).resolve().parent.parent / "tools" / "tc5ws_protocol_model.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for seed in range(600):
        for ntiles in (0, 1, 2, 3, 7):
            assert mod.run(ntiles, seed * 31 + ntiles) >= 0
