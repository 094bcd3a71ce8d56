"""CPU: the reference arm of bench.py (the unmodified reference from oracle/_ref, else the oracle port) prints one JSON
line that carries the contract's keys, and the host-side helpers of the module tree agree with the oracle."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "tiny", "--size", "56",
                        "--batch", "4", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "depth_frames_per_sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["global_batch"] == 4 and "workload" in d["config"]
    from oracle.build_ref import ref_available

    assert d["cpu_baseline"]["kind"] == ("reference" if ref_available() else "port")


def test_native_and_reference_arms_share_the_config_object():
    sys.path.insert(0, ROOT)
    import bench

    a = bench.workload_config("vitl", 32, 32, 1, 504, "bf16")
    assert a["global_batch"] == 32 and a["per_gpu_batch"] == 32 and "504" in a["workload"] and "518" in a["workload"]
    b = bench.workload_config("swinv2_large_384", 16, 2, 8, 384, "fp16")
    assert b["parallelism"].startswith("dp8") and "MiDaS" in b["workload"]


@pytest.mark.parametrize("patch,target", [(96, 24), (48, 24), (24, 24), (12, 24), (40, 16), (20, 16), (10, 16), (36, 24), (30, 12)])
def test_window_and_shift_helper_matches_the_oracle(patch, target):
    from muggled_dpt_b200.module_tree import swin_window_and_shift
    from oracle import dpt_oracle as O

    (wh, ww), (sh, sw) = O.swin_window_and_shift((patch, patch), (target, target))
    assert swin_window_and_shift(patch, target) == (wh, sh) == (ww, sw)
