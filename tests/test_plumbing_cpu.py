"""CPU-side plumbing checks: the drop-in registration against the reference's registry (only where the
reference tree is present, i.e. in the build container), and bench.py's contract lines."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "MetLib")), reason="reference tree not present on this machine")
def test_register_with_metlib_swaps_the_detectors_in_the_reference_registry():
    """MetLib.get_detector (MetLib/__init__.py:34-46) must hand out the CUDA classes after registration, so that an
    unmodified MetDetPy.detect_video (MetDetPy.py:137-142) constructs them.  No detector is created (no GPU here)."""
    code = (
        "import sys; sys.path[:0] = [%r, %r, %r]\n"
        "import MetLib, metdetpy_b200\n"
        "from metdetpy_b200.detector import M3Detector, ClassicDetector\n"
        "before = MetLib.get_detector('M3Detector')\n"
        "assert before is not M3Detector\n"
        "metdetpy_b200.register_with_metlib()\n"
        "assert MetLib.get_detector('M3Detector') is M3Detector\n"
        "assert MetLib.get_detector('ClassicDetector') is ClassicDetector\n"
        "import MetLib.Detector as D\n"
        "assert D.M3Detector is M3Detector and D.ClassicDetector is ClassicDetector\n"
        "import inspect\n"
        "ref_args = list(inspect.signature(before.__init__).parameters)[:7]\n"
        "our_args = list(inspect.signature(M3Detector.__init__).parameters)[:7]\n"
        "assert ref_args == our_args, (ref_args, our_args)\n"
        "print('ok')\n" % (os.path.join(REPO, "tests", "golden", "shims"), REF, REPO))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_bench_reference_arm_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours): one JSON line with the contract keys."""
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--width", "320", "--height", "192", "--window", "5"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["dtype"] == "u8"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_bench_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_without_cuda_fails_loudly():
    """The product arm has no CPU fallback: without a CUDA device bench.py must exit non-zero, not print a number."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not r.stdout.strip().startswith("{")


def test_bench_roofline_traffic_comes_from_the_committed_capture():
    """roofline.traffic is read from profiles/chain_traffic.json (written by profiles/ncu_summary.py --traffic from an ncu
    --set full capture), never a literal; a configuration without a capture gives null."""
    import json
    sys.path.insert(0, REPO)
    import bench
    tab = json.load(open(os.path.join(REPO, "profiles", "chain_traffic.json")))
    assert "3840x2160_n30_b512" in tab
    v = bench._traffic_from_profile(3840, 2160, 30, 512)
    e = tab["3840x2160_n30_b512"]
    assert v == e["bytes_per_launch"] and abs(v * e["launches_per_batch"] - e["bytes_per_batch"]) < 1
    # the chain moves less than the algorithmic 2*H*W per frame (mask bytes are only rewritten where they change)
    assert 0.5 * 2 * 3840 * 2160 * 512 < e["bytes_per_batch"] < 2 * 3840 * 2160 * 512
    assert bench._traffic_from_profile(1234, 567, 8, 9) is None


def test_bench_numa_binding_degrades_gracefully():
    sys.path.insert(0, REPO)
    import bench
    r = bench.bind_near_gpu(0)  # no NVML device here: must report, not raise
    assert r["bound"] is False and "why" in r
