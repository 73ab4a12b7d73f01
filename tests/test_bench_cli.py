"""bench.py contract checks that need no GPU: the reference arm (the unmodified reference from
baseline/_ref when installed, else the oracle port) prints exactly one JSON line with the
agreed keys; the byte models match SURVEY 8(d); both arms name the workload identically."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "melvin"))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "2", "--warmup", "1", "--nx", "128", "--nz", "128"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "grid-point-steps/s" and d["value"] > 0
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["scaling"] == "strong"
    assert d["cpu_baseline"]["kind"] == ("reference" if HAVE_REF else "port") and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]
    # the grid is never shrunk behind the label
    assert d["config"]["grid"] == [128, 128] and "128x128" in d["config"]["workload"]
    assert d["steps_timed"] == 2 and "128x128" in d["cpu_baseline"]["sample"]


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0", "--nx", "64", "--nz", "64"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_byte_models_match_survey():
    sys.path.insert(0, ROOT)
    import bench
    bm = bench.byte_model("kh", 4096, 4096)
    assert bm["S"] == 59_645_040 and bm["I"] == 89_456_640
    assert bm["step"] == 5 * bm["S"] + 8 * bm["I"] == 1_013_878_320
    assert abs(bm["step"] / 4096 ** 2 - 60.4) < 0.05                # B per grid-point-step
    assert abs(bench.byte_model("ddc", 8192, 8192)["step"] / 8192 ** 2 - 138.6) < 0.05
    assert abs(bench.byte_model("tearing", 16384, 16384)["step"] / 16384 ** 2 - 152.9) < 0.05
    assert abs(bench.byte_model("rbc", 4096, 2048)["step"] / (4096 * 2048) - 122.6) < 0.05
    # traffic figures are only reported for the kernel sources they were captured on
    traffic, src = bench.measured_traffic(256, 256)
    assert traffic == {} and src is None


def test_both_arms_use_the_same_workload_string():
    sys.path.insert(0, ROOT)
    import bench
    with open(os.path.join(ROOT, "bench.py")) as fp:
        src = fp.read()
    assert src.count('"workload": workload(') >= 3 and '"scaling": "weak"' not in src
    assert "4096x4096" in bench.workload("kh", 4096, 4096)
