"""bench.py contract checks that need no GPU: the reference arm (oracle port on host cores)
prints exactly one JSON line with the agreed keys; the byte model matches SURVEY 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "2", "--warmup", "1", "--nx", "128", "--nz", "128"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "grid-point-steps/s" and d["value"] > 0
    assert d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0", "--nx", "64", "--nz", "64"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_byte_model_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    bm = bench.byte_model(4096, 4096)
    assert bm["S"] == 59_645_040 and bm["I"] == 89_456_640
    assert bm["step"] == 5 * bm["S"] + 8 * bm["I"] == 1_013_878_320
    assert abs(bm["step"] / 4096 ** 2 - 60.4) < 0.05                # B per grid-point-step
    traffic, src = bench.measured_traffic(4096, 4096)
    assert set(traffic) == {"mlv_x_inverse", "mlv_advect_z", "mlv_x_forward"} and "ncu" in src
    assert bench.measured_traffic(256, 256) == ({}, None)
