"""CPU: the JSON contract of `bench.py --impl reference` (the reference arm: oracle port on the host cores, no GPU), on a tiny
sample so it stays within seconds.  The GPU arm of bench.py is exercised on the B200 box by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-sample", "1024"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sites/sec (predict)" and d["unit"] == "sites/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["sites_per_step_per_gpu"] == 1024 and "workload" in d["config"]


def test_roofline_flop_counts_match_the_survey():
    """The algorithmic FLOP counts the secondary bench legs report against are SURVEY 8d's: 113 379 328 conv FLOP per site for the
    shipped human MuRaL-indel configuration, 3 x stage-S + 2 x stage-G wgrad = 22.6 M per site for a MuRaL-snv training step."""
    import bench
    assert bench.unet_flops(8, 7, [1, 4, 5, 5, 5, 2], 8000, True) == 113379328
    assert bench.unet_flops(8, 7, [1, 4, 5, 5, 5, 2], 8000, False) == 113379328 - 2 * 2 * 8000 * 7 * 4 * 4
    assert bench.snv_train_flops(2001) == 3 * 6396008 + 2 * (768 * 2001 + 768 * 201)
    assert abs(bench.snv_train_flops(2001) - 22.6e6) < 0.1e6
