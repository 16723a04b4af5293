"""bench.py pieces that need no GPU: the configuration both arms print, and the reference arm end to end on a tiny
sample (the arm the driver runs with --impl reference)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_config_is_a_pure_function_of_the_workload():
    import bench
    from gridcodegenerator_b200 import load_named_robot
    robot = load_named_robot("iiwa14")
    a, nsets, L = bench.make_config("iiwa14", robot, "fd_grad", 65536)
    b, _, _ = bench.make_config("iiwa14", robot, "fd_grad", 65536)
    assert a == b and a["launches_per_step"] == L == bench.launches_per_step(robot, "fd_grad", 65536)
    assert L * 35e-6 * 20 > 0.02                 # K = 20 steps time at least 20 ms of 35 us launches
    assert nsets * 4 * 65536 * (21 + 98) > 126e6  # buffers rotate over more than the L2
    atlas = load_named_robot("atlas")
    assert 1 <= bench.launches_per_step(atlas, "fd_grad", 65536) < L


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-sample", "16"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "fd_grad_evals_per_s_iiwa14_N65536"
    assert line["unit"] == "evals/s" and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    import bench
    from gridcodegenerator_b200 import load_named_robot
    assert line["config"] == bench.make_config("iiwa14", load_named_robot("iiwa14"), "fd_grad", 65536)[0]
