"""bench.py's reference arm runs on the CPU: check its one-line JSON contract here (the GPU arm's line is checked
by the driver on the GPU box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "groth16_withdraw_proofs_per_sec" and line["unit"] == "proofs/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["steps"] == 1 and line["warmup"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_launch_list_tool_reads_the_committed_ncu_csv(tmp_path):
    """tools/launch_list_summary.py turns the committed ncu launch list into the table under profiles/: the dominant kernel of
    the bench command must come out on top (this is the evidence the roofline share is checked against)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "ll.md"
    subprocess.run([sys.executable, os.path.join(root, "tools", "launch_list_summary.py"),
                    os.path.join(root, "profiles", "r2_launches_bench_steps2.csv"), str(out), "t"], check=True)
    rows = [l for l in out.read_text().splitlines() if l.startswith("| `")]
    assert rows and rows[0].startswith("| `k_bucket_acc_sm1"), rows[:2]
