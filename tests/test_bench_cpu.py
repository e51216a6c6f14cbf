"""CPU: the host-side legs of bench.py that need no GPU -- the reference arm's JSON contract and the reference PPO_AcM phase timer
(SURVEY 8d iv) -- run here so that a broken baseline leg is seen before the GPU box sees it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_ppo_phase_timer_reports_every_phase():
    sys.path.insert(0, ROOT)
    import bench
    r = bench.reference_ppo_block(transitions=600, limit_s=240)
    assert "unavailable" not in r, r
    assert r["transitions"] >= 600 and r["us_per_transition"] > 0
    assert set(r["phases_ms"]) == {"rollout", "critic_fit", "actor_epochs", "add_buffer", "acm_update"}
    assert all(v > 0 for v in r["phases_ms"].values())
    assert abs(sum(r["phases_ms"].values()) - r["ms_per_iteration"]) < 1e-6 * r["ms_per_iteration"] + 1.0


def test_reference_rollout_step_timer():
    sys.path.insert(0, ROOT)
    import bench
    r = bench.reference_rollout_block(transitions=600, limit_s=240)
    assert "unavailable" not in r, r
    assert r["transitions"] > 0 and r["us_per_transition"] > 0 and r["transitions_per_s"] > 0


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "updates/s" and line["value"] > 0
    assert line["steps"] == 1 and line["warmup"] == 1 and line["ms_per_step"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ms_per_step is a measured time of real steps: value = updates per step / that time
    assert abs(line["value"] - line["config"]["updates_per_step"] / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"]
