"""CPU: the reference arm of bench.py (`--impl reference`: the oracle port timed on the host cores) prints ONE JSON line
with the keys the driver reads, and the non-zero ranks of a torchrun launch exit 0 without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None, *flags):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                           *flags], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "distill-loss fwd+bwd samples/s" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"] == "soft_kd_logits_b256_c1000_bf16"          # BASELINE.json configs[1]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "256" in cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_do_nothing():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workload_registry_covers_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    names = set(bench.WORKLOADS)
    assert bench.HEADLINE.name == "soft_kd_logits_b256_c1000_bf16"
    for must in ("curkd_early_3layers_b512_f32", "curkd_mid_4layers_b512_f32", "mgd_b512_f32", "saliency_mgd_m1_b512_f32",
                 "lrkd_r64_b512_f32", "wasskd_l1_b512_f32", "wasskd_sinkhorn_b512_f32", "deit_tiny_kd_step_soft_b256_bf16"):
        assert must in names, must
