"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) prints ONE JSON line with the
agreed keys, and the product arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")  # the reference arm is a CPU arm; the product arm must refuse
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=env, cwd=ROOT)


def test_reference_arm_json_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "chamfer_pairs_per_s_10k" and d["unit"] == "pairs/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "fp32"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_product_arm_refuses_without_gpu():
    p = _run("--steps", "1", "--warmup", "0")
    assert "no CUDA device" in (p.stdout + p.stderr)
    assert not any(l.strip().startswith("{") for l in p.stdout.splitlines()), "no result line without a GPU"
