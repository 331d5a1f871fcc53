"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the keys the driver reads, the
product arm refuses to run without a CUDA device (no CPU fallback), and the host-core detection respects cgroup quotas."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout,
                          env=e, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--res", "64"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                                      # exactly one line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    p = _run(["--steps", "1", "--warmup", "1"], timeout=300)
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout)


def test_usable_cores_respects_affinity():
    sys.path.insert(0, ROOT)
    import bench
    n = bench.usable_cores()
    assert 1 <= n <= (os.cpu_count() or 1)
    try:
        assert n <= len(os.sched_getaffinity(0))
    except AttributeError:
        pass
