"""bench.py contract checks that run without a GPU: the reference arm prints ONE JSON line with the keys the
driver reads, every workload names a kernel path, and the b200 arm refuses to run (loudly) without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=600)


def test_reference_arm_prints_one_json_line_with_contract_keys():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "cfg2")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Msps" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "cfg2",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workload_table_is_consistent():
    sys.path.insert(0, ROOT)
    import bench
    assert "cfg3" in bench.WORKLOADS and bench.WORKLOADS["cfg3"]["nchans"] == 1024 and bench.WORKLOADS["cfg3"]["ntaps"] == 256
    for name, w in bench.WORKLOADS.items():
        assert "desc" in w and w["log2n"] >= 20, name
        if w.get("kind") is None:
            assert w["out"] in ("fm", "iq", "iq+fm"), name
            assert w["ntaps"] >= 1 and w["nchans"] in (64, 256, 1024), name
