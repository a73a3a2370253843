"""CPU tests of bench.py's contract pieces that do not need a GPU: the reference arm (oracle port on the host cores) prints
one well-formed JSON line, non-zero ranks of a torchrun launch stay silent, and the byte accounting of SURVEY.md App. E is
what DESIGN.md states."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                          timeout=600)


def test_reference_arm_json_line():
    out = _run(["--impl", "reference", "--cells", "48", "--steps", "1", "--warmup", "1"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "FCT steps/sec" and d["unit"] == "steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["dtype"] == "f64"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert "C/OpenMP port" in cb["sample"] or "replicas" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_are_silent():
    out = _run(["--impl", "reference", "--cells", "48", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_byte_accounting():
    sys.path.insert(0, ROOT)
    import bench
    n, N = 4097 ** 2, 4097
    nnz = n + 2 * (2 * N * (N - 1) + (N - 1) ** 2)
    cells = 2 * 4096 ** 2
    assert nnz == 117465089
    # 93.24 GB per FCT step with 18 Jacobi sweeps
    assert abs(bench.step_bytes(n, nnz, cells, 18) / 1e9 - 93.24) < 0.01
    assert bench.cheb_iter_bytes(n, nnz) == 12 * nnz + 4 * n + 40 * n
