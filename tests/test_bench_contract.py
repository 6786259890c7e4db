"""bench.py's output contract on the CPU tier: the reference arm (the C oracle port timed on the host cores) prints
exactly ONE line on stdout, a JSON object with the keys the driver reads and the same `config` the GPU arm reports;
ranks other than 0 print nothing and exit 0; anything a library writes to file descriptor 1 ends up on stderr."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _run(extra_env=None, *flags):
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0", "--batch", "2", *flags], cwd=ROOT, env=env, capture_output=True, text=True,
                          timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same config dict as the GPU arm (the driver compares the two lines)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.config_for("resnet18", 2, 1)
    assert set(d["config"]) >= {"workload", "arch", "batch_per_gpu", "global_batch", "parallelism"}


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout == "", (r.stdout, r.stderr[-1000:])


def test_library_output_on_fd1_goes_to_stderr():
    """NCCL prints its version banner to file descriptor 1 under NCCL_DEBUG=VERSION (set on the GPU boxes): bench.py
    points fd 1 at stderr and writes its line to a private duplicate of the original stdout."""
    code = ("import os, sys\n"
            f"sys.path.insert(0, {ROOT!r})\n"
            "import bench\n"
            "bench.RESULT_OUT = bench._claim_stdout()\n"
            "os.write(1, b'banner from a library\\n')\n"
            "print('python print after the claim')\n"
            "bench.emit({'ok': 1})\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"ok": 1}\n', r.stdout
    assert "banner from a library" in r.stderr and "python print after the claim" in r.stderr
