"""bench.py's reference arm runs on the host cores only, so its JSON line can be checked without a GPU:
the keys the driver reads, the metric / unit / config it must share with the GPU arm, and the
tier's extra objects (cpu_baseline, zero-copy e2e)."""
import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_reference_arm_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["metric"] == "point-pairs/s (1M-pt ICP)" and line["unit"] == "point-pairs/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f32"
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 1e6   # a kd-tree ICP on any recent core does millions of point pairs per second


def test_matchers_per_gpu_respects_the_cores():
    """bench.py runs several matchers per GPU, one polling host thread each: never more than this rank's share of
    the host cores minus one (8 ranks on a 32-core box get three), and the environment override wins."""
    import os
    code = "import bench; print(bench.MATCHERS_PER_GPU)"
    env = {k: v for k, v in os.environ.items() if k not in ("WAVE_BENCH_MATCHERS", "LOCAL_WORLD_SIZE", "WORLD_SIZE")}
    cores = os.cpu_count() or 1
    one = int(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, env=env).stdout)
    assert one == max(1, min(6, cores - 1))
    many = int(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT,
                              env={**env, "LOCAL_WORLD_SIZE": "8"}).stdout)
    assert many == max(1, min(6, cores // 8 - 1))
    forced = int(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT,
                                env={**env, "WAVE_BENCH_MATCHERS": "2", "LOCAL_WORLD_SIZE": "8"}).stdout)
    assert forced == 2
