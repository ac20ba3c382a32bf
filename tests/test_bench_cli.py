"""bench.py contract checks that need no GPU: the reference arm (stock torch-DDP on gloo/CPU via the oracle)
prints exactly one JSON line with the agreed keys, also when launched the way the driver launches it for
N > 1 (under torchrun: rank 0 works, the other ranks exit 0 silently)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *args, timeout=240):
    env = dict(os.environ, **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, env=env, cwd=ROOT)


@pytest.mark.timeout(300)
def test_reference_arm_under_torchrun_environment():
    torchrun_env = {"WORLD_SIZE": "2", "LOCAL_WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29999",
                    "TORCHELASTIC_USE_AGENT_STORE": "True", "TORCHELASTIC_RESTART_COUNT": "0",
                    "TORCHELASTIC_RUN_ID": "none"}
    other = _run(dict(torchrun_env, RANK="1", LOCAL_RANK="1"), "--impl", "reference", "--gpus", "2", "--steps", "5",
                 "--warmup", "1")
    assert other.returncode == 0 and other.stdout.strip() == ""  # ranks != 0 do no work
    r0 = _run(dict(torchrun_env, RANK="0", LOCAL_RANK="0"), "--impl", "reference", "--gpus", "2", "--steps", "5",
              "--warmup", "1")
    assert r0.returncode == 0, r0.stderr[-2000:]
    lines = [l for l in r0.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["metric"] == "samples/sec"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[1]" in line["config"]["workload"]


@pytest.mark.timeout(120)
def test_product_arm_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU failure mode")
    out = _run({}, "--gpus", "1", "--steps", "3", "--warmup", "1", timeout=100)
    assert out.returncode != 0
    assert "no CPU fallback" in out.stdout or "no CUDA" in out.stdout


def test_bench_helpers_blocks_and_identical_config():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    # the driver passes --steps 20: thirty blocks of 20 replace one sub-millisecond sample
    assert bench.blocks_for("toy", 20, None) == 30
    assert bench.blocks_for("toy", 2000, None) == 1
    assert bench.blocks_for("resnet50", 20, None) == 2
    assert bench.blocks_for("gpt2m", 10, 4) == 4
    # `config` is the same dict in every arm (implementation details live in impl_config)
    cfg = bench.bench_config("toy", 8, 32)
    assert cfg == {"workload": bench.WORKLOAD_DESC["toy"], "global_batch": 256, "parallelism": "dp8", "l2": bench.L2_NOTE}
    st = bench.block_stats([3.0, 1.0, 2.0], 20)
    assert st == {"timed_blocks": 3, "block_ms_min": 1.0, "block_ms_median": 2.0, "block_ms_max": 3.0, "steps_per_block": 20}
