"""bench.py contract checks that need no GPU: the reference arm (CPU port of the reference path) prints one
JSON line with the agreed keys, and the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import helpers


def _run(args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(helpers.ROOT, "bench.py")] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          text=True, timeout=timeout)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--unique", "16"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "GB/s" and j["higher_is_better"] is True and j["value"] > 0
    assert j["metric"] == "decompressed GB/s on 256Kx64KiB brotli batch"
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1", "--warmup", "1", "--streams", "64", "--unique", "16", "--no-e2e", "--no-cpu"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
