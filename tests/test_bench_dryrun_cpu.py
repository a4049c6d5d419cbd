"""CPU: bench.py's host logic end to end (tools/bench_dryrun.py: recorder instead of the CUDA library, inert torch.cuda) in a
subprocess -- model build, CUDA-graph capture flow, timed loop, e2e prefetch loop, roofline passes and the JSON contract keys."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
        "config", "e2e", "gpu_launches", "clocks"}


@pytest.mark.parametrize("mode", [["train"], ["caption"], ["train", "reference", "nosecondary"], ["train", "large_pretrain224", "nosecondary"]],
                         ids=lambda m: "+".join(m))
def test_bench_host_logic(mode):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_dryrun.py")] + mode, cwd=ROOT, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert KEYS <= set(out), KEYS - set(out)
    assert set(out["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and out["e2e"]["h2d_bytes_per_step"] > 0
    assert "workload" in out["config"] and out["n_gpus"] == 1
    if mode[0] == "train":
        want = "Prismer-LARGE pretrain images/sec" if "large_pretrain224" in mode else "Prismer-BASE caption-train images/sec"
        assert out["metric"] == want and {"roofline", "step_mfu"} <= set(out)
        if "nosecondary" not in mode:          # the driver-run line carries greedy captions/s (+ decode roofline) and the fp32-input e2e
            assert out["secondary"]["unit"] == "captions/s" and out["secondary"]["roofline"]["bound"] == "hbm"
            assert out["e2e_reference_inputs"]["h2d_bytes_per_step"] > 10 * out["e2e"]["h2d_bytes_per_step"]
        assert set(out["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
