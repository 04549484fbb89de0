"""Data-parallel correctness on REAL NCCL (needs >= 2 GPUs; skipped on a single-GPU box): two replicas train the flagship graph
through both exchange forms (CUDA-graph replay; bf16 gradients all-reduced in one call after backward = the default, or segment by
segment beside the split backward graphs) and must (a) hold bit-identical weights after every step, (b) have exchanged the SUM of their local gradients (to
bf16 bucket precision), (c) have drawn different dropout masks (per-rank seed).  The gloo tests (test_parallel_cpu.py) cover
the host logic; this covers the collective itself."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("overlap", ["0", "1"], ids=["one all-reduce after backward (default)", "overlapped with the split backward graphs"])
def test_two_nccl_replicas_stay_identical_and_exchange_the_gradient_sum(lib_built, tmp_path, overlap):
    out = str(tmp_path / "dp")
    env = dict(os.environ, SAP3D_OUT=out, SAP3D_DP_OVERLAP=overlap)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tests", "_dp_nccl_worker.py")], env=env, cwd=ROOT, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res = [json.load(open(f"{out}.{k}")) for k in range(2)]
    print(json.dumps(res[0]))
    for d in res:
        assert d["checksums"][0] == d["checksums"][1], d["checksums"]                   # replicas bit-identical after 3 steps
        assert d["exchanged_vs_fp32_sum_rel"] < 1e-2, d["exchanged_vs_fp32_sum_rel"]     # bf16 exchange: 2^-9 per element
        assert d["overlap"] == (overlap == "1")
        assert d["overlap_graphs"] == (1 + d["segments"] if overlap == "1" else 2) and d["segments"] >= 2
        assert d["dropout_seeds"][0] != d["dropout_seeds"][1]
        assert all(l == l for l in d["losses"])
    assert res[0]["losses"] != res[1]["losses"]                                          # different shards
