"""sap3d_gemm_nt_batched: `batch` independent C_n = A_n B_n^T products in one launch (per-sample B through a rank-3 tensor map,
row tiles that never span samples) against torch.bmm on the same bf16 operands, at the shapes of the three attention blocks
that run per-sample loops today (utils/network.py:176-180 at 49, 392 and 3136 positions)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [
    # name, batch, M, N, K, rows_b, out_f32
    ("x_4_0 scores 49x49 (padded to 56 cols)", 8, 49, 56, 64, 49, 1),
    ("x_3_1 scores 392x392", 8, 392, 392, 64, 392, 1),
    ("x_3_1 P.V: 392 x 512 over 448 padded keys", 8, 392, 512, 448, 512, 0),
    ("x_2_2 scores 3136x3136", 4, 3136, 3136, 64, 3136, 0),
    ("one sample == the unbatched entry point", 1, 300, 128, 128, 128, 0),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_batched_gemm_nt_matches_bmm(lib_built, case):
    from sap3d_tensorflow_b200 import _abi as A

    _, batch, M, N, K, rows_b, out_f32 = case
    torch.manual_seed(0)
    a = (torch.randn(batch, M, K, device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn(batch, rows_b, K, device="cuda") * 0.5).to(torch.bfloat16)
    c = torch.full((batch, M, N), 7.0, device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    A.check(A.lib.sap3d_gemm_nt_batched(A.ptr(a), K, M * K, A.ptr(b), K, rows_b * K, rows_b, A.ptr(c), N, M * N, M, N, K, batch, out_f32, 0, st),
            "gemm_nt_batched")
    torch.cuda.synchronize()
    ref = torch.bmm(a.float(), b.float().transpose(1, 2))
    got = c.float()
    assert torch.allclose(got[:, :, :rows_b], ref, rtol=2e-2, atol=2e-2)
    err = ((got[:, :, :rows_b] - ref).norm() / ref.norm()).item()
    assert err < (1e-5 if out_f32 else 4e-3), err
    if rows_b < N:
        assert (got[:, :, rows_b:] == 0).all()          # columns against the zero-filled rows of B
    if batch > 1:                                       # and sample n really used B_n, not B_0
        wrong = torch.bmm(a.float(), b[:1].float().expand(batch, -1, -1).transpose(1, 2))
        assert ((got[1:, :, :rows_b] - wrong[1:]).norm() / ref[1:].norm()).item() > 0.5


def test_model_parity_with_batched_attention(lib_built):
    """the engine's attention core with one launch per product for the whole batch (SAP3D_ATTN_BATCHED=1, read at import):
    the graph-level parity tests of the attention graphs must hold with it switched on (child process)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SAP3D_ATTN_BATCHED="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_model_gpu.py"), "-q", "-x", "-m", "gpu", "-k",
                        "forward_parity or training_step_parity or reduces_loss or gradcheck"], env=env, cwd=root, capture_output=True,
                       text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
