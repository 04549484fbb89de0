"""Fused attention core (csrc/flash_attn.cu) against softmax(q k^T) v in fp32 on the same bf16 inputs
(utils/network.py:184-186).  Tolerance 1e-2 relative (bf16 probabilities), log-sum-exp 1e-4."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A(lib_built):
    from sap3d_tensorflow_b200 import _abi

    assert _abi.lib.sap3d_device_ok() == 1, _abi.lib.sap3d_last_error()
    return _abi


def rel(a, b):
    a, b = a.detach().float().cpu().reshape(-1), b.detach().float().cpu().reshape(-1)
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def stream():
    return torch.cuda.current_stream().cuda_stream


def make(B, Nq, Nk, dk, dv, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.zeros(B, Nq, 64, device="cuda")
    k = torch.zeros(B, Nk, 64, device="cuda")
    q[..., :dk] = torch.randn(B, Nq, dk, device="cuda", generator=g) * scale
    k[..., :dk] = torch.randn(B, Nk, dk, device="cuda", generator=g) * scale
    v = torch.randn(B, Nk, dv, device="cuda", generator=g)
    return q.bfloat16(), k.bfloat16(), v.bfloat16()


SHAPES = [(2, 300, 200, 16, 128), (1, 256, 3136, 16, 128), (2, 130, 129, 32, 256), (1, 128, 128, 64, 128), (3, 1, 1, 16, 128)]


@pytest.mark.parametrize("B,Nq,Nk,dk,dv", SHAPES)
@pytest.mark.parametrize("scale", [1.0, 3.0])
def test_flash_fwd(A, B, Nq, Nk, dk, dv, scale):
    q, k, v = make(B, Nq, Nk, dk, dv, scale)
    o = torch.full((B, Nq, dv), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, Nq, device="cuda")
    A.check(A.lib.sap3d_flash_attn_fwd(A.ptr(q), A.ptr(k), A.ptr(v), A.ptr(o), A.ptr(lse), B, Nq, Nk, 64, dv, stream()), "flash fwd")
    torch.cuda.synchronize()
    s = q.float() @ k.float().transpose(1, 2)
    ref = torch.softmax(s, -1) @ v.float()
    assert torch.isfinite(o.float()).all()
    assert rel(o, ref) < 1e-2, rel(o, ref)
    assert rel(lse, torch.logsumexp(s, -1)) < 1e-4
