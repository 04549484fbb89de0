"""Op-level parity of the CUDA kernels (through the C ABI) against the torch oracle primitives."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tf_semantics as tfs  # noqa: E402


@pytest.fixture(scope="module")
def A(lib_built):
    from sap3d_tensorflow_b200 import _abi

    assert _abi.lib.sap3d_device_ok() == 1, _abi.lib.sap3d_last_error()
    return _abi


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def stream():
    return torch.cuda.current_stream().cuda_stream


DT = [("f32", torch.float32, 2e-5), ("bf16", torch.bfloat16, 1.5e-2)]


def _stats(x):
    xf = x.float().reshape(-1, x.shape[-1])
    return torch.stack([xf.sum(0), (xf * xf).sum(0)], 0).reshape(1, 2, -1).contiguous()


@pytest.mark.parametrize("dtn,tdt,tol", DT)
@pytest.mark.parametrize("pattern", ["relu_bn", "bn_plus_relu_bn", "act_plus_relu_bn", "relu_bn_plus_t", "relu_bn_plus_bn"])
@pytest.mark.parametrize("shape", [(2, 3, 5, 6, 64), (8, 2, 7, 7, 1024), (3, 2, 7, 7, 72), (2, 8, 28, 28, 64), (2, 4, 40, 64, 256)],
                         ids=["slab", "slab_stage3_tail", "slab_ragged", "coop_two_level", "three_launch"])
def test_affine_act_fwd_bwd(A, dtn, tdt, tol, pattern, shape):
    """<= 1024 positions (stage 3 of the backbone): the single-block-per-16-channels slab kernel; small tensors the single
    cooperative backward launch; > 4 M elements the reduce/finalize/apply triple"""
    if shape[-1] == 64 and shape[1] == 3:
        pass
    elif pattern not in ("bn_plus_relu_bn", "relu_bn_plus_t") and shape not in ((8, 2, 7, 7, 1024), (3, 2, 7, 7, 72)):
        pytest.skip("large shapes: two representative patterns")
    torch.manual_seed(0)
    dev = "cuda"
    dt = A.BF16 if dtn == "bf16" else A.F32
    N, D, H, W, Cc = shape
    P = N * D * H * W
    a = (torch.randn(N, D, H, W, Cc, device=dev) * 1.5 + 0.7).to(tdt)
    b = (torch.randn(N, D, H, W, Cc, device=dev) * 0.8 - 0.2).to(tdt)
    g1, b1 = torch.rand(Cc, device=dev) + 0.5, torch.randn(Cc, device=dev) * 0.1
    g2, b2 = torch.rand(Cc, device=dev) + 0.5, torch.randn(Cc, device=dev) * 0.1
    dy = torch.randn(N, D, H, W, Cc, device=dev).to(tdt)
    relu1, relu2, relu_out, use_b, bn2 = {
        "relu_bn": (1, 0, 0, False, False),
        "bn_plus_relu_bn": (1, 1, 0, True, True),       # ST_B
        "act_plus_relu_bn": (1, 0, 0, True, False),     # ST_C
        "relu_bn_plus_t": (0, 0, 1, True, False),       # identity shortcut
        "relu_bn_plus_bn": (0, 0, 1, True, True),       # projection shortcut
    }[pattern]
    f = lambda: torch.empty(Cc, device=dev)  # noqa: E731
    s1, t1, m1, r1, s2, t2, m2, r2 = f(), f(), f(), f(), f(), f(), f(), f()
    mm, mv = torch.zeros(Cc, device=dev), torch.ones(Cc, device=dev)
    sa = _stats(a)
    A.check(A.lib.sap3d_bn_finalize(A.ptr(sa), 1, Cc, float(P), A.ptr(g1), A.ptr(b1), A.ptr(mm), A.ptr(mv), 1, 0.99, 1e-3,
                                    A.ptr(s1), A.ptr(t1), A.ptr(m1), A.ptr(r1), stream()), "fin1")
    if bn2:
        sb = _stats(b)
        A.check(A.lib.sap3d_bn_finalize(A.ptr(sb), 1, Cc, float(P), A.ptr(g2), A.ptr(b2), A.ptr(mm), A.ptr(mv), 1, 0.99, 1e-3,
                                        A.ptr(s2), A.ptr(t2), A.ptr(m2), A.ptr(r2), stream()), "fin2")
    y = torch.empty_like(a)
    A.check(A.lib.sap3d_affine_act(dt, A.ptr(a), A.ptr(s1), A.ptr(t1), relu1, A.ptr(b) if use_b else None,
                                   A.ptr(s2) if bn2 else None, A.ptr(t2) if bn2 else None, relu2, relu_out, A.ptr(y), P, Cc, 0,
                                   stream()), "apply")
    # oracle
    af, bf = a.float().requires_grad_(True), b.float().requires_grad_(True)
    g1r, b1r, g2r, b2r = [t.clone().requires_grad_(True) for t in (g1, b1, g2, b2)]
    z1, _, _ = tfs.batch_norm(af, g1r, b1r, mm, mv, True)
    z1 = torch.relu(z1) if relu1 else z1
    if use_b:
        z2 = tfs.batch_norm(bf, g2r, b2r, mm, mv, True)[0] if bn2 else bf
        z2 = torch.relu(z2) if relu2 else z2
        z1 = z1 + z2
    ref = torch.relu(z1) if relu_out else z1
    assert rel(y, ref) < tol
    ref.backward(dy.float())
    da, db = torch.empty_like(a), torch.empty_like(b)
    dg1, db1, dg2, db2 = [torch.zeros(Cc, device=dev) for _ in range(4)]
    ws = torch.zeros(A.lib.sap3d_affine_act_bwd_workspace(Cc) // 4 + 16, device=dev)
    A.check(A.lib.sap3d_affine_act_bwd(dt, A.ptr(dy), A.ptr(a), A.ptr(s1), A.ptr(t1), A.ptr(m1), A.ptr(r1), relu1,
                                       A.ptr(b) if use_b else None, A.ptr(s2) if bn2 else None, A.ptr(t2) if bn2 else None,
                                       A.ptr(m2) if bn2 else None, A.ptr(r2) if bn2 else None, relu2, relu_out, P, Cc,
                                       A.ptr(da), 0, A.ptr(db) if use_b else None, 0, A.ptr(dg1), A.ptr(db1),
                                       A.ptr(dg2) if bn2 else None, A.ptr(db2) if bn2 else None, A.ptr(ws), stream()), "bwd")
    torch.cuda.synchronize()
    if P > 10000:   # 0.8 - 5.2 M elements: ONE ReLU-mask flip at a |z| ~ 1e-7 element is sqrt(1/5e6) = 4e-4 relative
        tol = max(tol, 1e-3)
    assert rel(da, af.grad) < 3 * tol
    assert rel(dg1, g1r.grad) < 3 * tol and rel(db1, b1r.grad) < 3 * tol
    if use_b:
        assert rel(db, bf.grad) < 3 * tol
    if bn2:
        assert rel(dg2, g2r.grad) < 3 * tol and rel(db2, b2r.grad) < 3 * tol


@pytest.mark.parametrize("dtn,tdt,tol", DT)
@pytest.mark.parametrize("variant", ["relu_bn", "bn", "bn_relu_out", "frozen_relu", "relu_bn_acc"])
@pytest.mark.parametrize("shape", [(2, 4, 40, 64, 256), (3, 5, 61, 67, 72)], ids=["decoder", "ragged"])
def test_bn_backward_without_second_operand_large(A, dtn, tdt, tol, variant, shape):
    """y = relu?(norm(a)) on tensors above the cooperative-launch limit: the register-resident reduce / apply kernels
    (apply_bwd_reduce_nob / apply_bwd_nob: four fixed channels per thread, four positions in flight).  "ragged": channels not a
    multiple of 64, positions not a multiple of anything; "frozen": moving-statistics norm (no reductions in d a); "acc": the
    data gradient is accumulated onto an existing one."""
    torch.manual_seed(11)
    dev = "cuda"
    dt = A.BF16 if dtn == "bf16" else A.F32
    N, D, H, W, Cc = shape
    P = N * D * H * W
    relu1 = int(variant in ("relu_bn", "frozen_relu", "relu_bn_acc"))
    relu_out = int(variant == "bn_relu_out")
    training = variant != "frozen_relu"
    acc = int(variant == "relu_bn_acc")
    a = (torch.randn(N, D, H, W, Cc, device=dev) * 1.5 + 0.7).to(tdt)
    dy = torch.randn(N, D, H, W, Cc, device=dev).to(tdt)
    g1, b1 = torch.rand(Cc, device=dev) + 0.5, torch.randn(Cc, device=dev) * 0.1
    mm, mv = torch.randn(Cc, device=dev) * 0.3 + 0.6, torch.rand(Cc, device=dev) + 1.5
    mm0, mv0 = mm.clone(), mv.clone()
    f = lambda: torch.empty(Cc, device=dev)  # noqa: E731
    s1, t1, m1, r1 = f(), f(), f(), f()
    A.check(A.lib.sap3d_bn_finalize(A.ptr(_stats(a)), 1, Cc, float(P), A.ptr(g1), A.ptr(b1), A.ptr(mm), A.ptr(mv), int(training), 0.99,
                                    1e-3, A.ptr(s1), A.ptr(t1), A.ptr(m1), A.ptr(r1), stream()), "fin")
    af = a.float().requires_grad_(True)
    g1r, b1r = g1.clone().requires_grad_(True), b1.clone().requires_grad_(True)
    z = tfs.batch_norm(af, g1r, b1r, mm0, mv0, training)[0]
    ref = torch.relu(z) if (relu1 or relu_out) else z
    ref.backward(dy.float())
    base = (torch.randn(N, D, H, W, Cc, device=dev) * 0.5).to(tdt)
    da = base.clone() if acc else torch.empty_like(a)
    dg1, db1 = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev)
    ws = torch.zeros(A.lib.sap3d_affine_act_bwd_workspace(Cc) // 4 + 16, device=dev)
    A.check(A.lib.sap3d_affine_act_bwd(dt, A.ptr(dy), A.ptr(a), A.ptr(s1), A.ptr(t1), A.ptr(m1) if training else None,
                                       A.ptr(r1) if training else None, relu1, None, None, None, None, None, 0, relu_out, P, Cc,
                                       A.ptr(da), acc, None, 0, A.ptr(dg1), A.ptr(db1), None, None, A.ptr(ws), stream()), "bwd")
    torch.cuda.synchronize()
    tol = max(tol, 1e-3)            # millions of elements: single ReLU-mask flips at |z| ~ 1e-7
    want = af.grad + (base.float() if acc else 0)
    assert rel(da, want) < 3 * tol, rel(da, want)
    assert rel(db1, b1r.grad) < 3 * tol
    if training:        # (a frozen norm's gamma is not trained by any graph: its x-hat is not formed)
        assert rel(dg1, g1r.grad) < 3 * tol


@pytest.mark.parametrize("dtn,tdt,tol", DT)
@pytest.mark.parametrize("P", [2 * 37, 2 * 4 * 28 * 28], ids=["tiny", "stage1"])
def test_bn_backward_in_two_phases_over_two_replicas(A, dtn, tdt, tol, P):
    """sap3d_affine_act_bwd_sync (synchronised BatchNorm): two 'replicas' each own half of the positions; phase 1 leaves
    (local sums / global count) in each workspace, the caller adds them up, phase 2 applies.  The result equals the one-shot
    backward over all positions, and d(gamma), d(beta) add up to the global ones."""
    torch.manual_seed(3)
    dev, Cc = "cuda", 192
    dt = A.BF16 if dtn == "bf16" else A.F32
    a = (torch.randn(P, Cc, device=dev) * 1.5 + 0.7).to(tdt)
    b = (torch.randn(P, Cc, device=dev) * 0.8 - 0.2).to(tdt)
    dy = torch.randn(P, Cc, device=dev).to(tdt)
    gam = [torch.rand(Cc, device=dev) + 0.5 for _ in range(2)]
    bet = [torch.randn(Cc, device=dev) * 0.1 for _ in range(2)]
    mm, mv = torch.zeros(Cc, device=dev), torch.ones(Cc, device=dev)
    f = lambda: torch.empty(Cc, device=dev)  # noqa: E731
    sc, sh, me, rs = [f(), f()], [f(), f()], [f(), f()], [f(), f()]
    for i, x in enumerate((a, b)):   # global statistics (what the all-reduced forward rows finalise to)
        A.check(A.lib.sap3d_bn_finalize(A.ptr(_stats(x)), 1, Cc, float(P), A.ptr(gam[i]), A.ptr(bet[i]), A.ptr(mm), A.ptr(mv), 1, 0.99,
                                        1e-3, A.ptr(sc[i]), A.ptr(sh[i]), A.ptr(me[i]), A.ptr(rs[i]), stream()), "fin")
    nws = A.lib.sap3d_affine_act_bwd_workspace(Cc) // 4 + 16

    def args(lo, hi, da, db, dg, ws):
        return (dt, A.ptr(dy[lo:hi]), A.ptr(a[lo:hi]), A.ptr(sc[0]), A.ptr(sh[0]), A.ptr(me[0]), A.ptr(rs[0]), 1, A.ptr(b[lo:hi]),
                A.ptr(sc[1]), A.ptr(sh[1]), A.ptr(me[1]), A.ptr(rs[1]), 0, 1, hi - lo, Cc, A.ptr(da[lo:hi]), 0, A.ptr(db[lo:hi]), 0,
                A.ptr(dg[0]), A.ptr(dg[1]), A.ptr(dg[2]), A.ptr(dg[3]), A.ptr(ws), stream())

    da0, db0 = torch.empty_like(a), torch.empty_like(b)
    dg0 = [torch.zeros(Cc, device=dev) for _ in range(4)]
    A.check(A.lib.sap3d_affine_act_bwd(*args(0, P, da0, db0, dg0, torch.zeros(nws, device=dev))), "one shot")
    da1, db1 = torch.empty_like(a), torch.empty_like(b)
    halves = [(0, P // 2), (P // 2, P)]
    wss = [torch.zeros(nws, device=dev) for _ in halves]
    dgs = [[torch.zeros(Cc, device=dev) for _ in range(4)] for _ in halves]
    for (lo, hi), ws, dg in zip(halves, wss, dgs):
        A.check(A.lib.sap3d_affine_act_bwd_sync(*args(lo, hi, da1, db1, dg, ws), float(P), 1), "phase 1")
    total = wss[0][:4 * Cc] + wss[1][:4 * Cc]                    # the all-reduce
    for (lo, hi), ws, dg in zip(halves, wss, dgs):
        ws[:4 * Cc].copy_(total)
        A.check(A.lib.sap3d_affine_act_bwd_sync(*args(lo, hi, da1, db1, dg, ws), float(P), 2), "phase 2")
    torch.cuda.synchronize()
    assert rel(da1, da0) < 10 * tol * (1e-2 if dtn == "bf16" else 1) + 1e-5, rel(da1, da0)
    assert rel(db1, db0) < 10 * tol * (1e-2 if dtn == "bf16" else 1) + 1e-5, rel(db1, db0)
    for k in range(4):
        assert rel(dgs[0][k] + dgs[1][k], dg0[k]) < 1e-4, k
    # argument checking
    assert A.lib.sap3d_affine_act_bwd_sync(*args(0, P, da1, db1, dg0, wss[0]), float(P), 0) != 0
    assert A.lib.sap3d_affine_act_bwd_sync(*args(0, P, da1, db1, dg0, wss[0]), float(P - 1), 1) != 0


@pytest.mark.parametrize("dtn,tdt,tol", DT)
@pytest.mark.parametrize("two_norms,training", [(False, True), (True, True), (True, False)])
def test_bn_apply_fused_matches_finalize_plus_apply(A, dtn, tdt, tol, two_norms, training):
    """the one-launch form (finalize folded into apply) is bit-identical to sap3d_bn_finalize + sap3d_affine_act"""
    torch.manual_seed(5)
    dev = "cuda"
    dt = A.BF16 if dtn == "bf16" else A.F32
    P, Cc, rows = 7 * 11 * 3, 192, 5
    a = (torch.randn(P, Cc, device=dev) * 1.5 + 0.7).to(tdt)
    b = (torch.randn(P, Cc, device=dev) * 0.8 - 0.2).to(tdt)

    def split_stats(x):   # per-"tile" partial sums as a conv epilogue would write them
        xf = x.float()
        parts = torch.chunk(xf, rows, 0)
        return torch.stack([torch.stack([q.sum(0), (q * q).sum(0)], 0) for q in parts], 0).contiguous()

    sa, sb = split_stats(a), split_stats(b)
    gam = [torch.rand(Cc, device=dev) + 0.5 for _ in range(2)]
    bet = [torch.randn(Cc, device=dev) * 0.1 for _ in range(2)]
    outs = []
    for fused in (False, True):
        mm = [torch.randn(Cc, device=dev) * 0.1 for _ in range(2)]
        mv = [torch.rand(Cc, device=dev) + 0.5 for _ in range(2)]
        torch.manual_seed(6)
        mm = [torch.full((Cc,), 0.05, device=dev), torch.full((Cc,), -0.02, device=dev)]
        mv = [torch.full((Cc,), 0.9, device=dev), torch.full((Cc,), 1.1, device=dev)]
        f = lambda: torch.zeros(Cc, device=dev)  # noqa: E731
        sc, sh, me, rs = [f(), f()], [f(), f()], [f(), f()], [f(), f()]
        y = torch.empty_like(a)
        tr = int(training)
        if fused:
            A.check(A.lib.sap3d_bn_apply_fused(dt, A.ptr(a), A.ptr(sa), rows, A.ptr(gam[0]), A.ptr(bet[0]), A.ptr(mm[0]), A.ptr(mv[0]), tr,
                                               A.ptr(sc[0]), A.ptr(sh[0]), A.ptr(me[0]), A.ptr(rs[0]), 1, A.ptr(b), int(two_norms),
                                               A.ptr(sb) if two_norms else None, rows, A.ptr(gam[1]), A.ptr(bet[1]), A.ptr(mm[1]), A.ptr(mv[1]),
                                               tr, A.ptr(sc[1]), A.ptr(sh[1]), A.ptr(me[1]), A.ptr(rs[1]), int(two_norms), 1, A.ptr(y), P, Cc,
                                               float(P), 0.99, 1e-3, stream()), "fused")
        else:
            A.check(A.lib.sap3d_bn_finalize(A.ptr(sa), rows, Cc, float(P), A.ptr(gam[0]), A.ptr(bet[0]), A.ptr(mm[0]), A.ptr(mv[0]), tr, 0.99,
                                            1e-3, A.ptr(sc[0]), A.ptr(sh[0]), A.ptr(me[0]), A.ptr(rs[0]), stream()), "fin1")
            if two_norms:
                A.check(A.lib.sap3d_bn_finalize(A.ptr(sb), rows, Cc, float(P), A.ptr(gam[1]), A.ptr(bet[1]), A.ptr(mm[1]), A.ptr(mv[1]), tr,
                                                0.99, 1e-3, A.ptr(sc[1]), A.ptr(sh[1]), A.ptr(me[1]), A.ptr(rs[1]), stream()), "fin2")
            A.check(A.lib.sap3d_affine_act(dt, A.ptr(a), A.ptr(sc[0]), A.ptr(sh[0]), 1, A.ptr(b), A.ptr(sc[1]) if two_norms else None,
                                           A.ptr(sh[1]) if two_norms else None, int(two_norms), 1, A.ptr(y), P, Cc, 0, stream()), "apply")
        torch.cuda.synchronize()
        outs.append((y, sc[0], sh[0], me[0], rs[0], mm[0], mv[0], sc[1], mm[1]))
    for u, v in zip(*outs):
        assert rel(u, v) < 1e-6


@pytest.mark.parametrize("dtn,tdt,tol", DT)
@pytest.mark.parametrize("k,s,same", [((2, 1, 1), (2, 1, 1), 1), ((2, 3, 3), (2, 2, 2), 1), ((2, 2, 2), (2, 2, 2), 0)])
def test_maxpool(A, dtn, tdt, tol, k, s, same):
    torch.manual_seed(1)
    dt = A.BF16 if dtn == "bf16" else A.F32
    N, D, H, W, Cc = 2, 4, 9, 10, 16
    x = torch.randn(N, D, H, W, Cc, device="cuda").to(tdt)
    xr = x.float().requires_grad_(True)
    ref = tfs.max_pool3d_same(xr, k, s) if same else tfs.max_pool3d_valid(xr, 2)
    y = torch.empty(ref.shape, device="cuda", dtype=tdt)
    amax = torch.empty(ref.shape, device="cuda", dtype=torch.uint8)
    A.check(A.lib.sap3d_maxpool3d_fwd(dt, A.ptr(x), N, D, H, W, Cc, A.i3(k), A.i3(s), same, A.ptr(y), A.ptr(amax), stream()), "pool")
    assert rel(y, ref) == 0.0
    dy = torch.randn_like(ref).to(tdt)
    ref.backward(dy.float())
    for idx in (None, amax):   # window re-scan and saved-arg-max forms of the backward pass
        dx = torch.empty_like(x)
        A.check(A.lib.sap3d_maxpool3d_bwd(dt, A.ptr(x), A.ptr(dy), N, D, H, W, Cc, A.i3(k), A.i3(s), same, A.ptr(idx), A.ptr(dx), 0,
                                          stream()), "poolb")
        torch.cuda.synchronize()
        assert rel(dx, xr.grad) < tol


def test_maxpool_ties_go_to_the_first_maximum(A):
    """post-ReLU activations are full of exact ties (zeros): both backward forms route a window's gradient to the first
    maximal element in (d, h, w) scan order, as torch's max_pool3d backward does"""
    N, D, H, W, Cc = 1, 4, 8, 8, 8
    x = torch.relu(torch.randn(N, D, H, W, Cc, device="cuda") - 0.8).to(torch.bfloat16)
    k, s = (2, 3, 3), (2, 2, 2)
    xr = x.float().requires_grad_(True)
    ref = tfs.max_pool3d_same(xr, k, s)
    y = torch.empty(ref.shape, device="cuda", dtype=torch.bfloat16)
    amax = torch.empty(ref.shape, device="cuda", dtype=torch.uint8)
    A.check(A.lib.sap3d_maxpool3d_fwd(A.BF16, A.ptr(x), N, D, H, W, Cc, A.i3(k), A.i3(s), 1, A.ptr(y), A.ptr(amax), stream()), "pool")
    dy = torch.randn_like(ref).to(torch.bfloat16)
    outs = []
    for idx in (None, amax):
        dx = torch.empty_like(x)
        A.check(A.lib.sap3d_maxpool3d_bwd(A.BF16, A.ptr(x), A.ptr(dy), N, D, H, W, Cc, A.i3(k), A.i3(s), 1, A.ptr(idx), A.ptr(dx), 0,
                                          stream()), "poolb")
        outs.append(dx)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    assert rel(outs[0].float().sum(), dy.float().sum()) < 1e-2   # every window's gradient lands exactly once


@pytest.mark.parametrize("shape", [(2, 3, 5, 6, 128), (1, 4, 7, 9, 64), (2, 2, 8, 8, 256)])
def test_head_tensor_core_form(A, shape):
    """k3 s2 head as GEMM + col2im (forward) and GEMMs over im2col(dlogits) (backward) vs the oracle's conv3d_transpose
    (p3d.py:393) on the same bf16 inputs; weights and dlogits are rounded to bf16 on this path (tolerance 1e-2)"""
    torch.manual_seed(4)
    N, D, H, W, Cc = shape
    x = torch.randn(N, D, H, W, Cc, device="cuda").bfloat16()
    w = torch.randn(3, 3, 3, 1, Cc, device="cuda") * 0.05
    b = torch.randn(1, device="cuda")
    logits = torch.empty(N, 2 * D, 2 * H, 2 * W, 1, device="cuda")
    pred = torch.empty_like(logits)
    ws = torch.empty(A.lib.sap3d_head_tc_workspace(N, D, H, W, Cc) // 4 + 16, device="cuda")
    A.check(A.lib.sap3d_head_tc_fwd(A.ptr(x), N, D, H, W, Cc, A.ptr(w), A.ptr(b), A.ptr(logits), A.ptr(pred), A.ptr(ws), stream()), "head tc")
    xr, wr = x.float().requires_grad_(True), w.clone().requires_grad_(True)
    lref = tfs.conv3d_transpose_same(xr, wr, (2, 2, 2), b)
    assert rel(logits, lref) < 1e-2
    assert rel(pred, torch.sigmoid(lref)) < 1e-2
    dl = torch.randn_like(lref)
    lref.backward(dl)
    dx = torch.ones_like(x)
    dw = torch.zeros_like(w)
    A.check(A.lib.sap3d_head_tc_bwd(A.ptr(dl), A.ptr(x), N, D, H, W, Cc, A.ptr(dx), 1, A.ptr(dw), A.ptr(ws), stream(), None), "head tc bwd")
    torch.cuda.synchronize()
    assert rel(dx.float() - 1.0, xr.grad) < 2e-2     # accumulated onto ones in bf16
    assert rel(dw, wr.grad) < 1e-2


@pytest.mark.parametrize("dtn,tdt,tol", DT)
def test_head_loss(A, dtn, tdt, tol):
    torch.manual_seed(2)
    dt = A.BF16 if dtn == "bf16" else A.F32
    N, D, H, W, Cc = 2, 3, 5, 6, 128
    x = torch.randn(N, D, H, W, Cc, device="cuda").to(tdt)
    w = torch.randn(3, 3, 3, 1, Cc, device="cuda") * 0.05
    b = torch.randn(1, device="cuda")
    tgt = torch.rand(N, 2 * D, 2 * H, 2 * W, device="cuda")
    logits = torch.empty(N, 2 * D, 2 * H, 2 * W, 1, device="cuda")
    pred = torch.empty_like(logits)
    A.check(A.lib.sap3d_head_fwd(dt, A.ptr(x), N, D, H, W, Cc, A.i3((3, 3, 3)), 2, A.ptr(w), A.ptr(b), A.ptr(logits), A.ptr(pred),
                                 stream()), "head")
    xr, wr, br = x.float().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    lref = tfs.conv3d_transpose_same(xr, wr, (2, 2, 2), br)
    pref = torch.sigmoid(lref)
    assert rel(logits, lref) < tol and rel(pred, pref) < tol
    loss_ref = tfs.smooth_l1_loss(pref.reshape(tgt.shape), tgt)
    loss_ref.backward()
    dlog = torch.empty_like(logits)
    loss = torch.zeros(1, device="cuda", dtype=torch.float64)
    dbias = torch.zeros(1, device="cuda")
    A.check(A.lib.sap3d_loss_smooth_l1(A.ptr(logits), A.ptr(tgt), logits.numel(), 1, None, A.ptr(dlog), A.ptr(loss), A.ptr(dbias),
                                       stream()), "loss")
    dx = torch.empty_like(x)
    dw = torch.zeros_like(w)
    A.check(A.lib.sap3d_head_bwd(dt, A.ptr(dlog), A.ptr(x), N, D, H, W, Cc, A.i3((3, 3, 3)), 2, A.ptr(w), A.ptr(dx), 0, A.ptr(dw),
                                 stream()), "headb")
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) / loss_ref.item() < 1e-4 if dtn == "f32" else 5e-3
    assert rel(dx, xr.grad) < 3 * tol
    assert rel(dw, wr.grad) < 3 * tol
    assert rel(dbias, br.grad) < 3 * tol


def test_adam_tf_formula(A):
    torch.manual_seed(3)
    n = 1000
    w, g = torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step = torch.zeros(1, device="cuda", dtype=torch.int32)
    wr, mr, vr = w.clone(), m.clone(), v.clone()
    for t in (1, 2, 3):
        A.check(A.lib.sap3d_step_increment(A.ptr(step), stream()), "inc")
        A.check(A.lib.sap3d_adam_step(A.ptr(w), A.ptr(g), A.ptr(m), A.ptr(v), n, A.ptr(step), 1e-4, 0.9, 0.999, 1e-8, 1.0, stream()), "adam")
        wr, mr, vr = tfs.adam_step_tf(wr, g, mr, vr, t)
    torch.cuda.synchronize()
    # the kernel evaluates (1 - beta2) in fp32 like TF's ApplyAdam (0.0010000467 instead of 0.001): 5e-5 on v
    assert rel(w, wr) < 1e-6 and rel(m, mr) < 1e-6 and rel(v, vr) < 1e-4


def test_dropout_matches_numpy_hash(A):
    import numpy as np

    n = 4096
    x = torch.ones(n, device="cuda")
    y = torch.empty_like(x)
    step = torch.full((1,), 5, device="cuda", dtype=torch.int32)
    A.check(A.lib.sap3d_dropout(A.F32, A.ptr(x), A.ptr(y), n, 0.5, 1234, A.ptr(step), 0, stream()), "dropout")
    torch.cuda.synchronize()
    from oracle.dropout_hash import keep_mask

    ref = keep_mask(1234 + 5, n, 0.5).astype(np.float32) * 2.0
    assert np.array_equal(y.cpu().numpy(), ref)


@pytest.mark.parametrize("dtn,tdt,tol", DT)
def test_attention_gate(A, dtn, tdt, tol):
    """x = o * gamma + x (utils/network.py:191-192) and its gradients (d_o, dx accumulate, dgamma)"""
    torch.manual_seed(5)
    dt = A.BF16 if dtn == "bf16" else A.F32
    n = 2 * 3 * 4 * 5 * 64
    o, x, dy = [torch.randn(n, device="cuda").to(tdt) for _ in range(3)]
    gamma = torch.tensor([0.37], device="cuda")
    y = torch.empty_like(o)
    A.check(A.lib.sap3d_gate_fwd(dt, A.ptr(o), A.ptr(x), A.ptr(gamma), A.ptr(y), n, stream()), "gate")
    assert rel(y, o.float() * 0.37 + x.float()) < tol
    d_o, dx, dg = torch.empty_like(o), torch.ones_like(o), torch.zeros(1, device="cuda")
    A.check(A.lib.sap3d_gate_bwd(dt, A.ptr(dy), A.ptr(o), A.ptr(gamma), A.ptr(d_o), A.ptr(dx), 1, A.ptr(dg), n, stream()), "gateb")
    torch.cuda.synchronize()
    assert rel(d_o, dy.float() * 0.37) < tol and rel(dx, dy.float() + 1) < tol
    assert rel(dg, (dy.float() * o.float()).sum().reshape(1)) < 1e-3


@pytest.mark.parametrize("shape", [(2, 49, 49, 128, 1024), (1, 392, 392, 64, 512), (2, 256, 64, 16, 128)])
@pytest.mark.parametrize("dtn,tdt,tol", DT)
def test_attention_core(A, dtn, tdt, tol, shape):
    """o = softmax(g f^T) h and its gradients: CUDA-core kernels (fp32 / tiny sites) against torch autograd"""
    torch.manual_seed(6)
    dt = A.BF16 if dtn == "bf16" else A.F32
    B, Nq, Nk, dk, dv = shape
    g = (torch.randn(B, Nq, dk, device="cuda") * 0.3).to(tdt)
    f = (torch.randn(B, Nk, dk, device="cuda") * 0.3).to(tdt)
    h = torch.randn(B, Nk, dv, device="cuda").to(tdt)
    d_o = torch.randn(B, Nq, dv, device="cuda").to(tdt)
    beta = torch.empty(B, Nq, Nk, device="cuda", dtype=tdt)
    o = torch.empty(B, Nq, dv, device="cuda", dtype=tdt)
    A.check(A.lib.sap3d_attention_fwd(dt, A.ptr(g), A.ptr(f), A.ptr(h), A.ptr(beta), A.ptr(o), B, Nq, Nk, dk, dv, dk, dk, dv, Nk, dv,
                                      stream()), "attn")
    gr, fr, hr = [t.float().requires_grad_(True) for t in (g, f, h)]
    ref = torch.softmax(gr @ fr.transpose(1, 2), dim=-1) @ hr
    assert rel(o, ref) < tol
    ref.backward(d_o.float())
    ds = torch.empty_like(beta)
    dg, df, dh = torch.empty_like(g), torch.empty_like(f), torch.empty_like(h)
    A.check(A.lib.sap3d_attention_bwd(dt, A.ptr(g), A.ptr(f), A.ptr(h), A.ptr(beta), A.ptr(d_o), A.ptr(ds), A.ptr(dg), A.ptr(df), A.ptr(dh),
                                      B, Nq, Nk, dk, dv, dk, dk, dv, Nk, dv, stream()), "attnb")
    torch.cuda.synchronize()
    assert rel(dg, gr.grad) < 4 * tol and rel(df, fr.grad) < 4 * tol and rel(dh, hr.grad) < 4 * tol


@pytest.mark.parametrize("dtn,tdt,tol", DT)
@pytest.mark.parametrize("case", [("stage 3 tail: bn(a) + x, relu", 4, 98, 1024, False, True, False, False, True),
                                  ("ST_B: relu(bn(t)) + relu(bn(s))", 3, 784, 128, True, True, True, True, False),
                                  ("plain bn + relu, ragged C", 5, 37, 72, True, False, False, False, False),
                                  ("ST_C: s + relu(bn(t))", 2, 1000, 64, True, True, False, False, False)],
                         ids=lambda c: c[0] if isinstance(c, tuple) else None)
def test_per_clip_batchnorm_in_one_launch(A, dtn, tdt, tol, case):
    """sap3d_sample_norm_apply (one launch) == sap3d_sample_channel_partials + sap3d_gn_finalize(G = C) + sap3d_affine_act, and both
    == the per-clip batch-statistics BatchNorm written out in torch"""
    _, N, S, Cc, relu1, with_b, norm2, relu2, relu_out = case
    dev = "cuda"
    dt = A.BF16 if dtn == "bf16" else A.F32
    torch.manual_seed(N * 1000 + S + Cc)
    a = (torch.randn(N, S, Cc, device=dev) * 1.5 + 0.3).to(tdt)
    b = (torch.randn(N, S, Cc, device=dev) * 0.7 - 0.2).to(tdt) if with_b else None
    g1, b1 = torch.rand(Cc, device=dev) + 0.5, torch.randn(Cc, device=dev) * 0.3
    g2, b2 = (torch.rand(Cc, device=dev) + 0.5, torch.randn(Cc, device=dev) * 0.3) if norm2 else (None, None)
    eps = 1e-3
    if A.lib.sap3d_sample_norm_apply_supported(dt, S, Cc) != 1:
        assert S * 64 * (2 if dtn == "bf16" else 4) > (128 << 10)       # only slabs over the 128 KB limit are refused
        pytest.skip("slab larger than the one-launch form takes")

    def clip_bn(x, g, bt):
        xf = x.float()
        m = xf.mean(dim=1, keepdim=True)
        v = xf.var(dim=1, unbiased=False, keepdim=True)
        return (xf - m) * torch.rsqrt(v + eps) * g + bt

    ref = clip_bn(a, g1, b1)
    if relu1:
        ref = torch.relu(ref)
    if with_b:
        q = clip_bn(b, g2, b2) if norm2 else b.float()
        ref = ref + (torch.relu(q) if relu2 else q)
    if relu_out:
        ref = torch.relu(ref)
    y = torch.full_like(a, float("nan"))
    A.check(A.lib.sap3d_sample_norm_apply(dt, A.ptr(a), A.ptr(g1), A.ptr(b1), int(relu1), A.ptr(b), A.ptr(g2), A.ptr(b2), int(relu2), int(relu_out),
                                          A.ptr(y), N, S, Cc, eps, stream()), "sample_norm_apply")
    torch.cuda.synchronize()
    assert not torch.isnan(y.float()).any()
    assert rel(y, ref) < 3 * tol, rel(y, ref)
    if Cc % 8 == 0:     # the three-launch path it replaces (needs C % 8 == 0)
        rows = A.lib.sap3d_sample_stats_rows(S, Cc, N)
        f = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)  # noqa: E731
        coef = []
        for x, g, bt in ((a, g1, b1),) + (((b, g2, b2),) if norm2 else ()):
            part, sc, sh, mean, rstd = f(N, rows, 3, Cc), f(N, Cc), f(N, Cc), f(N, Cc), f(N, Cc)
            A.check(A.lib.sap3d_sample_channel_partials(dt, A.ptr(x), None, N, S, Cc, rows, A.ptr(part), stream()), "partials")
            A.check(A.lib.sap3d_gn_finalize(A.ptr(part), rows, N, S, Cc, Cc, A.ptr(g), A.ptr(bt), eps, A.ptr(sc), A.ptr(sh), A.ptr(mean), A.ptr(rstd),
                                            stream()), "finalize")
            coef += [sc, sh]
        y3 = torch.full_like(a, float("nan"))
        A.check(A.lib.sap3d_affine_act(dt, A.ptr(a), A.ptr(coef[0]), A.ptr(coef[1]), int(relu1), A.ptr(b), A.ptr(coef[2]) if norm2 else None,
                                       A.ptr(coef[3]) if norm2 else None, int(relu2), int(relu_out), A.ptr(y3), N * S, Cc, S, stream()), "affine_act")
        torch.cuda.synchronize()
        assert rel(y, y3) < (2e-3 if dtn == "bf16" else 1e-5), rel(y, y3)
