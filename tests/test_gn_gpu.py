"""GroupNorm + CBAM variant (gn/p3d_gn.py, utils/network.py:65-87,198-274): op-level forward/backward parity of the
CUDA kernels (through the C ABI) against torch autograd on the oracle primitives, and whole-graph parity of
inference_p3d / inference_p3d_concat (forward, and one training step of gn/train_p3d_gn_dataset.py:169-199).

Tolerances: fp32 storage 1e-4 relative (Frobenius); bf16 storage 2e-2 (inputs are the same bf16-rounded values)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import p3d_oracle as O  # noqa: E402
from oracle import tf_semantics as tfs  # noqa: E402


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


DT = [("f32", torch.float32, 1e-4), ("bf16", torch.bfloat16, 2e-2)]


def gn_forward(A, dt, x, gamma, beta, G):
    """per-sample statistics of x through the ABI: returns scale, shift [N][C], mean, rstd [N][G]"""
    N, Cc = x.shape[0], x.shape[-1]
    S = x.numel() // (N * Cc)
    rows = A.lib.sap3d_sample_stats_rows(S, Cc, N)
    part = torch.empty(N, rows, 3, Cc, device="cuda")
    sc, sh = torch.empty(N, Cc, device="cuda"), torch.empty(N, Cc, device="cuda")
    mean, rstd = torch.empty(N, G, device="cuda"), torch.empty(N, G, device="cuda")
    A.check(A.lib.sap3d_sample_channel_partials(dt, A.ptr(x), None, N, S, Cc, rows, A.ptr(part), stream()), "partials")
    A.check(A.lib.sap3d_gn_finalize(A.ptr(part), rows, N, S, Cc, G, A.ptr(gamma), A.ptr(beta), 1e-5, A.ptr(sc), A.ptr(sh), A.ptr(mean),
                                    A.ptr(rstd), stream()), "gn_finalize")
    return sc, sh, mean, rstd


@pytest.mark.parametrize("dtn,tdt,tol", DT)
@pytest.mark.parametrize("Cc", [16, 64, 256])
@pytest.mark.parametrize("pattern", ["gn_relu", "gn_only", "relu_gn_plus_relu_gn", "relu_gn_plus_t"])
def test_gn_act_fwd_bwd(A, dtn, tdt, tol, Cc, pattern):
    torch.manual_seed(1)
    dev = "cuda"
    dt = A.BF16 if dtn == "bf16" else A.F32
    N, D, H, W = 3, 2, 5, 7
    S, G = D * H * W, min(32, Cc)
    a = (torch.randn(N, D, H, W, Cc, device=dev) * 1.5 + 0.7).to(tdt)
    b = (torch.randn(N, D, H, W, Cc, device=dev) * 0.8 - 0.2).to(tdt)
    dy = torch.randn(N, D, H, W, Cc, device=dev).to(tdt)
    g1, b1 = torch.rand(Cc, device=dev) + 0.5, torch.randn(Cc, device=dev) * 0.1
    g2, b2 = torch.rand(Cc, device=dev) + 0.5, torch.randn(Cc, device=dev) * 0.1
    relu1, relu2, relu_out, use_b, gn2 = {
        "gn_relu": (1, 0, 0, False, False),
        "gn_only": (0, 0, 0, False, False),                # GroupNorm(dw3d) without ReLU
        "relu_gn_plus_relu_gn": (1, 1, 0, True, True),     # ST_B
        "relu_gn_plus_t": (1, 0, 0, True, False),          # ST_C
    }[pattern]
    s1, t1, m1, r1 = gn_forward(A, dt, a, g1, b1, G)
    s2 = t2 = m2 = r2 = None
    if gn2:
        s2, t2, m2, r2 = gn_forward(A, dt, b, g2, b2, G)
    y = torch.empty_like(a)
    A.check(A.lib.sap3d_affine_act(dt, A.ptr(a), A.ptr(s1), A.ptr(t1), relu1, A.ptr(b) if use_b else None, A.ptr(s2), A.ptr(t2), relu2,
                                   relu_out, A.ptr(y), N * S, Cc, S, stream()), "apply")
    af, bf = a.float().requires_grad_(True), b.float().requires_grad_(True)
    g1r, b1r, g2r, b2r = [t.clone().requires_grad_(True) for t in (g1, b1, g2, b2)]
    z = tfs.group_norm(af, g1r, b1r)
    z = torch.relu(z) if relu1 else z
    if use_b:
        z2 = tfs.group_norm(bf, g2r, b2r) if gn2 else bf
        z = z + (torch.relu(z2) if relu2 else z2)
    ref = torch.relu(z) if relu_out else z
    assert rel(y, ref) < tol
    ref.backward(dy.float())
    da, db = torch.empty_like(a), torch.empty_like(b)
    dg1, db1, dg2, db2 = [torch.zeros(Cc, device=dev) for _ in range(4)]
    ws = torch.zeros(A.lib.sap3d_gn_bwd_workspace(N, S, Cc) // 4 + 16, device=dev)
    A.check(A.lib.sap3d_gn_act_bwd(dt, A.ptr(dy), A.ptr(a), A.ptr(s1), A.ptr(t1), A.ptr(m1), A.ptr(r1), A.ptr(g1), relu1,
                                   A.ptr(b) if use_b else None, A.ptr(s2), A.ptr(t2), A.ptr(m2), A.ptr(r2), A.ptr(g2) if gn2 else None,
                                   relu2, relu_out, N, S, Cc, G, A.ptr(da), 0, A.ptr(db) if use_b else None, 0, A.ptr(dg1), A.ptr(db1),
                                   A.ptr(dg2) if gn2 else None, A.ptr(db2) if gn2 else None, A.ptr(ws), stream()), "gn_act_bwd")
    torch.cuda.synchronize()
    assert rel(da, af.grad) < tol
    assert rel(dg1, g1r.grad) < tol and rel(db1, b1r.grad) < tol
    if use_b:
        assert rel(db, bf.grad) < tol
    if gn2:
        assert rel(dg2, g2r.grad) < tol and rel(db2, b2r.grad) < tol
    # accumulation into existing gradients
    da2 = torch.ones_like(a)
    A.check(A.lib.sap3d_gn_act_bwd(dt, A.ptr(dy), A.ptr(a), A.ptr(s1), A.ptr(t1), A.ptr(m1), A.ptr(r1), A.ptr(g1), relu1,
                                   A.ptr(b) if use_b else None, A.ptr(s2), A.ptr(t2), A.ptr(m2), A.ptr(r2), A.ptr(g2) if gn2 else None,
                                   relu2, relu_out, N, S, Cc, G, A.ptr(da2), 1, None, 0, A.ptr(dg1), A.ptr(db1), None, None, A.ptr(ws),
                                   stream()), "gn_act_bwd acc")
    assert rel(da2.float() - 1.0, af.grad) < 2 * tol + (2e-2 if dtn == "bf16" else 0)


def _cbam_ref(r, w0, b0, w1, b1, wsp):
    n, c = r.shape[0], r.shape[-1]
    avg, mx = r.mean(dim=(1, 2, 3)), r.amax(dim=(1, 2, 3))
    mlp = lambda v: torch.relu(v @ w0 + b0) @ w1 + b1  # noqa: E731
    u = r * torch.sigmoid(mlp(avg) + mlp(mx)).view(n, 1, 1, 1, c)
    cat = torch.cat([u.mean(dim=4, keepdim=True), u.amax(dim=4, keepdim=True)], dim=4)
    return u * torch.sigmoid(tfs.conv3d_same(cat, wsp, (1, 1, 1)))


@pytest.mark.parametrize("dtn,tdt,tol", DT)
@pytest.mark.parametrize("shape", [(2, 4, 9, 9, 256), (3, 2, 5, 5, 512), (2, 8, 6, 6, 64)])
def test_cbam_tail_fwd_bwd(A, dtn, tdt, tol, shape):
    """y = relu(GN(c3) + cbam_block(r)) (gn/p3d_gn.py:175-177) and all of its gradients"""
    torch.manual_seed(2)
    dev = "cuda"
    dt = A.BF16 if dtn == "bf16" else A.F32
    N, D, H, W, Cc = shape
    S, G, hid = D * H * W, min(32, Cc), Cc // 8
    c3 = (torch.randn(*shape, device=dev) * 1.2 + 0.3).to(tdt)
    r = torch.relu(torch.randn(*shape, device=dev) + 0.3).to(tdt)          # block inputs are post-ReLU
    dy = torch.randn(*shape, device=dev).to(tdt)
    g3, be3 = torch.rand(Cc, device=dev) * 0.5 + 0.2, torch.randn(Cc, device=dev) * 0.1
    w0, b0 = torch.randn(Cc, hid, device=dev) * (2.0 / Cc) ** 0.5, torch.randn(hid, device=dev) * 0.1
    w1, b1 = torch.randn(hid, Cc, device=dev) * (2.0 / hid) ** 0.5, torch.randn(Cc, device=dev) * 0.1
    wsp = torch.randn(7, 7, 7, 2, 1, device=dev) * 0.08
    s3, t3, m3, r3 = gn_forward(A, dt, c3, g3, be3, G)
    rows = A.lib.sap3d_sample_stats_rows(S, Cc, N)
    part = torch.empty(N, rows, 3, Cc, device=dev)
    cs, sp, att = torch.empty(N, Cc, device=dev), torch.empty(N, S, 2, device=dev), torch.empty(N, S, device=dev)
    save = torch.empty(N, 2 * Cc + 2 * hid, device=dev)
    A.check(A.lib.sap3d_cbam_fwd(dt, A.ptr(r), N, D, H, W, Cc, hid, A.ptr(w0), A.ptr(b0), A.ptr(w1), A.ptr(b1), A.ptr(wsp), A.ptr(part),
                                 rows, A.ptr(cs), A.ptr(sp), A.ptr(att), A.ptr(save), stream()), "cbam_fwd")
    y = torch.empty_like(r)
    A.check(A.lib.sap3d_cbam_merge(dt, A.ptr(c3), A.ptr(s3), A.ptr(t3), A.ptr(r), A.ptr(cs), A.ptr(att), A.ptr(y), N, S, Cc, stream()), "merge")
    leaves = [t.float().clone().requires_grad_(True) for t in (c3, r, g3, be3, w0, b0, w1, b1, wsp)]
    c3f, rf, g3f, be3f, w0f, b0f, w1f, b1f, wspf = leaves
    ref = torch.relu(tfs.group_norm(c3f, g3f, be3f) + _cbam_ref(rf, w0f, b0f, w1f, b1f, wspf))
    assert rel(y, ref) < tol
    ref.backward(dy.float())
    dc3, dr = torch.empty_like(c3), torch.empty_like(r)
    grads = [torch.zeros_like(t) for t in (g3, be3, w0, b0, w1, b1, wsp)]
    ws = torch.zeros(A.lib.sap3d_gn_bwd_workspace(N, S, Cc) // 4 + 16, device=dev)
    A.check(A.lib.sap3d_cbam_tail_bwd(dt, A.ptr(dy), A.ptr(y), A.ptr(c3), A.ptr(s3), A.ptr(m3), A.ptr(r3), A.ptr(g3), A.ptr(r), N, D, H, W,
                                      Cc, G, hid, A.ptr(w0), A.ptr(w1), A.ptr(wsp), A.ptr(cs), A.ptr(sp), A.ptr(att), A.ptr(save),
                                      A.ptr(dc3), 0, A.ptr(dr), 0, *[A.ptr(g) for g in grads], A.ptr(ws), stream()), "cbam_tail_bwd")
    torch.cuda.synchronize()
    # in bf16 the stored y decides the ReLU mask; elements within rounding of zero can differ from the fp32 oracle mask
    btol = tol if dtn == "f32" else 3e-2
    assert rel(dc3, c3f.grad) < btol
    assert rel(dr, rf.grad) < btol
    for name, g, leaf in zip(("gamma3", "beta3", "w0", "b0", "w1", "b1", "w_sp"), grads, leaves[2:]):
        assert rel(g, leaf.grad) < btol, name


@pytest.mark.parametrize("dtn,tdt,tol", DT)
def test_concat_split(A, dtn, tdt, tol):
    dt = A.BF16 if dtn == "bf16" else A.F32
    P, ca, cb = 77, 64, 24
    a, b = torch.randn(P, ca, device="cuda").to(tdt), torch.randn(P, cb, device="cuda").to(tdt)
    y = torch.empty(P, ca + cb, device="cuda", dtype=tdt)
    A.check(A.lib.sap3d_concat_channels(dt, A.ptr(a), A.ptr(b), A.ptr(y), P, ca, cb, stream()), "concat")
    assert torch.equal(y, torch.cat([a, b], -1))
    da, db = torch.zeros_like(a), torch.ones_like(b)
    A.check(A.lib.sap3d_split_channels(dt, A.ptr(y), A.ptr(da), 0, A.ptr(db), 1, P, ca, cb, stream()), "split")
    assert torch.equal(da, a) and rel(db.float(), b.float() + 1) < 1e-2


# ---- whole graphs --------------------------------------------------------------------------------
def build(graph, dtype, training, batch, size, dropout=0.0):
    import sap3d_tensorflow_b200 as sp
    from sap3d_tensorflow_b200.gn import p3d_gn

    xin = sp.placeholder([batch, 16, size, size, 3], dtype=dtype, training_graph=training)
    head = getattr(p3d_gn, graph)(xin, dropout, batch, training)
    return sp.Session(head)


@pytest.mark.parametrize("graph", ["inference_p3d", "inference_p3d_concat", "inference_p3d_decoder_block"])
def test_gn_forward_parity(lib_built, graph):
    batch, size = 2, 64
    x = O.synthetic_clip(batch, 16, size, seed=0)
    vs = O.VarStore(seed=0)
    taps = {}
    with torch.no_grad():
        ref = O.forward(graph, x, vs, False, taps=taps)
    # The GN builders return LOGITS (gn/p3d_gn.py:257-258).  fp32 storage: 1e-4 on logits, saliency map and every tap.
    # bf16 storage: in this random-weight network CBAM scales the residual path by ~0.25 per block, so the 47-block chain
    # amplifies ANY storage rounding (stage-2 output already differs by 1e-1 from the fp32 oracle, printed below); the
    # end-to-end bf16 numbers are therefore only bounded (logits 2e-1, saliency map 5e-2; 1.0e-2 / 1.0e-2 / 3.6e-2
    # measured) while bf16 parity at 1e-2 is asserted op by op on equal inputs (test_gn_act_fwd_bwd, test_cbam_tail_fwd_bwd,
    # test_conv_gpu.py).
    for dtype, tol in (("f32", 1e-4), ("bf16", 2e-1)):
        sess = build(graph, dtype, False, batch, size)
        assert set(sess.eng.params) == set(vs.params)
        sess.eng.load_params(vs.params)
        pred = sess.run(x.cuda())
        torch.cuda.synchronize()
        sig = rel(torch.sigmoid(pred.float().cpu()), torch.sigmoid(ref))
        print(f"[{graph}/{dtype}] logits rel {rel(pred, ref):.3e}  sigmoid rel {sig:.3e}  " +
              " ".join(f"{k}={rel(sess.eng.taps[k].buf, t):.2e}" for k, t in taps.items() if k in sess.eng.taps and not k.startswith("b")))
        assert rel(pred, ref) < tol, (dtype, rel(pred, ref))
        assert sig < (1e-4 if dtype == "f32" else 5e-2)
        if dtype == "f32":
            for name, t in taps.items():
                if name in sess.eng.taps:
                    assert rel(sess.eng.taps[name].buf, t) < 2e-4, name
        del sess
        torch.cuda.empty_cache()


@pytest.mark.parametrize("graph", ["inference_p3d", "inference_p3d_decoder_block"])
def test_gn_training_step_parity_fp32(lib_built, graph):
    """one iteration of gn/train_p3d_gn_dataset.py:186-199 (smooth-L1 on the logits, Adam) in the fp32 path: loss,
    gradients of every variable (cosine / norm, see test_model_gpu.py for why not element-wise) and post-Adam values"""
    batch, size = 2, 64
    x = O.synthetic_clip(batch, 16, size, seed=0)
    y = O.synthetic_target(batch, 16, size, seed=1)
    vs = O.VarStore(seed=0)
    with torch.no_grad():
        O.forward(graph, x, vs, True)
    init = {k: v.clone() for k, v in vs.params.items()}
    loss_ref, grads_ref = O.train_step(graph, x, y, vs, {}, 1)
    sess = build(graph, "f32", True, batch, size)
    sess.eng.load_params(init)
    loss = float(sess.train_step(x.cuda(), y.cuda()).item())
    torch.cuda.synchronize()
    assert abs(loss - loss_ref) / loss_ref < 1e-4, (loss, loss_ref)
    gmax = max(float(g.norm()) for g in grads_ref.values())
    n = 0
    for name, g in sess.gradients().items():
        gr = grads_ref[name]
        if float(gr.norm()) < 1e-6 * gmax:
            continue
        a, b = g.float().cpu().reshape(-1), gr.reshape(-1)
        cos = float((a @ b) / (a.norm() * b.norm() + 1e-30))
        assert cos > 0.995, (name, cos)
        # (norms of the deepest CBAM gradients move by a few % with the summation order of the fp32 reductions: every
        #  reduce_max / ReLU decision that flips in stage 2 reroutes a whole gradient path)
        assert abs(float(a.norm() / b.norm()) - 1) < 0.10, name
        n += 1
    assert n > 700
    worst = max(rel(p, vs.params[name]) for name, p in sess.variables().items())
    # step-1 Adam moves every weight by ~lr*sign(g): sign flips of near-zero gradient entries bound the agreement of
    # small-valued variables (biases ~0.05) at a few lr/|w| = 1e-3
    assert worst < 5e-3, worst


@pytest.mark.parametrize("graph", ["inference_p3d", "inference_p3d_decoder_block"])
def test_gn_training_reduces_loss_bf16(lib_built, graph):
    batch, size = 2, 64
    x = O.synthetic_clip(batch, 16, size, seed=0).cuda()
    y = O.synthetic_target(batch, 16, size, seed=1).cuda()
    sess = build(graph, "bf16", True, batch, size, dropout=0.5)
    sess.lr = 1e-3
    losses = [float(sess.train_step(x, y, graph=True).item()) for _ in range(8)]
    assert all(l == l and l < 1e12 for l in losses)
    assert losses[-1] < losses[0]
