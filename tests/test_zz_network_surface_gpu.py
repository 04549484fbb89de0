"""The callable surface of the reference's utils/network.py that its drivers use directly: smooth_l1_loss (network.py:49-62,
called at train.py:159 / gn/train_p3d_gn_dataset.py:186), GroupNorm / normalize(mode='gn') (network.py:65-94) and the
stand-alone cbam_block / channel_attention / spatial_attention (network.py:198-274) -- forward, loss and every gradient of a
small graph built only from those calls, against torch autograd over the oracle's restatement of the same functions."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import p3d_oracle as O  # noqa: E402
from oracle import tf_semantics as tfs  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _ref_loss(pred, y, sigma, w_in, w_out):
    s2 = sigma ** 2
    d = w_in * (pred - y)
    a = d.abs()
    sign = (a < 1.0 / s2).to(d.dtype)
    return (w_out * (d * d * (s2 / 2.0) * sign + (a - 0.5 / s2) * (1.0 - sign))).sum()


def _oracle(x, params, mode, y, sigma, w_in, w_out):
    """stem conv -> GroupNorm (no ReLU: the feature keeps its negative half) -> CBAM variant -> 3x3x3 conv to 1 channel -> loss"""
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    vs = O.VarStore(seed=0, params=p)
    ctx = O.Ctx(vs, True)
    f = tfs.conv3d_same(x, p["firstconv1"], (1, 2, 2))
    f = tfs.group_norm(f, p["group_norm/gamma"], p["group_norm/beta"])
    n, d, h, w, c = f.shape
    if mode == "both":
        g = O.cbam_block(ctx, f, "att")
    elif mode == "channel":
        w0, b0, w1, b1 = (p["att/mlp_0/kernel"], p["att/mlp_0/bias"], p["att/mlp_1/kernel"], p["att/mlp_1/bias"])
        mlp = lambda v: torch.relu(v @ w0 + b0) @ w1 + b1  # noqa: E731
        g = f * torch.sigmoid(mlp(f.mean(dim=(1, 2, 3))) + mlp(f.amax(dim=(1, 2, 3)))).view(n, 1, 1, 1, c)
    else:
        cat = torch.cat([f.mean(dim=4, keepdim=True), f.amax(dim=4, keepdim=True)], dim=4)
        g = f * torch.sigmoid(tfs.conv3d_same(cat, p["att/conv3d/kernel"], (1, 1, 1)))
    logits = tfs.conv3d_same(g, p["results/kernel"], (1, 1, 1), p["results/bias"])
    loss = _ref_loss(logits.reshape(y.shape), y, sigma, w_in, w_out)
    loss.backward()
    return logits.detach(), float(loss.detach()), {k: v.grad for k, v in p.items()}, g.detach()


@pytest.mark.parametrize("mode", ["both", "channel", "spatial"])
def test_network_surface_small_graph(lib_built, mode):
    import sap3d_tensorflow_b200 as sp
    from sap3d_tensorflow_b200 import network as nw
    from sap3d_tensorflow_b200.p3d import get_conv_weight

    B, size = 2, 32
    sigma, w_in, w_out = 3.0, 2.0, 0.5
    x = O.synthetic_clip(B, 16, size, seed=5)
    y = O.synthetic_target(B, 16, size // 2, seed=6) * 0.2          # differences straddle 1/sigma^2: both loss branches
    xin = sp.placeholder([B, 16, size, size, 3], dtype="f32", training_graph=True)
    eng = xin.eng
    c = eng.conv([xin], 64, (1, 7, 7), (1, 2, 2), get_conv_weight(eng, "firstconv1", [1, 7, 7, 3, 64]), name="firstconv1", want_stats=False)
    f = nw.normalize(c, True, mode="gn")                             # == GroupNorm(c)
    eng.tap("feature", f)
    if mode == "both":
        g = nw.cbam_block(f, "att")
    elif mode == "channel":
        g = nw.channel_attention(f, "att")
    else:
        g = nw.spatial_attention(f, "att")
    eng.tap("refined", g)
    w = eng.param("results/kernel", [3, 3, 3, 64, 1], "glorot")
    b = eng.param("results/bias", [1], "zeros")
    head = eng.logits_loss(eng.conv([g], 1, (3, 3, 3), (1, 1, 1), w, b, want_stats=False, name="results", out_f32=True), name="results")
    loss_h = nw.smooth_l1_loss(head, None, w_in, w_out, sigma=sigma)
    sess = sp.Session(loss_h)
    g0 = torch.Generator().manual_seed(3)
    params = {}
    for n, p in eng.params.items():     # filters keep their TF initialisation; norm affines and biases get non-trivial values
        if p.kind == "ones":
            params[n] = 1.0 + 0.3 * torch.randn(p.shape, generator=g0)
        elif p.kind in ("zeros", "bias"):
            params[n] = 0.1 * torch.randn(p.shape, generator=g0)
        else:
            params[n] = p.w.detach().cpu().clone()
    sess.eng.load_params(params)
    loss = float(sess.train_step(x.cuda(), y.cuda()).item())
    torch.cuda.synchronize()
    logits_ref, loss_ref, grads_ref, refined_ref = _oracle(x, params, mode, y, sigma, w_in, w_out)
    assert rel(sess.tap("refined"), refined_ref) < 1e-5, rel(sess.tap("refined"), refined_ref)
    assert float(refined_ref.min()) < 0 < float(refined_ref.max())        # the stand-alone op must not clamp negatives
    assert rel(sess.head.output, logits_ref) < 1e-5
    assert abs(loss - loss_ref) / abs(loss_ref) < 1e-5, (loss, loss_ref)
    for name, g_ in sess.gradients().items():
        gr = grads_ref[name]
        assert rel(g_, gr) < 2e-4, (name, rel(g_, gr))
    assert "group_norm/gamma" in eng.params


def test_smooth_l1_loss_defaults_match_the_drivers_call(lib_built):
    """train.py:159: smooth_l1_loss(pred, y, 1, 1, sigma=1.0) on the saliency head == the fused default path of train_step"""
    import sap3d_tensorflow_b200 as sp
    from sap3d_tensorflow_b200 import network as nw

    B, size = 1, 32
    x = O.synthetic_clip(B, 16, size, seed=0).cuda()
    y = O.synthetic_target(B, 16, size, seed=1).cuda()
    out = []
    for wrap in (False, True):
        xin = sp.placeholder([B, 16, size, size, 3], dtype="bf16", training_graph=True)
        head = sp.p3d.p3d_unet(xin, 0.0, B, True)
        sess = sp.Session(nw.smooth_l1_loss(head, None, 1, 1, sigma=1.0) if wrap else head)
        out.append(float(sess.train_step(x, y).item()))
        pred = sess.head.output.double().reshape(y.shape)
        expect = float(tfs.smooth_l1_loss(pred.cpu(), y.double().cpu()))
        assert abs(out[-1] - expect) / expect < 1e-6
    assert abs(out[0] - out[1]) / out[0] < 1e-5
    with pytest.raises(sp._abi.Sap3dError):
        nw.smooth_l1_loss(head, None, torch.ones(3), 1)
