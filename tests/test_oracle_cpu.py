"""CPU tests of the oracle itself: TF semantics vs an independent fp64 direct-loop restatement, graph
structure vs the parameter counts derived from the reference, loss / Adam formulas."""
import math

import numpy as np
import pytest
import torch

from oracle import np_direct as npd
from oracle import p3d_oracle as O
from oracle import tf_semantics as tfs


def test_same_padding_cases_from_the_reference():
    # stem: I=112,k=7,s=2 -> (2,3); pool1 H/W: I=56,k=3,s=2 -> (0,1); k=2,s=1 -> (0,1); 1x1 s2 -> 0
    assert tfs.same_pad(112, 7, 2) == (56, 2, 3)
    assert tfs.same_pad(56, 3, 2) == (28, 0, 1)
    assert tfs.same_pad(2, 2, 1) == (2, 0, 1)
    assert tfs.same_pad(28, 1, 2) == (14, 0, 0)
    assert tfs.same_pad(16, 2, 2) == (8, 0, 0)


CONV_CASES = [
    ((1, 3, 12, 10, 3), (1, 7, 7), (1, 2, 2)),   # stem geometry (even size, k7 s2: pads (2,3))
    ((2, 4, 5, 6, 4), (1, 3, 3), (1, 1, 1)),     # convS
    ((2, 4, 5, 6, 4), (3, 1, 1), (1, 1, 1)),     # convT
    ((1, 2, 5, 5, 4), (2, 3, 3), (1, 1, 1)),     # x_3_1: k_d = 2 pads only AFTER in D
    ((1, 4, 6, 6, 4), (1, 1, 1), (1, 2, 2)),     # strided 1x1x1 (samples even indices)
    ((1, 3, 5, 7, 2), (3, 3, 3), (1, 1, 1)),
]


@pytest.mark.parametrize("xs,k,s", CONV_CASES)
def test_conv3d_same_matches_direct_loops(xs, k, s):
    rng = np.random.RandomState(0)
    x = rng.randn(*xs)
    w = rng.randn(*k, xs[-1], 5)
    b = rng.randn(5)
    ref = npd.conv3d_same(x, w, s, b)
    got = tfs.conv3d_same(torch.tensor(x), torch.tensor(w), s, torch.tensor(b)).numpy()
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-10)


DECONV_CASES = [
    ((1, 2, 3, 3, 4), (3, 3, 3), (2, 2, 2)),     # upx_2_x
    ((1, 1, 3, 3, 4), (1, 3, 3), (2, 2, 2)),     # upx_4_0: odd-D planes hold only the bias
    ((1, 2, 3, 3, 4), (2, 3, 3), (2, 2, 2)),     # upx_3_x
    ((1, 1, 2, 2, 4), (3, 3, 3), (4, 4, 4)),     # p3d_concat deconv_pool4: k < s holes
    ((1, 2, 3, 4, 4), (3, 3, 3), (1, 1, 1)),     # p3d_concat deconv_pool2: pb = 1
]


@pytest.mark.parametrize("xs,k,s", DECONV_CASES)
def test_conv3d_transpose_matches_direct_loops(xs, k, s):
    rng = np.random.RandomState(1)
    x = rng.randn(*xs)
    w = rng.randn(*k, 3, xs[-1])
    b = rng.randn(3)
    ref = npd.conv3d_transpose_same(x, w, s, b)
    got = tfs.conv3d_transpose_same(torch.tensor(x), torch.tensor(w), s, torch.tensor(b)).numpy()
    assert got.shape == ref.shape == (xs[0], xs[1] * s[0], xs[2] * s[1], xs[3] * s[2], 3)
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-10)


def test_transposed_conv_is_the_gradient_of_the_same_conv():
    # tf.layers.conv3d_transpose 'same' == input-gradient of a SAME conv from the I*s tensor
    torch.manual_seed(0)
    x = torch.randn(1, 2, 3, 3, 4, dtype=torch.float64)
    w = torch.randn(3, 3, 3, 5, 4, dtype=torch.float64)  # [k, Cout, Cin] of the transpose == DHWIO of the forward conv
    big = torch.zeros(1, 4, 6, 6, 5, dtype=torch.float64, requires_grad=True)
    y = tfs.conv3d_same(big, w, (2, 2, 2))
    (y * x).sum().backward()
    got = tfs.conv3d_transpose_same(x, w, (2, 2, 2))
    torch.testing.assert_close(got, big.grad)


@pytest.mark.parametrize("k,s", [((2, 1, 1), (2, 1, 1)), ((2, 3, 3), (2, 2, 2))])
def test_max_pool_same(k, s):
    rng = np.random.RandomState(2)
    x = rng.randn(2, 4, 6, 6, 3) - 3.0  # all negative: zero padding would win, -inf padding must not
    ref = npd.max_pool3d_same(x, k, s)
    got = tfs.max_pool3d_same(torch.tensor(x), k, s).numpy()
    np.testing.assert_array_equal(got, ref)


def test_batch_norm_training_and_inference():
    torch.manual_seed(0)
    x = torch.randn(2, 3, 4, 5, 6, dtype=torch.float64) * 2 + 1
    g, b = torch.rand(6, dtype=torch.float64) + 0.5, torch.randn(6, dtype=torch.float64)
    mm, mv = torch.zeros(6, dtype=torch.float64), torch.ones(6, dtype=torch.float64)
    y, nmm, nmv = tfs.batch_norm(x, g, b, mm, mv, True)
    flat = x.reshape(-1, 6)
    mean, var = flat.mean(0), flat.var(0, unbiased=False)
    torch.testing.assert_close(y.reshape(-1, 6), (flat - mean) / torch.sqrt(var + 1e-3) * g + b)
    torch.testing.assert_close(nmm, 0.01 * mean)
    torch.testing.assert_close(nmv, 0.99 + 0.01 * var)  # biased variance (non-fused 5-D TF path)
    y2, _, _ = tfs.batch_norm(x, g, b, mm, mv, False)
    torch.testing.assert_close(y2, x / math.sqrt(1 + 1e-3) * g + b)


def test_group_norm_matches_reference_formula():
    # network.py:65-87 literally: transpose to NCDHW, reshape [N,G,C/G,D,H,W], moments over [2,3,4,5]
    torch.manual_seed(1)
    x = torch.randn(2, 2, 3, 3, 64, dtype=torch.float64)
    g, b = torch.rand(64, dtype=torch.float64), torch.randn(64, dtype=torch.float64)
    xt = x.permute(0, 4, 1, 2, 3).reshape(2, 32, 2, 2, 3, 3)
    mean = xt.mean(dim=(2, 3, 4, 5), keepdim=True)
    var = xt.var(dim=(2, 3, 4, 5), keepdim=True, unbiased=False)
    ref = ((xt - mean) / torch.sqrt(var + 1e-5)).reshape(2, 64, 2, 3, 3) * g.view(1, 64, 1, 1, 1) + b.view(1, 64, 1, 1, 1)
    torch.testing.assert_close(tfs.group_norm(x, g, b), ref.permute(0, 2, 3, 4, 1))


def test_smooth_l1_is_half_squared_error_for_sigmoid_outputs():
    p, t = torch.rand(4, 5), torch.rand(4, 5)
    torch.testing.assert_close(tfs.smooth_l1_loss(p, t), 0.5 * ((p - t) ** 2).sum())
    d = torch.tensor([2.0, -3.0, 0.5])
    torch.testing.assert_close(tfs.smooth_l1_loss(d, torch.zeros(3)), torch.tensor(1.5 + 2.5 + 0.125))


def test_adam_is_the_tf_formula():
    p, g = torch.tensor([1.0], dtype=torch.float64), torch.tensor([0.5], dtype=torch.float64)
    p1, m, v = tfs.adam_step_tf(p, g, torch.zeros(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64), 1)
    lr_t = 1e-4 * math.sqrt(1 - 0.999) / (1 - 0.9)
    assert abs(p1.item() - (1.0 - lr_t * 0.05 / (math.sqrt(0.00025) + 1e-8))) < 1e-9


# parameter counts derived from the reference graphs: kernels + biases + 2 per norm channel,
# attention gammas and BN moving statistics excluded
PARAM_COUNTS = {
    "p3d_unetplusplus_ds": 84_921_761,
    "p3d_unetplusplus_nonsa": 81_780_161,
    "p3d_unet": 61_943_105,
    "p3d_concat": 83_472_065,
    "inference_p3d": 152_846_483,
    "inference_p3d_decoder_block": 103_446_323,
}


@pytest.mark.parametrize("graph", list(PARAM_COUNTS))
def test_graph_structure_matches_reference_parameter_count(graph):
    x = O.synthetic_clip(1, 16, 32, seed=0)
    vs = O.VarStore(seed=0)
    with torch.no_grad():
        y = O.forward(graph, x, vs, True)
    assert tuple(y.shape) == (1, 16, 32, 32, 1)
    n = sum(v.numel() for k, v in vs.params.items() if vs.trainable[k] and not k.startswith("gamma"))
    assert n == PARAM_COUNTS[graph]


def test_variable_names_follow_the_reference():
    x = O.synthetic_clip(1, 16, 32, seed=0)
    vs = O.VarStore(seed=0)
    with torch.no_grad():
        O.forward("p3d_unetplusplus_ds", x, vs, True)
    names = set(vs.params)
    for n in ["firstconv1", "conv3_0_1", "STA_0_2_S", "STA_0_2_S_bias", "STB_1_2_T", "STC_2_2_S", "dw3d_0", "dw3d_3", "dw3d_11",
              "conv3_46_3", "batch_normalization/gamma", "batch_normalization_1/moving_mean", "upx_4_0/kernel", "x_3_1/bias",
              "x_4_0_sa/conv3d/kernel", "x_4_0_sa/conv3d_2/bias", "gammax_1_3_sa", "x_0_1/kernel", "conv3d/kernel", "conv3d_3/kernel"]:
        assert n in names, n
    assert vs.params["upx_4_0/kernel"].shape == (1, 3, 3, 512, 1024)   # transposed conv: [k, Cout, Cin]
    assert vs.params["x_3_1/kernel"].shape == (2, 3, 3, 1024, 512)
    assert sum(1 for n in names if n.startswith("dw3d_")) == 3
    assert sum(1 for n in names if n.endswith("/moving_mean")) == 208


def test_train_step_decreases_loss_on_repeated_batch():
    x = O.synthetic_clip(1, 16, 32, seed=0)
    y = O.synthetic_target(1, 16, 32, seed=1)
    vs = O.VarStore(seed=0)
    with torch.no_grad():
        O.forward("p3d_unet", x, vs, True)
    adam = {}
    losses = [O.train_step("p3d_unet", x, y, vs, adam, i + 1, lr=1e-3)[0] for i in range(6)]
    assert losses[-1] < losses[0] and all(math.isfinite(v) for v in losses)
