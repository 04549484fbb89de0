"""Parity of the CUDA path against the CPU oracle AT THE BENCHMARK'S OWN CONFIGURATIONS (BASELINE.json configs[1], [2]).

configs[1]  p3d_unetplusplus_ds, 8 clips of 16 x 112 x 112, training step (reference train.py:156-172,217; p3d.py:340-399):
            saliency map, loss, BatchNorm moving statistics, gradients and the post-Adam variables.
configs[2]  gn/p3d_gn.inference_p3d (GroupNorm + CBAM) at 16 x 160 x 160 (reference gn/p3d_gn.py:214-258); the oracle runs
            2 clips, not the benchmark's 16: GroupNorm / CBAM statistics are per sample, so the batch size does not change
            the conditioning of the comparison, and 16 clips of fp32 autograd do not fit the CPU side's time budget.

Three references per configuration:
  fp32 oracle                      the reference graph in fp32
  bf16-storage oracle              the same graph with every HBM storage point of the CUDA path rounded to bf16
                                   (oracle.p3d_oracle.Ctx(bf16=True)); accumulation / statistics stay fp32
  d(fp32 oracle, bf16 oracle)      how far ANY bf16-storage implementation of this random-weight network is from the fp32
                                   graph -- printed, and used as the yardstick for "bf16 CUDA path vs fp32 oracle".
Tolerances (north_star): fp32 path 1e-4; bf16 path 1e-2.

Measured r02 (B200, configs[1]): with the default synthetic weights the 47 batch-statistics blocks amplify any perturbation
~200x from stem to output: the bf16-storage ORACLE is 7.4e-2 away from the fp32 oracle on the saliency map (gradient cosine
0.07), so no bf16 implementation can meet 1e-2 end to end there, and even equal rounding points do not help (one differing
rounding in a million grows to 6e-2).  The tests therefore run each configuration with TWO sets of synthetic weights:
  "default"            gamma of the last norm of every residual branch ~ U(0.1, 0.3): fp32 path strict; the bf16 path is
                       bounded by the printed storage noise d(bf16 oracle, fp32 oracle);
  "well-conditioned"   ~ U(0.02, 0.06) (residual branches as small as in a zero-init-residual / trained ResNet), attention
                       gates ~ U(0.05, 0.15): fp32 strict AND bf16 strict (1e-2 on the saliency map against BOTH oracles).
"""
import os
import time

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import p3d_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _grad_agreement(got, ref):
    """global cosine over all variables, worst per-variable cosine among variables that carry gradient mass, and the share
    of the gradient mass whose Adam update direction (sign) agrees"""
    gmax = max(float(g.norm()) for g in ref.values())
    num = den_a = den_b = 0.0
    worst, worst_name, n = 1.0, "", 0
    agree_w = tot_w = 0.0
    for name, gr in ref.items():
        if float(gr.norm()) < 1e-5 * gmax:      # mathematically-zero gradients (bias in front of a batch-statistics norm)
            continue
        a, b = got[name].double().cpu().reshape(-1), gr.double().reshape(-1)
        num += float(a @ b); den_a += float(a @ a); den_b += float(b @ b)
        cos = float((a @ b) / (a.norm() * b.norm() + 1e-300))
        if cos < worst:
            worst, worst_name = cos, name
        agree_w += float((b.abs() * (torch.sign(a) == torch.sign(b))).sum())
        tot_w += float(b.abs().sum())
        n += 1
    return num / (den_a ** 0.5 * den_b ** 0.5 + 1e-300), worst, worst_name, agree_w / tot_w, n


def _oracle_step(graph, x, y, init, bf16):
    vs = O.VarStore(seed=0, params={k: v.clone() for k, v in init.items()})
    vs.frozen = False
    vs.trainable = dict(TRAINABLE)
    t0 = time.time()
    taps = {}
    loss, grads = O.train_step(graph, x, y, vs, {}, 1, bf16=bf16, taps=taps)
    pred = O.train_step.last_pred
    taps = {k: v.detach() for k, v in taps.items()}
    print(f"  oracle {'bf16-storage' if bf16 else 'fp32'} training step: {time.time() - t0:.1f} s, loss {loss:.4f}", flush=True)
    return loss, grads, pred, vs.params, taps


TRAINABLE = {}
DIAG_TAPS = ("b0", "b2", "b5", "b10", "b15", "b20", "b30", "b40", "b46")
FAILURES = []     # every comparison is made and printed before the test asserts (one GPU run shows all numbers)


def _engine_step(builder, dtype, batch, size, init, x, y):
    import sap3d_tensorflow_b200 as sp

    xin = sp.placeholder([batch, 16, size, size, 3], dtype=dtype, training_graph=True)
    sess = sp.Session(builder(xin, 0.0, batch, True))
    assert set(sess.eng.params) == set(init)
    sess.eng.load_params(init)
    loss = float(sess.train_step(x.cuda(), y.cuda()).item())
    torch.cuda.synchronize()
    pred = sess.head.output.float().cpu().clone()
    grads = {n: g.detach().float().cpu().clone() for n, g in sess.gradients().items()}
    variables = {n: v.detach().float().cpu().clone() for n, v in sess.variables().items()}
    taps = {n: t.buf.float().cpu().clone() for n, t in sess.eng.taps.items()}
    del sess
    torch.cuda.empty_cache()
    return loss, grads, pred, variables, taps


def _compare(tag, got, ref, tol_pred, tol_loss, tol_stats, cos_min, shape, tol_w=1e-3):
    loss, grads, pred, variables = got[:4]
    loss_r, grads_r, pred_r, vars_r = ref[:4]
    e_pred = rel(torch.sigmoid(pred) if shape == "logits" else pred, torch.sigmoid(pred_r) if shape == "logits" else pred_r)
    e_loss = abs(loss - loss_r) / abs(loss_r)
    e_stats = max([rel(variables[n], vars_r[n]) for n in variables if n.endswith(("moving_mean", "moving_variance"))] or [0.0])
    cos, worst, worst_name, agree, n = _grad_agreement(grads, grads_r)
    # post-Adam variables: the first TF-Adam step moves every weight by ~lr * sign(g); compare the UPDATES, not the weights
    tr = [k for k in grads_r]
    e_w = rel(torch.cat([variables[k].reshape(-1) for k in tr]), torch.cat([vars_r[k].reshape(-1) for k in tr]))
    print(f"  [{tag}] saliency rel {e_pred:.3e} | loss rel {e_loss:.3e} | moving stats worst rel {e_stats:.3e} | gradient cosine "
          f"global {cos:.5f}, worst variable {worst:.4f} ({worst_name}), update-direction agreement {agree:.4f} over {n} variables | "
          f"post-Adam variables rel {e_w:.3e}", flush=True)
    for what, v, ok in (("saliency", e_pred, e_pred < tol_pred), ("loss", e_loss, e_loss < tol_loss),
                        ("moving statistics", e_stats, e_stats < tol_stats), ("gradient cosine", cos, cos > cos_min),
                        ("post-Adam variables", e_w, e_w < tol_w)):
        if not ok:
            FAILURES.append((tag, what, v))
    return e_pred


CONDITIONING = {"default": {}, "well-conditioned": {"gamma_res": (0.02, 0.06), "sa_gamma": (0.05, 0.15)}}


def _run_config(graph, builder, batch, size, out_kind, conditioning):
    torch.set_num_threads(os.cpu_count() or 1)
    strict = conditioning == "well-conditioned"
    x = O.synthetic_clip(batch, 16, size, seed=0)
    y = O.synthetic_target(batch, 16, size, seed=1)
    vs = O.VarStore(seed=0, **CONDITIONING[conditioning])
    with torch.no_grad():
        O.forward(graph, x, vs, True)
    init = {k: v.clone() for k, v in vs.params.items()}
    TRAINABLE.clear()
    TRAINABLE.update(vs.trainable)
    FAILURES.clear()
    print(f"\n{graph}: {batch} clips of 16 x {size} x {size}, training step, {conditioning} synthetic weights", flush=True)
    ref32 = _oracle_step(graph, x, y, init, bf16=False)
    refbf = _oracle_step(graph, x, y, init, bf16=True)
    sal = (lambda p: torch.sigmoid(p)) if out_kind == "logits" else (lambda p: p)
    d_pred = rel(sal(refbf[2]), sal(ref32[2]))
    d_loss = abs(refbf[0] - ref32[0]) / ref32[0]
    cos, worst, wn, agree, _ = _grad_agreement(refbf[1], ref32[1])
    print(f"  d(bf16-storage oracle, fp32 oracle): saliency rel {d_pred:.3e}, loss rel {d_loss:.3e}, gradient cosine global {cos:.5f} "
          f"worst {worst:.4f} ({wn}), update-direction agreement {agree:.4f}   <- bf16 storage noise of ANY implementation", flush=True)
    got32 = _engine_step(builder, "f32", batch, size, init, x, y)
    _compare("fp32 CUDA vs fp32 oracle", got32, ref32, 1e-4, 1e-5, 1e-4, 0.999 if strict else 0.99, out_kind)
    del got32
    gotbf = _engine_step(builder, "bf16", batch, size, init, x, y)
    # strict: 1e-2 (north_star); otherwise the yardstick is the storage noise the oracle itself shows.  Gradients of the bf16
    # path are asserted relative to how well the two ORACLES agree with each other (the engine also stores gradients in bf16).
    bound = 1e-2 if strict else 1.5 * d_pred + 1e-2
    e_eq = _compare("bf16 CUDA vs bf16-storage oracle", gotbf, refbf, bound, 1e-2, 1e-2 if strict else 5e-2, 0.8 * cos - 0.1, out_kind, tol_w=5e-3)
    e_32 = rel(sal(gotbf[2]), sal(ref32[2]))
    print(f"  [bf16 CUDA vs fp32 oracle] saliency rel {e_32:.3e} (bound {bound:.3e}: "
          f"{'north_star 1e-2' if strict else '1.5 x d(bf16 oracle, fp32 oracle) + 1e-2'})", flush=True)
    if not e_32 < bound:
        FAILURES.append(("bf16 CUDA vs fp32 oracle", "saliency", e_32))
    # where the bf16 path leaves the equal-rounding oracle, tap by tap (diagnostic; the same taps for the storage noise itself)
    for name, t in refbf[4].items():
        if name in gotbf[4] and name in ref32[4] and (name in DIAG_TAPS or not name.startswith("b")):
            print(f"    tap {name:12s} bf16 CUDA vs bf16 oracle {rel(gotbf[4][name], t):.3e} | bf16 oracle vs fp32 oracle {rel(t, ref32[4][name]):.3e}",
                  flush=True)
    assert not FAILURES, FAILURES
    return e_eq, d_pred


@pytest.mark.parametrize("conditioning", ["well-conditioned", "default"])
def test_config1_training_step_b8_112(lib_built, conditioning):
    """BASELINE configs[1] exactly: p3d_unetplusplus_ds, batch 8, 16 x 112 x 112, training=True, dropout 0"""
    import sap3d_tensorflow_b200 as sp

    _run_config("p3d_unetplusplus_ds", sp.p3d.p3d_unetplusplus_ds, 8, 112, "pred", conditioning)


@pytest.mark.parametrize("conditioning", ["well-conditioned", "default"])
def test_config2_gn_cbam_training_step_160(lib_built, conditioning):
    """BASELINE configs[2] at its spatial size: gn/p3d_gn.inference_p3d, 16 x 160 x 160, training step (2 clips, see above)"""
    from sap3d_tensorflow_b200.gn import p3d_gn

    _run_config("inference_p3d", p3d_gn.inference_p3d, 2, 160, "logits", conditioning)
