"""BASELINE.json configs[0] at its exact size: the reference's own runnable case -- `p3d_unetplusplus_ds(x, 0, 1, False)` forward
(gen_pred.py:45-46) on ONE 16 x 112 x 112 clip, batch 1 -- against the CPU oracle with the SURVEY §8d recipe (seed-0 uint8 clip
through `mapf`, randomised BatchNorm parameters and moving statistics, non-zero attention gates).  Tolerances are the
north-star's: saliency map within 1e-4 relative in the fp32 path and 1e-2 in bf16.  (The other forward-parity tests use 64 x 64
clips to keep the CPU oracle fast; this one costs it about a second.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import p3d_oracle as O  # noqa: E402


def test_reference_run_forward_batch1_112(lib_built):
    import sap3d_tensorflow_b200 as sp

    graph, batch, size = "p3d_unetplusplus_ds", 1, 112
    x = O.synthetic_clip(batch, 16, size, seed=0)
    vs = O.VarStore(seed=0)
    with torch.no_grad():
        ref = O.forward(graph, x, vs, False)
    assert tuple(ref.shape) == (1, 16, 112, 112, 1)
    for dtype, tol in (("f32", 1e-4), ("bf16", 1e-2)):
        xin = sp.placeholder([batch, 16, size, size, 3], dtype=dtype, training_graph=False)
        sess = sp.Session(sp.p3d.p3d_unetplusplus_ds(xin, 0, batch, False))
        sess.eng.load_params(vs.params)
        pred = sess.run(x.cuda(), graph=True).float().cpu()
        err = ((pred - ref).norm() / ref.norm()).item()
        print(f"configs[0] {dtype}: saliency rel err vs oracle {err:.3e}")
        assert err < tol, (dtype, err)
        del sess
        torch.cuda.empty_cache()
