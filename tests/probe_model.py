"""GPU probe: whole-graph forward / training-step parity against the CPU oracle.
    python tests/probe_model.py [graph] [size] [batch] > gpurun_out/probe_model.log 2>&1
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import p3d_oracle as O  # noqa: E402
import sap3d_tensorflow_b200 as sp  # noqa: E402
from sap3d_tensorflow_b200 import _abi as A  # noqa: E402


def relerr(a, b):
    a = a.detach().float().cpu().reshape(-1)
    b = b.detach().float().cpu().reshape(-1)
    return ((a - b).norm() / (b.norm() + 1e-30)).item(), (a - b).abs().max().item()


def main():
    graph = sys.argv[1] if len(sys.argv) > 1 else "p3d_unetplusplus_nonsa"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    torch.set_num_threads(os.cpu_count())
    x = O.synthetic_clip(batch, 16, size, seed=0)
    y = O.synthetic_target(batch, 16, size, seed=1)
    builder = getattr(sp.p3d, graph) if hasattr(sp.p3d, graph) else getattr(sp.gn.p3d_gn, graph)
    modes = ("eval",) if graph.startswith("inference") else ("eval", "train")
    ok = True
    for mode in modes:
        training = mode == "train"
        vs = O.VarStore(seed=0)
        taps_ref = {}
        t0 = time.time()
        with torch.no_grad():
            ref = O.forward(graph, x, vs, training, taps=taps_ref)
        print(f"[{mode}] oracle forward {time.time() - t0:.1f}s, {len(vs.params)} variables", flush=True)
        for dtype in ("f32", "bf16"):
            xin = sp.placeholder([batch, 16, size, size, 3], dtype=dtype, training_graph=training)
            head = builder(xin, 0.0, batch, training)
            sess = sp.Session(head)
            missing = set(sess.eng.params) ^ set(vs.params)
            if missing:
                print("  VARIABLE NAME MISMATCH:", sorted(missing)[:10], flush=True)
                ok = False
            sess.eng.load_params(vs.params, strict=False)
            pred = sess.run(x.cuda())
            torch.cuda.synchronize()
            tol = 2e-4 if dtype == "f32" else 3e-2
            worst = 0.0
            for name, t in taps_ref.items():
                if name in sess.eng.taps:
                    r, m = relerr(sess.eng.taps[name].buf, t)
                    worst = max(worst, r)
                    flag = "" if r < tol else "  <-- FAIL"
                    if flag or name in ("stem", "pool1", "x_1_0", "x_2_0", "x_3_0", "x_4_0", "b0", "b2", "b10", "b46", "x_3_1", "x_2_2", "x_1_3", "upx_4_0"):
                        print(f"  [{mode}/{dtype}] tap {name:12s} rel={r:.2e} max={m:.2e}{flag}", flush=True)
            r, m = relerr(pred, ref)
            print(f"  [{mode}/{dtype}] pred rel={r:.2e} max={m:.2e} worst-tap={worst:.2e} launches={sess.eng.launches_fwd}", flush=True)
            ok = ok and r < tol
            # CUDA-graph replay must give the same result
            pred_g = sess.run(x.cuda(), graph=True).clone()
            torch.cuda.synchronize()
            rg, _ = relerr(pred_g, pred)
            print(f"  [{mode}/{dtype}] graph-replay vs eager rel={rg:.2e}", flush=True)
            ok = ok and rg < 1e-6
            if training:
                vs2 = O.VarStore(seed=0)
                with torch.no_grad():
                    O.forward(graph, x, vs2, True)
                adam = {}
                t0 = time.time()
                loss_ref, grads_ref = O.train_step(graph, x, y, vs2, adam, 1)
                print(f"  oracle train step {time.time() - t0:.1f}s loss={loss_ref:.4f}", flush=True)
                sess.eng.load_params(vs.params, strict=False)
                loss = sess.train_step(x.cuda(), y.cuda())
                torch.cuda.synchronize()
                if dtype == "f32":
                    # activation-gradient comparison (localises backward bugs)
                    vs3 = O.VarStore(seed=0, params={k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in vs.params.items()})
                    taps3 = {}
                    pr3 = O.forward(graph, x, vs3, True, taps=taps3)
                    for t in taps3.values():
                        if t.requires_grad:
                            t.retain_grad()
                    O.tfs.smooth_l1_loss(pr3.reshape(y.shape), y).backward()
                    for name in reversed(list(taps3.keys())):
                        t = taps3[name]
                        et = sess.eng.taps.get(name)
                        if t.grad is None or et is None or et.grad is None:
                            continue
                        r, m = relerr(et.grad, t.grad)
                        print(f"    dtap {name:12s} rel={r:.2e} max={m:.2e} gflag={et.gflag}", flush=True)
                lv = float(loss.item())
                print(f"  [{mode}/{dtype}] loss={lv:.4f} ref={loss_ref:.4f} rel={(lv - loss_ref) / loss_ref:.2e} bwd launches={sess.eng.launches_bwd}", flush=True)
                gtol = 5e-3 if dtype == "f32" else 1e-1
                bad = 0
                worst = (0.0, "")
                for name, g in sess.gradients().items():
                    gr = grads_ref[name]
                    if gr.abs().max() < 1e-6 * max(1.0, float(g.abs().max())) and g.abs().max() < 1e-4:
                        continue  # mathematically-zero gradients (bias before batch-stat BN)
                    r, m = relerr(g, gr)
                    if r > worst[0]:
                        worst = (r, name)
                    if r > gtol:
                        bad += 1
                        if bad <= 12:
                            print(f"    grad {name:40s} rel={r:.2e} max={m:.2e} |ref|={gr.abs().max():.2e}", flush=True)
                print(f"  [{mode}/{dtype}] grads: {bad} above tol {gtol}, worst {worst[0]:.2e} ({worst[1]})", flush=True)
                ok = ok and bad == 0
                # parameters after Adam and BN moving statistics
                wr = 0.0
                for name, p in sess.variables().items():
                    r, _ = relerr(p, vs2.params[name])
                    wr = max(wr, r)
                print(f"  [{mode}/{dtype}] post-step variables worst rel={wr:.2e}", flush=True)
            del sess, head, xin
            torch.cuda.empty_cache()
    print("ALL PASS" if ok else "SOME FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
