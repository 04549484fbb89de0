"""one replica of the NCCL data-parallel check (launched by test_zz_dp_nccl_gpu.py through torch.distributed.run, one process
per GPU): the production exchange path -- bf16 buckets, overlapped start / finish around the split backward graphs, CUDA-graph
replay -- on real NCCL.  Writes its findings as JSON to SAP3D_OUT.<rank>."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import p3d_oracle as O  # noqa: E402  (synthetic inputs only)


def checksum(t):
    bits = t.contiguous().view(torch.int32).to(torch.int64)
    return [int(bits.sum()), int((bits * (torch.arange(bits.numel(), device=t.device) % 8191 + 1)).sum())]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import sap3d_tensorflow_b200 as sp
    from sap3d_tensorflow_b200 import parallel

    graph, per, size, steps = "p3d_unetplusplus_ds", 2, 64, 3
    dev = f"cuda:{local}"
    x = O.synthetic_clip(per * world, 16, size, seed=0)[rank * per:(rank + 1) * per].to(dev)
    y = O.synthetic_target(per * world, 16, size, seed=1)[rank * per:(rank + 1) * per].to(dev)

    def build(seed):
        xin = sp.placeholder([per, 16, size, size, 3], dtype="bf16", training_graph=True, device=dev, dropout_seed=seed)
        return sp.Session(sp.p3d.p3d_unetplusplus_ds(xin, 0.5, per, True))

    # (1) local gradient of this replica, no exchange (same weights everywhere: rank 0's, broadcast below)
    sess = build(1234)
    parallel.attach_data_parallel(sess)                   # broadcast + bf16 exchange (SAP3D_DP_OVERLAP picks the form) + per-rank dropout seed
    w0 = sess.eng.flat_w.clone()
    seeds = [None] * world
    dist.all_gather_object(seeds, int(sess.eng.dropout_seed))
    hook = sess.grad_hook
    sess.grad_hook = None
    sess.train_step(x, y, graph=False)                    # local step (its Adam update is discarded below)
    g_local = sess.eng.flat_g[:sess.eng.n_train].clone()
    sess.eng.flat_w.copy_(w0); sess.eng.flat_m.zero_(); sess.eng.flat_v.zero_(); sess.eng.step.zero_()
    sess.eng.pack_weights()
    g_sum = g_local.clone()
    dist.all_reduce(g_sum)                                # fp32 reference of the exchanged gradient
    # (2) the production path: CUDA graphs + overlapped bf16 bucket exchange
    sess.grad_hook = hook
    sess.graph_train = None
    losses = []
    for i in range(steps):
        losses.append(float(sess.train_step(x, y, graph=True).item()))
        if i == 0:
            torch.cuda.synchronize()
            g_ex = sess.grad_hook.gradient()[:sess.eng.n_train]      # the all-reduced bf16 buckets Adam consumes
    torch.cuda.synchronize()
    rel = float((g_ex - g_sum).norm() / g_sum.norm())
    cs = checksum(sess.eng.flat_w[:sess.eng.n_train])
    all_cs = [None] * world
    dist.all_gather_object(all_cs, cs)
    out = {"rank": rank, "world": world, "exchanged_vs_fp32_sum_rel": rel, "checksums": all_cs, "losses": losses, "dropout_seeds": seeds,
           "overlap_graphs": 2 + len(sess.graph_train[2]), "segments": len(sess.eng.dp_segments), "overlap": bool(hook.overlap)}
    json.dump(out, open(os.environ["SAP3D_OUT"] + f".{rank}", "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
