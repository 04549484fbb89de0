"""one replica of the synchronised-BatchNorm parity run (launched by test_syncbn_gpu.py; RANK / WORLD_SIZE / SAP3D_PORT /
SAP3D_OUT in the environment).  Replicas share GPUs round-robin and talk over gloo, so the test needs only one GPU."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import p3d_oracle as O  # noqa: E402  (synthetic inputs only)


def build_graph(sp, graph, xin, batch):
    """a reference builder by name, or "shallow": stem -> pool1 -> the three stage-1 bottlenecks (ST_A with the projection
    shortcut, ST_B, ST_C) -> conv 3x3x3 + BN + ReLU -> 1-channel head.  Same ops and norm wirings as the full graphs, but 15
    BatchNorms instead of ~195, so rounding differences are not amplified and gradients can be compared tightly."""
    if graph != "shallow":
        return getattr(sp.p3d, graph)(xin, 0.0, batch, True)
    from sap3d_tensorflow_b200 import network as nw
    eng = xin.eng
    x = eng.maxpool(sp.p3d._stem(xin, True), (2, 3, 3), (2, 2, 2), name="pool1")
    x = sp.p3d.make_block(x, 64, 3, 64, 0).infer()
    d = nw.bn_relu(nw.layers_conv3d(x, 64, 3, 1, "dec"), True, name="dec_bn")
    w = eng.param("out/kernel", [3, 3, 3, 1, 64], "glorot_t")
    b = eng.param("out/bias", [1], "zeros")
    return eng.head(d, w, b, (3, 3, 3), 2, sigmoid=True, name="out")


def targets(graph, batch, size):
    return O.synthetic_target(batch, 16, size // 2 if graph == "shallow" else size, seed=1)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    graph, per, size, dtype = os.environ["SAP3D_GRAPH"], int(os.environ["SAP3D_PER"]), int(os.environ["SAP3D_SIZE"]), os.environ["SAP3D_DTYPE"]
    dev = f"cuda:{rank % torch.cuda.device_count()}"
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['SAP3D_PORT']}", rank=rank, world_size=world)
    import sap3d_tensorflow_b200 as sp
    from sap3d_tensorflow_b200 import parallel

    xin = sp.placeholder([per, 16, size, size, 3], dtype=dtype, training_graph=True, device=dev)
    head = build_graph(sp, graph, xin, per)
    sess = sp.Session(head)
    parallel.attach_data_parallel(sess, sync_bn=True, exact=True)
    x = O.synthetic_clip(per * world, 16, size, seed=0)[rank * per:(rank + 1) * per].to(dev)
    y = targets(graph, per * world, size)[rank * per:(rank + 1) * per].to(dev)
    loss = sess.train_step(x, y, graph=False)
    torch.cuda.synchronize()
    out = {
        "loss": float(loss.item()),
        "pred": head.output.detach().float().cpu(),
        "grads": {n: g.detach().cpu().clone() for n, g in sess.gradients().items()},
        "vars": {n: v.detach().cpu().clone() for n, v in sess.variables().items()},
        "sync_calls": sess.eng.sync_bn.calls,
    }
    try:
        sess.train_step(x, y, graph=True)
        out["graph_refused"] = False
    except sp._abi.Sap3dError:
        out["graph_refused"] = True
    torch.save(out, os.path.join(os.environ["SAP3D_OUT"], f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
