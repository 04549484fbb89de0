"""Where the data-parallel step spends its time (run under torch.distributed.run, one process per GPU):
per-interval CUDA-event times on the main stream of rank 0 -- front graph (forward + last backward segment), every later
backward segment, the exchange tail (finish), the optimizer graph -- for three modes:
  none    : graphs split as in production, no collective at all (hook that does nothing)   -> cost of splitting the graph
  serial  : one exchange after the whole backward (no overlap)                              -> raw all-reduce time
  overlap : production
  *_1bucket : the same with one all-reduce call per segment instead of 32 MB buckets
Prints one line per mode.  usage: torchrun ... tools/dp_timeline.py [batch] [size] [steps]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class NoComm:
    def start(self, eng, k=0):
        pass

    def finish(self, eng):
        pass

    def __call__(self, eng):
        pass


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 112
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import sap3d_tensorflow_b200 as sp
    from sap3d_tensorflow_b200 import parallel

    dev = torch.device("cuda", local)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    x = ((torch.randint(0, 256, (B, 16, size, size, 3), generator=g).float() - torch.tensor([90.0, 102.0, 98.0])) / 255.0).to(dev)
    y = (torch.randint(0, 256, (B, 16, size, size), generator=g).float() / 255.0).to(dev)
    for mode in ("none", "serial", "serial_1bucket", "overlap", "overlap_1bucket"):
        xin = sp.placeholder([B, 16, size, size, 3], dtype="bf16", training_graph=True, device=f"cuda:{local}")
        sess = sp.Session(sp.p3d.p3d_unetplusplus_ds(xin, 0.5, B, True))
        if world > 1:
            parallel.attach_data_parallel(sess, bucket_mb=0 if mode.endswith("1bucket") else 32, overlap=True)
        if mode == "none" or world == 1:
            sess.grad_hook = NoComm()
            sess.eng.grad_source = None
        elif mode.startswith("serial"):
            hook = sess.grad_hook

            class Serial:            # no start/finish attributes -> Session captures the unsplit graphs
                def __call__(self, eng):
                    hook(eng)

            sess.grad_hook = Serial()
        e = sess.eng
        for _ in range(3):
            sess.train_step(x, y, graph=True)
        ga, gb, gms = sess.graph_train
        nseg = len(gms)
        names = ["front"] + [f"seg{k + 1}" for k in range(nseg)] + ["exchange_tail", "optimizer"]
        tot = [0.0] * len(names)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for _ in range(steps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            ev[0].record()
            ga.replay()
            ev[1].record()
            if gms:
                for k, gm in enumerate(gms):
                    sess.grad_hook.start(e, k)
                    gm.replay()
                    ev[2 + k].record()
                sess.grad_hook.finish(e)
            else:
                sess.grad_hook(e)
            ev[2 + nseg].record()
            gb.replay()
            ev[3 + nseg].record()
            torch.cuda.synchronize()
            for i in range(len(names)):
                tot[i] += ev[i].elapsed_time(ev[i + 1])
        if rank == 0:
            parts = "  ".join(f"{n} {t / steps:.3f}" for n, t in zip(names, tot))
            print(f"[dp_timeline] world={world} mode={mode} segments={len(e.dp_segments)} total {sum(tot) / steps:.3f} ms :: {parts}", flush=True)
            print("   segments (ops_lo, ops_hi, grad_lo, grad_hi):", e.dp_segments, "n_train", e.n_train, flush=True)
        del sess
        torch.cuda.synchronize()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
