"""Generates tests/golden/reference_graphs_golden.npz by EXECUTING THE REFERENCE'S OWN builder code (/root/reference/p3d.py,
utils/network.py, gn/p3d_gn.py -- read where they lie, never copied) over the TF-1.x emulation of tests/golden/tf1_emulation.py,
with the oracle's synthetic variables (oracle.p3d_oracle.VarStore(seed=0)) and the synthetic clip synthetic_clip(1, 16, 32, seed=0).

    python tests/golden/make_reference_graph_golden.py            (needs /root/reference; run in the build container)

Per graph and mode the fixture keeps: the output map (fp32, every 2nd pixel of every 2nd frame + sum / sum of squares of the whole
map), the ordered list of variable names the reference code created, and for the training-mode graphs the reference's own
smooth_l1_loss (utils/network.py:49-62) of the output against synthetic_target(seed=1).
tests/test_reference_wiring_cpu.py compares the oracle with these vectors (everywhere) and re-runs the reference live when
/root/reference exists."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import tf1_emulation as E  # noqa: E402
from oracle import p3d_oracle as O  # noqa: E402

REF = "/root/reference"
SIZE, BATCH = 32, 1
GRAPHS = [  # (reference module, builder == oracle graph name, modes)
    ("p3d", "p3d_unetplusplus_ds", (False, True)),
    ("p3d", "p3d_unetplusplus_nonsa", (True,)),
    ("p3d", "p3d_unet", (True,)),
    ("p3d", "p3d_concat", (True,)),
    ("gn", "inference_p3d", (True,)),
    ("gn", "inference_p3d_concat", (True,)),
    ("gn", "inference_p3d_decoder_block", (True,)),
]


def reference_run(module, builder, training, params=None):
    """returns (output [B,16,H,W,1] tensor, created variable names, reference smooth-L1 loss, the variable values used)"""
    x = O.synthetic_clip(BATCH, 16, SIZE, seed=0)
    y = O.synthetic_target(BATCH, 16, SIZE, seed=1)
    if params is None:
        vs = O.VarStore(seed=0)
        with torch.no_grad():
            O.forward(builder, x, vs, training)      # creates the synthetic variables (names are checked against the reference's below)
        params = dict(vs.params)
    out, created = E.run_reference_builder(REF, module, builder, x, params, training, BATCH)
    # the reference's own loss function on its own output (train.py:156-159: reshape to [B,16,H,W], weights 1, sigma 1)
    tf = E.build_module({})
    net = E.load_reference_module(os.path.join(REF, "utils", "network.py"), "utils.network", tf)
    loss = float(net.smooth_l1_loss(E._t(out.reshape(y.shape)), E._t(y), 1, 1, sigma=1.0))
    return out, created, loss, params


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    fix = {}
    for module, builder, modes in GRAPHS:
        for training in modes:
            out, created, loss, _ = reference_run(module, builder, training)
            key = f"{builder}/{'train' if training else 'infer'}"
            o = out.reshape(BATCH, 16, SIZE, SIZE).numpy().astype(np.float32)
            fix[key + "/sample"] = o[:, ::2, ::2, ::2].copy()
            fix[key + "/sums"] = np.array([o.astype(np.float64).sum(), (o.astype(np.float64) ** 2).sum()])
            fix[key + "/loss"] = np.array([loss])
            fix[key + "/variables"] = np.array("\n".join(created))
            print(f"{key}: {len(created)} variables, loss {loss:.6f}, mean {o.mean():.6f}")
    np.savez_compressed(os.path.join(HERE, "reference_graphs_golden.npz"), **fix)
    print("wrote", os.path.join(HERE, "reference_graphs_golden.npz"))


if __name__ == "__main__":
    main()
