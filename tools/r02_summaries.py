"""gpurun_out/r02_*.ncu-rep + launch lists (tools/r02_evidence.sh) -> committed summaries under profiles/:
   r02_launches_{train,infer}_ds_b8.{csv,txt}, r02_ncu_full_summary.txt, r02_ncu_bandwidth_summary.txt, dominant_kernel_traffic.json
   python tools/r02_summaries.py"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = "r02"
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
tscale = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}

for mode in ("train", "infer"):
    src = os.path.join(G, f"{tag}_launches_{mode}.csv")
    if os.path.exists(src):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"), src], capture_output=True, text=True).stdout
        open(os.path.join(P, f"{tag}_launches_{mode}_ds_b8.txt"), "w").write(out)
        open(os.path.join(P, f"{tag}_launches_{mode}_ds_b8.csv"), "w").write(open(src).read())


def rows_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    return rows[0], rows[1], rows[2:]


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|sap3d::|void |<unnamed>::", "", name)
    return re.split(r"\((?!int)", name)[0].replace("(int)", "")


FULL = [
    ("r02_full_conv_dominant", "FLOP-dominant conv, timed alone (tools/run_dominant_kernel.py fwd): x_1_2 / x_1_3 forward, 3x3x3, 128+128 -> 128 channels at "
                               "8 x 8 x 56 x 56 positions (M = 200704, K = 6912, N = 128; 355.1 GFLOP, 156 MB algorithmic)"),
    ("r02_full_wgrad_dominant", "filter gradient of the same layer, timed alone (tools/run_dominant_kernel.py wgrad): dW[27][256][128] += P^T Q over 200704 "
                                "positions, per 128-channel input segment (2 launches; 355.1 GFLOP together; algorithmic DRAM per launch: x segment 51.4 MB + dy 51.4 MB)"),
    ("r02_full_conv_halo", "the same launch on the un-swapped halo-tile kernel (SAP3D_CONV_SWAP=0 form, captured before the swap existed): "
                           "positions on the M side, three taps per activation box"),
    ("r02_full_wgrad_pair", "opt-in two-tap filter-gradient kernel on the same layer (SAP3D_WGRAD_PAIR=1; measured slower, DESIGN 3.3)"),
    ("r02_full_conv_splitk", "split-K cluster conv inside the training step (stage-3 backbone layer, 784 positions)"),
    ("r02_full_bn_slab", "BatchNorm backward of a stage-3 tensor (784 positions) inside the training step: one block per 8 channels, no grid barrier"),
    ("r02_full_bn_nob", "BatchNorm backward of a decoder tensor (8 x 8 x 56 x 56 x 128 = 25.7 M elements) inside the training step: reduce, then apply"),
    ("r02_full_metrics", "saliency metrics CC / SIM / NSS / KLdiv over 8192 map triples of 112 x 112 (tools/profile_metrics.py)"),
]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum"]
dom_kernel = ""
lines, traffic, l2sm = [f"ncu --set full --clock-control none --import-source on (cold caches, serialised launches); measured peaks: {peaks['hbm_gbs']} GB/s HBM, "
                        f"{peaks['bf16_tflops']} / {peaks['bf16_tflops_sustained']} TFLOP/s bf16 burst / sustained (MEASURED_PEAKS.json)", ""], None, None
for f, desc in FULL:
    rep = os.path.join(G, f + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    hdr, units, rows = rows_of(rep)
    lines.append(f"== {f}.ncu-rep\n   {desc}")
    for r in rows:
        lines.append("  kernel: " + short(r[hdr.index("Kernel Name")])[:110])
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"     {w:95s} {r[i]:>16s} {units[i]}")
        i0, i1, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        b = float(r[i0]) * scale[units[i0]] + float(r[i1]) * scale[units[i1]]
        t = float(r[it]) * tscale[units[it]]
        lines.append(f"     -> DRAM traffic {b / 1e6:.1f} MB in {t * 1e6:.1f} us = {b / t / 1e9:.0f} GB/s = {b / t / 1e9 / peaks['hbm_gbs']:.2f} of the measured HBM peak")
        if f == "r02_full_conv_dominant":
            traffic = b
            dom_kernel = short(r[hdr.index("Kernel Name")])[:80]
            i2 = hdr.index("l1tex__m_xbar2l1tex_read_bytes.sum")
            l2sm = float(r[i2]) * scale[units[i2]]
            lines.append(f"     -> 355.14 GFLOP / {t * 1e6:.1f} us = {355.14e9 / t / 1e12:.0f} TFLOP/s under ncu (the bench's CUDA-event timing is the reported one)")
    lines.append("")
open(os.path.join(P, f"{tag}_ncu_full_summary.txt"), "w").write("\n".join(lines) + "\n")
if traffic is not None:
    json.dump({"kernel": dom_kernel + " x_1_2 fwd B=8", "dram_bytes_per_launch": traffic,
               "algorithmic_bytes_per_launch": 2 * 51380224 + 1769472 + 51380224, "l2_to_sm_bytes_per_launch": l2sm,
               "source": f"profiles/{tag}_ncu_full_summary.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"},
              open(os.path.join(P, "dominant_kernel_traffic.json"), "w"), indent=1)

SOL = [
    ("r02_sol_step_kernels", "bandwidth / latency kernels inside the p3d_unetplusplus_ds training step (B = 8, 112 px), first launches of each"),
    ("r02_sol_gn_cbam", "GroupNorm + CBAM kernels of gn/inference_p3d at configs[2] size (B = 16, 160 px), first launches (stem, first bottleneck)"),
    ("r02_full_bn_nob", "decoder BatchNorm backward (reduce, apply) on 25.7 M-element tensors"),
    ("r02_full_bn_slab", "stage-3 BatchNorm backward"),
    ("r02_full_metrics", "saliency metrics over 8192 maps of 112 x 112"),
]
out = [f"measured HBM peak (MEASURED_PEAKS.json, device copy): {peaks['hbm_gbs']} GB/s; ncu --clock-control none (cold caches, serialised launches);",
       "DRAM MB = dram__bytes_read.sum + dram__bytes_write.sum of the launch", ""]
for f, d in SOL:
    rep = os.path.join(G, f + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    hdr, units, rows = rows_of(rep)
    col = {n: hdr.index(n) for n in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
                                      "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size",
                                      "sm__warps_active.avg.pct_of_peak_sustained_active") if n in hdr}
    rate = {"Gbyte/s": 1e9, "Tbyte/s": 1e12, "Mbyte/s": 1e6, "Kbyte/s": 1e3, "byte/s": 1.0, "Gbyte/second": 1e9, "Tbyte/second": 1e12, "Mbyte/second": 1e6,
            "Kbyte/second": 1e3, "byte/second": 1.0}
    out.append(f"== {f}.ncu-rep  ({d})")
    out.append(f"   {'kernel':52s} {'grid':>7s} {'us':>9s} {'DRAM MB':>9s} {'GB/s':>8s} {'of peak':>8s} {'ncu dram%':>9s} {'occ%':>6s}")
    for r in rows:
        t = float(r[col["gpu__time_duration.sum"]]) * tscale[units[col["gpu__time_duration.sum"]]]
        if "dram__bytes_read.sum" in col:
            b = sum(float(r[col[k]]) * scale[units[col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        else:   # section captures carry the rate, not the byte counts
            b = float(r[col["dram__bytes.sum.per_second"]]) * rate[units[col["dram__bytes.sum.per_second"]]] * t
        out.append(f"   {short(r[col['Kernel Name']])[:52]:52s} {r[col['launch__grid_size']]:>7s} {t * 1e6:9.1f} {b / 1e6:9.1f} {b / t / 1e9:8.0f} "
                   f"{b / t / 1e9 / peaks['hbm_gbs']:8.2f} {float(r[col['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']]):9.1f} "
                   f"{float(r[col['sm__warps_active.avg.pct_of_peak_sustained_active']]):6.1f}")
    out.append("")
open(os.path.join(P, f"{tag}_ncu_bandwidth_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(lines[:60]))
print("\n".join(out))
