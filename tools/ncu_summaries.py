"""Turns the artefacts of tools/ncu_round.sh (gpurun_out/) into the committed summaries under profiles/.
    python tools/ncu_summaries.py <tag>      e.g. r01_h"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for mode in ("train", "infer"):
    src = os.path.join(G, f"launches_{mode}.csv")
    if os.path.exists(src):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"), src], capture_output=True, text=True).stdout
        open(os.path.join(P, f"{tag}_launches_{mode}_ds_b8.txt"), "w").write(out)
        open(os.path.join(P, f"{tag}_launches_{mode}_ds_b8.csv"), "w").write(open(src).read())
names = {"full_conv_fwd": "dominant conv (x_1_2/x_1_3 3x3x3 128+128->128, B=8) timed alone: tools/run_dominant_kernel.py fwd",
         "full_flash": "attention kernels of the x_1_3_sa block inside the training step (tools/profile_step.py)",
         "full_bn_coop": "single-launch BatchNorm backward (cooperative grid) inside the training step"}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
lines, traffic = [], None
for f, desc in names.items():
    rep = os.path.join(G, f + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines.append(f"== {f}.ncu-rep  ({desc})\n   ncu --set full --clock-control none --import-source on")
    for r in rows[2:]:
        lines.append("  kernel: " + r[hdr.index("Kernel Name")][:100])
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"     {w:95s} {r[i]:>16s} {units[i]}")
        if f == "full_conv_fwd":
            i0, i1 = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            traffic = float(r[i0]) * scale[units[i0]] + float(r[i1]) * scale[units[i1]]
            i2 = hdr.index("l1tex__m_xbar2l1tex_read_bytes.sum")
            l2sm = float(r[i2]) * scale[units[i2]]
open(os.path.join(P, f"{tag}_ncu_full_summary.txt"), "w").write("\n".join(lines) + "\n")
if traffic is not None:
    json.dump({"kernel": "conv_tc_persist_kernel<128,4,2> x_1_2 fwd B=8", "dram_bytes_per_launch": traffic,
               "algorithmic_bytes_per_launch": 2 * 51380224 + 1769472 + 51380224, "l2_to_sm_bytes_per_launch": l2sm,
               "source": f"profiles/{tag}_ncu_full_summary.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"},
              open(os.path.join(P, "dominant_kernel_traffic.json"), "w"), indent=1)
print("\n".join(lines[:24]))
