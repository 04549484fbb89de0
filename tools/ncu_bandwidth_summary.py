"""gpurun_out/full_bw_*.ncu-rep (tools/ncu_bandwidth.sh) -> profiles/<tag>_ncu_bandwidth_summary.txt: per launch the DRAM bytes,
duration, achieved GB/s and the fraction of the measured HBM peak (MEASURED_PEAKS.json).   python tools/ncu_bandwidth_summary.py r01_k"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
tscale = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
desc = {"full_bw_norm": "BatchNorm apply / backward kernels of the decoder inside the p3d_unetplusplus_ds training step (B=8, 112 px)",
        "full_bw_gn": "GroupNorm statistics + CBAM forward kernels of gn/inference_p3d (B=16, 160 px), stem and first bottleneck",
        "full_bw_gn_bwd": "GroupNorm / CBAM backward kernels of gn/inference_p3d (B=16, 160 px)",
        "full_bw_metrics": "saliency metrics: CC/SIM/NSS/KLdiv over 8192 maps 112x112; resize to 1080x960 + AUC over 8 maps"}
out = [f"measured HBM peak (MEASURED_PEAKS.json, device copy): {peak} GB/s; ncu --set full --clock-control none (cold caches, serialised)", ""]
for f, d in desc.items():
    rep = os.path.join(G, f + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {n: hdr.index(n) for n in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                      "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size",
                                      "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct")}
    out.append(f"== {f}.ncu-rep  ({d})")
    out.append(f"   {'kernel':48s} {'grid':>7s} {'us':>9s} {'DRAM MB':>9s} {'GB/s':>8s} {'of peak':>8s} {'ncu dram%':>9s} {'occ%':>6s} {'L2 hit%':>8s}")
    for r in rows[2:]:
        t = float(r[col["gpu__time_duration.sum"]]) * tscale[units[col["gpu__time_duration.sum"]]]
        b = sum(float(r[col[k]]) * scale[units[col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        name = r[col["Kernel Name"]].replace("(anonymous namespace)::", "").replace("sap3d::", "")
        name = name.split("(")[0][:48]
        out.append(f"   {name:48s} {r[col['launch__grid_size']]:>7s} {t * 1e6:9.1f} {b / 1e6:9.1f} {b / t / 1e9:8.0f} {b / t / 1e9 / peak:8.2f} "
                   f"{float(r[col['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']]):9.1f} "
                   f"{float(r[col['sm__warps_active.avg.pct_of_peak_sustained_active']]):6.1f} {float(r[col['lts__t_sector_hit_rate.pct']]):8.1f}")
    out.append("")
open(os.path.join(P, f"{tag}_ncu_bandwidth_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
