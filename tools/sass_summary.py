"""profiles/<tag>_sass_summary.txt: per kernel of lib/libsap3d_b200.so the counts of the SASS mnemonics that prove the Blackwell
paths (cuobjdump -sass): UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA tensor loads), UBLKCP (bulk copies: DSMEM
exchange), UTCBAR (tcgen05.commit), SYNCS (mbarrier), ACQBULK / griddepcontrol (programmatic dependent launch), BAR / cluster ops.
    python tools/sass_summary.py r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
lib = os.path.join(ROOT, "sap3d_tensorflow_b200", "lib", "libsap3d_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "UCGABAR", "ACQBULK", "PREEXIT", "LDGSTS", "HMMA", "WARPSYNC"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_n"] += 1
        for w in WANT:
            if op.startswith(w):
                key = w
                if w == "UTMALDG":
                    dims = re.search(r"UTMALDG\.(\dD)", op)
                    key = "UTMALDG." + (dims.group(1) if dims else "?") + (".MULTICAST" if "MULTICAST" in op else "")
                if w == "UTCHMMA" and "2CTA" in op:
                    key = "UTCHMMA.2CTA"
                per[cur][key] += 1
demangled = subprocess.run(["cu++filt"] + list(per), capture_output=True, text=True).stdout.splitlines() if per else []
out = [f"cuobjdump -sass {os.path.relpath(lib, ROOT)}  ({len(per)} kernels; sm_100a)", ""]
tot = collections.Counter()
for (name, c), dm in zip(per.items(), demangled if len(demangled) == len(per) else list(per)):
    keys = [k for k in c if k != "_n"]
    for k in keys:
        tot[k] += c[k]
    if not keys:
        continue
    short = re.sub(r"\(anonymous namespace\)::|sap3d::|void |\(int\)|<unnamed>::", "", dm)
    short = re.split(r"\((?!anonymous)", short)[0]
    out.append(f"{short[:78]:78s} {c['_n']:6d} instr  " + "  ".join(f"{k}={c[k]}" for k in sorted(keys)))
out.append("")
out.append("totals: " + "  ".join(f"{k}={v}" for k, v in sorted(tot.items())))
path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out[-12:]))
print("wrote", path)
