#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
SAP3D_CONV_TRACE=1 timeout 300 python tools/profile_step.py p3d_unetplusplus_ds 8 112 train > $O/c30_trace.out 2> $O/c30_conv_trace.err; echo "rc=$?"
python - <<'PY'
import re,collections
L=[l for l in open('gpurun_out/c30_conv_trace.err') if l.startswith('[conv_trace]')]
n=len(L)//3   # three steps were run: take the last third
L=L[-n:]
agg=collections.Counter()
for l in L:
    m=re.search(r'ext=(\S+) cout=(\d+) ncls=(\d+) taps=(\d+) nkb=(\d+) views=(\d+) box=(\S+) m_tiles=(\d+) block_n=(\d+) mt=(\d+) grid=(\d+) split=(\d+) (\w+)',l)
    agg[m.groups()]+=1
print(len(L),'launches per step')
for k,c in sorted(agg.items(), key=lambda kv:(-int(kv[0][7])*int(kv[0][4]), kv[0])):
    ext,cout,ncls,taps,nkb,views,box,mt_,bn,mt,grid,split,kind=k
    if kind in ('persist','halo','swap'):
        print(f"{c:3d}x {kind:8s} ext={ext:16s} cout={cout:4s} ncls={ncls} taps={taps:3s} nkb={nkb:4s} box={box:10s} tiles={mt_:6s} bn={bn} mt={mt} grid={grid}")
PY
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
