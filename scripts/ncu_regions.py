#!/usr/bin/env python
"""Where does a kernel stall?  Splits the SASS of one profiled kernel at its barriers / mbarrier waits and sums the
warp-state samples of every region (needs an ncu report captured with --set full --import-source on).

    python scripts/ncu_regions.py gpurun_out/prof_heads_r1f.ncu-rep heads_bwd_color
"""
import csv, subprocess, sys
rep, kn = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+kn,"--launch-count","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[1]
ia=hdr.index("Source"); isamp=hdr.index("# Samples"); iex=hdr.index("Instructions Executed")
data=[]
for r in rows[2:]:
    try: data.append((int(r[isamp]), int(r[iex]), r[ia]))
    except Exception: pass
tot=sum(d[0] for d in data); print(kn, "total samples", tot, "instr", sum(d[1] for d in data))
marks=[i for i,d in enumerate(data) if ("BAR.SYNC" in d[2] or "UTCBAR" in d[2] or "SYNCS.PHASECHK" in d[2])]
prev=0
for m in marks+[len(data)-1]:
    s=sum(d[0] for d in data[prev:m+1]); ins=sum(d[1] for d in data[prev:m+1])
    if s>0.03*tot:
        top=sorted(range(prev,m+1), key=lambda i:-data[i][0])[:2]
        print(f"{prev:5d}-{m:5d} {100*s/tot:5.1f}% instr {ins:9d} end: {data[m][2][:38]:38s} | hot: " + " ; ".join(f"{data[i][0]}:{data[i][2][:34]}" for i in top))
    prev=m+1
