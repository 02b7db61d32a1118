#!/usr/bin/env python3
"""Hot code footprint of a kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view): how many KB of 128-byte
instruction lines cover 90 / 95 / 99 / 99.9 % of the executed warp instructions. usage: tools/ncu_hot_footprint.py src.csv [...]"""
import csv,sys
for f in sys.argv[1:]:
    rows=list(csv.reader(open(f)))
    ins=[r for r in rows[2:] if len(r)>6 and r[0].startswith("0x")]
    cnt=[int(r[5]) for r in ins]
    tot=sum(cnt); n=len(cnt)
    # 128B lines = 8 instrs
    lines=[sum(cnt[i:i+8]) for i in range(0,n,8)]
    srt=sorted(lines,reverse=True)
    acc=0; 
    out=[]
    for k,v in enumerate(srt):
        acc+=v
        for th in (0.9,0.95,0.99,0.999):
            if acc>=th*tot and not any(o[0]==th for o in out): out.append((th,(k+1)*128/1024))
    print(f, "instrs",n, "KB",n*16/1024, "coverage->KB", out, "lines touched >0:", sum(1 for v in lines if v>0)*128/1024)
    # per-line executions relative: lines executed at least once per pair-step...
