import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
cur=None; data=[]
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0] in("Function Name","Line No"): 
        if r[0]=="Line No": hdr=r
        continue
    if r[0]!="" and r[0].isdigit():
        try: inst=int(r[7]); samp=int(r[4])
        except: continue
        data.append((cur,int(r[0]),inst,samp,r[1].strip()[:110]))
tot=sum(d[2] for d in data); ts=sum(d[3] for d in data)
print("total inst",tot,"samples",ts)
byfile=collections.Counter()
for d in data: byfile[d[0]]+=d[2]
print(byfile)
# ranges
import bisect
def rng(f,a,b): 
    i=sum(d[2] for d in data if d[0]==f and a<=d[1]<=b); s=sum(d[3] for d in data if d[0]==f and a<=d[1]<=b)
    return i/tot*100, s/ts*100
for name,f,a,b in [("place_rect_r","ba_kernel.cuh",118,341),("place_rect+borders helpers","ba_kernel.cuh",342,424),("trace etc","ba_kernel.cuh",445,575),("init/grow/elig","ba_kernel.cuh",576,660),("run_generic","ba_kernel.cuh",661,852),("finish","ba_kernel.cuh",853,887),("fast load/spill","ba_kernel.cuh",902,932),("pk_fast_step","ba_kernel.cuh",933,1100),("warp_main","ba_kernel.cuh",1101,1249),("pk scorer","ba_packed.cuh",1,100),("pk_cols8","ba_packed.cuh",101,222),("pk_lane_key..","ba_packed.cuh",223,290),("pk_rect_ok","ba_packed.cuh",291,312),("place_rect_pk","ba_packed.cuh",313,345)]:
    print("%-28s inst %5.1f%%  samples %5.1f%%"%((name,)+rng(f,a,b)))
print()
for d in sorted(data,key=lambda d:-d[2])[:45]: print("%-14s %5d %5.2f%% s%5.2f%% %s"%(d[0],d[1],d[2]/tot*100,d[3]/ts*100,d[4]))
