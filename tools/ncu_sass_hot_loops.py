import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
ins=[]
for r in rows[2:]:
    if len(r)>6 and r[0].startswith("0x"):
        ins.append((int(r[0],16),r[1].strip(),int(r[5]),int(r[2])))
base=ins[0][0]
tot=sum(i[2] for i in ins); print("n sass",len(ins),"total",tot)
# cluster contiguous runs with same count
runs=[]
s=0
for k in range(1,len(ins)+1):
    if k==len(ins) or abs(ins[k][2]-ins[s][2])>0.02*max(ins[s][2],1):
        runs.append((s,k-1,ins[s][2])); s=k
big=sorted(runs,key=lambda r:-(r[1]-r[0]+1)*r[2])[:40]
for a,b,c in sorted(big):
    n=b-a+1
    print("idx %5d-%5d n=%4d count=%10d share=%5.2f%% samples=%5.2f%%"%(a,b,n,c,n*c/tot*100, sum(i[3] for i in ins[a:b+1])/sum(i[3] for i in ins)*100))
