import csv,re,collections,sys
dis="/tmp/dis_pk.txt"; key=sys.argv[1]; srcf=sys.argv[2]
line_of={}; cur=None; on=False
for l in open(dis):
    if l.startswith("//---") and ".text." in l:
        on = key in l; continue
    if not on: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split("/")[-1],int(m.group(2))); continue
    m=re.match(r"\s*/\*([0-9a-f]{4,})\*/",l)
    if m and cur: line_of[int(m.group(1),16)]=cur
rows=list(csv.reader(open(srcf))); h=rows[1]
ai,si,ii=h.index("Address"),h.index("# Samples"),h.index("Instructions Executed")
cols=["stall_long_sb","stall_mio","stall_short_sb","stall_barrier","stall_wait","stall_lg","stall_membar","stall_sleep"]
ci=[h.index(c) for c in cols]
base=None; agg=collections.defaultdict(lambda:[0,0]+[0]*len(cols))
for r in rows[2:]:
    try: a=int(r[ai],16)
    except ValueError: continue
    if base is None: base=a
    k=line_of.get(a-base,("?",0)); v=agg[k]; v[0]+=int(r[si]); v[1]+=int(r[ii])
    for j,c in enumerate(ci): v[2+j]+=int(r[c])
ts=sum(v[0] for v in agg.values())
src=open("/root/repo/mujoco_rl_manipulate_unknown_objects_b200/csrc/policy_kernels.cu").read().splitlines()
print("samp%  inst | "+" ".join(c[6:] for c in cols))
for (f,n),v in sorted(agg.items(), key=lambda x:-x[1][0])[:int(sys.argv[3]) if len(sys.argv)>3 else 18]:
    t=src[n-1].strip()[:80] if f=="policy_kernels.cu" and 0<n<=len(src) else f
    print("%5.1f %8d | %s | %4d %s"%(100*v[0]/ts, v[1], " ".join("%5d"%x for x in v[2:]), n, t))
