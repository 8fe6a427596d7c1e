import numpy as np
rng=np.random.default_rng(1)
G=2_000_000; k=32; m=15
g=rng.integers(0,4,G,dtype=np.uint64)
# forward m-mers at each position
def mmers(seq):
    n=len(seq)-m+1
    v=np.zeros(n,np.uint64)
    for j in range(m):
        v=(v<<np.uint64(2))|seq[j:j+n]
    return v
f=mmers(g)
rcg=(np.uint64(3)-g)[::-1]
r=mmers(rcg)[::-1]   # r[i] = revcomp of f[i]
c=np.minimum(f,r).astype(np.uint32)
def mix(c):
    c=(c*np.uint32(0x9E3779B1)); c^=c>>np.uint32(15); c=c*np.uint32(0x85EBCA6B); c^=c>>np.uint32(13); return c
h=mix(c)
W=k-m
nk=G-k+1
# window min over W+1
from numpy.lib.stride_tricks import sliding_window_view
mh=sliding_window_view(h,W+1).min(axis=1)[:nk]
print("distinct minimizers",len(np.unique(mh)),"kmers",nk, "avg run", nk/ (np.count_nonzero(np.diff(mh))+1))
for scale in (1,2,4):
    slots=int(nk*1.2/0.5*scale); lines=slots//16
    home=((mh.astype(np.uint64)*np.uint64(0x9E3779B97F4A7C15))>>np.uint64(32)).astype(np.uint64)
    # umulhi(x, lines): approximate with top 32 bits * lines >> 32
    line=(home*np.uint64(lines))>>np.uint64(32)
    cnt=np.bincount(line.astype(np.int64),minlength=lines)
    print("scale",scale,"lines",lines,"mean",cnt.mean(),"max",cnt.max(),"P(>16)",(cnt>16).mean(),"P(>32)",(cnt>32).mean(), "frac kmers in lines>16", cnt[cnt>16].sum()/nk)
    # linear spill simulation: overflow carried to next line
    carry=0; disp=[]
    occ=np.zeros(lines,np.int64)
    c2=0
    for L in range(lines):
        tot=cnt[L]+c2
        occ[L]=min(tot,16); c2=max(tot-16,0)
    print("   max carry run / frac full lines", (occ==16).mean())
print("---- region simulation (successful/unsuccessful probe counts in buckets of 4 slots)")
def simulate(mh, nk, scale, RB):  # RB buckets per region
    slots=int(nk*1.2/0.5*scale); nb=slots//(4*RB)*(RB); nreg=nb//RB
    home=((mh.astype(np.uint64)*np.uint64(0x9E3779B97F4A7C15))>>np.uint64(32)).astype(np.uint64)
    reg=((home*np.uint64(nreg))>>np.uint64(32)).astype(np.int64)
    keyh=rng.integers(0,RB,nk)   # sub-bucket start
    fill=np.zeros(nb,np.int8)
    probes=np.zeros(nk,np.int32)
    order=rng.permutation(nk)
    for i in order[:400000]:
        b=reg[i]*RB; t=0; p=0
        while True:
            bb=(b+((keyh[i]+t)%RB))%nb
            p+=1
            if fill[bb]<4:
                fill[bb]+=1; break
            t+=1
            if t==RB: t=0; b=(b+RB)%nb
        probes[i]=p
    return probes[order[:400000]]
# subsample: use first 400k kmers of a 2M set means load is partial; instead use smaller genome fully
