import numpy as np
exec(open('/tmp/sim.py').read().split('print("---- region')[0].replace('G=2_000_000','G=300_000'))
def run(scale,RB):
    slots=int(nk*1.2/0.5*scale); nreg=max(slots//(4*RB),1); nb=nreg*RB
    home=((mh.astype(np.uint64)*np.uint64(0x9E3779B97F4A7C15))>>np.uint64(32)).astype(np.uint64)
    reg=((home*np.uint64(nreg))>>np.uint64(32)).astype(np.int64)
    keyh=rng.integers(0,RB,nk)
    fill=np.zeros(nb,np.int8)
    tot=0
    for i in rng.permutation(nk):
        b=reg[i]*RB; t=0; p=0
        while True:
            bb=(b+((keyh[i]+t)%RB))%nb
            p+=1
            if fill[bb]<4:
                fill[bb]+=1; break
            t+=1
            if t==RB: t=0; b=(b+RB)%nb
        tot+=p
    # unsuccessful search cost from random (region of a random kmer, random start): count buckets until a non-full bucket
    un=0; N=100000
    idx=rng.integers(0,nk,N); ks=rng.integers(0,RB,N)
    for i,kh in zip(idx,ks):
        b=reg[i]*RB; t=0; p=0
        while True:
            bb=(b+((kh+t)%RB))%nb
            p+=1
            if fill[bb]<4: break
            t+=1
            if t==RB: t=0; b=(b+RB)%nb
        un+=p
    print("scale",scale,"RB",RB,"load",nk/(nb*4),"insert(avg probes at insert)",tot/nk,"unsuccessful",un/N)
for scale in (1,2):
    for RB in (1,4,8,16):
        run(scale,RB)
