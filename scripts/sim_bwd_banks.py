#!/usr/bin/env python
"""CPU simulation of the staged warp backward's shared-memory gathers (csrc/warp_bwd_staged.cu):
builds the per-tile transposed-warp CSR for tiles of the synthetic 1080p flow, assigns 16
consecutive pairs per thread and counts the bank wavefronts of every warp-wide gather under
different layouts of the grad_out tile.  Reproduces the measured 3.0 wavefronts per LDS (ncu,
profiles/r01_warp_bwd_staged_ncu_raw.txt) and shows that column permutations, odd pitches, XOR
swizzles and half-row rotations all stay at ~3.0: the lanes' source pixels are ~4.6 apart with
jitter, i.e. effectively random banks (32 balls in 32 bins).  Run: python scripts/sim_bwd_banks.py
"""
import os, sys
import numpy as np
import torch  # noqa: F401
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsvc_b200 import synthetic
H,W=1088,1920
inp=synthetic.make_pframe_inputs(B=1,H=H,W=W,seed=16)
f=inp['flow'][0].numpy()
xs=np.clip(np.arange(W)[None,:]+f[0],0,W-1); ys=np.clip(np.arange(H)[:,None]+f[1],0,H-1)
x0=np.floor(xs).astype(int); y0=np.floor(ys).astype(int)
rng=np.random.default_rng(0)
def wavefronts(banks, addrs):
    # banks, addrs: arrays for 32 lanes (active only); wavefronts = max over banks of distinct addresses
    best=0
    for b in np.unique(banks):
        best=max(best,len(np.unique(addrs[banks==b])))
    return best
layouts={
 'natural': lambda x,y: y*64+x,
 'perm5': lambda x,y: y*64+((x*5)%64),
 'perm3': lambda x,y: y*64+((x*3)%64),
 'perm7': lambda x,y: y*64+((x*7)%64),
 'perm9': lambda x,y: y*64+((x*9)%64),
 'perm11': lambda x,y: y*64+((x*11)%64),
 'perm13': lambda x,y: y*64+((x*13)%64),
 'pitch65': lambda x,y: y*65+x,
 'xorrow': lambda x,y: y*64+(x^((y&7)<<2)),
}
tot={k:0 for k in layouts}; n=0
tiles=[(ty,tx) for ty in range(0,H,16) for tx in range(0,W,64)]
for (ty,tx) in [tiles[i] for i in rng.choice(len(tiles),60,replace=False)]:
    X=x0[ty:ty+16,tx:tx+64]; Y=y0[ty:ty+16,tx:tx+64]
    bx0=X.min()&~3; by0=Y.min()
    pairs=[]
    for yy in range(16):
        for xx in range(64):
            for dy in (0,1):
                for dx in (0,1):
                    ex=min(X[yy,xx]+dx,W-1)-bx0; ey=min(Y[yy,xx]+dy,H-1)-by0
                    if (dx and X[yy,xx]+1>W-1) or (dy and Y[yy,xx]+1>H-1): continue
                    pairs.append((ey*96+ex, xx, yy))
    pairs.sort(key=lambda p:(p[0], rng.random()))
    P=np.array(pairs)
    T=len(P)
    for w in range(8):
        for j in range(16):
            idx=np.array([16*(w*32+l)+j for l in range(32)])
            idx=idx[idx<T]
            if len(idx)==0: continue
            x=P[idx,1]; y=P[idx,2]
            for k,fn in layouts.items():
                a=fn(x,y); tot[k]+=wavefronts(a%32,a)
            n+=1
for k in layouts: print(k, round(tot[k]/n,2))
print("search")
def evaluate(fn, ntile=30):
    rng=np.random.default_rng(1); tot=0;n=0
    for (ty,tx) in [tiles[i] for i in rng.choice(len(tiles),ntile,replace=False)]:
        X=x0[ty:ty+16,tx:tx+64]; Y=y0[ty:ty+16,tx:tx+64]
        bx0=X.min()&~3; by0=Y.min()
        yy,xx=np.meshgrid(np.arange(16),np.arange(64),indexing='ij')
        P=[]
        for dy in (0,1):
            for dx in (0,1):
                ex=np.minimum(X+dx,W-1)-bx0; ey=np.minimum(Y+dy,H-1)-by0
                P.append(np.stack([(ey*96+ex).ravel(), xx.ravel(), yy.ravel()],1))
        P=np.concatenate(P); P=P[np.lexsort((rng.random(len(P)),P[:,0]))]
        T=len(P)
        for w in range(8):
            for j in range(16):
                idx=16*(w*32+np.arange(32))+j; idx=idx[idx<T]
                a=fn(P[idx,1],P[idx,2]); b=a%32
                tot+=max(len(np.unique(a[b==bb])) for bb in np.unique(b)); n+=1
    return tot/n
res=[]
for Pp in (64,65,66,68,72,80):
    for K in (0,1,2,3,4,5,6,8,12,16):
        if 63+K>=Pp and Pp!=64: pass
        res.append((evaluate(lambda x,y:y*Pp+x+K*(x>>5)),Pp,K))
res.sort(); print(res[:12])
