"""CPU simulation behind DESIGN.md §2: tile-load balance of minimizer-owned tiles on a uniform random genome for three
strand-symmetric m-mer order hashes: sum = (f + r) * C (degenerate: depends on mirrored-base differences only),
xor = (f*C) ^ (r*C) (0 for every palindromic m-mer), mix = xor + (f & r) * C2 (what the library uses).
Usage: python tools/sim/minimizer_balance.py [n_tiles]   (2 Mb genome; 474 tiles = the C2 ratio of k-mers per tile)"""
import numpy as np, sys
rng=np.random.default_rng(1)
G=2_000_000
g=rng.integers(0,4,size=G,dtype=np.uint64)
k=25
def run(m, ntiles, mode):
    W=k-m+1
    mask=(1<<(2*m))-1
    # forward m-mers
    f=np.zeros(G-m+1,dtype=np.uint64); r=np.zeros(G-m+1,dtype=np.uint64)
    for i in range(m):
        f=(f<<np.uint64(2))|g[i:G-m+1+i]
        r=r|((np.uint64(3)-g[i:G-m+1+i])<<np.uint64(2*i))
    C=np.uint64(0x9E3779B1); sh=np.uint64(32-2*m); M32=np.uint64(0xFFFFFFFF)
    cm=(C<<sh)&M32
    if mode=='sum': y=((f+r)*cm)&M32
    elif mode=='xor': y=((f*cm)&M32)^((r*cm)&M32)
    else:
        c2=(np.uint64(0x85EBCA77)<<sh)&M32
        y=((((f*cm)&M32)^((r*cm)&M32))+(((f&r)*c2)&M32))&M32
    n=G-k+1
    ymin=y[:n].copy()
    for j in range(1,W): ymin=np.minimum(ymin,y[j:j+n])
    h=(ymin*np.uint64(0x9E3779B1))&M32
    h=((h^(h>>np.uint64(15)))*np.uint64(0x85EBCA77))&M32
    h^=h>>np.uint64(13); h&=np.uint64(0xFFFFFFF0)
    tile=(h*np.uint64(ntiles))>>np.uint64(32)
    cnt=np.bincount(tile.astype(np.int64),minlength=ntiles)
    chg=np.count_nonzero(ymin[1:]!=ymin[:-1])
    print(f"m={m} W={W} mode={mode} tiles={ntiles}: mean {cnt.mean():.0f} cv {cnt.std()/cnt.mean():.3f} max/mean {cnt.max()/cnt.mean():.2f} distinct-ymin {np.unique(ymin).size} density {chg/n:.3f}")
nt=int(sys.argv[1]) if len(sys.argv)>1 else 474
for m,mode in [(10,'sum'),(12,'sum'),(10,'xor'),(12,'xor'),(14,'xor'),(10,'mix'),(12,'mix'),(14,'mix')]:
    run(m,nt,mode)
