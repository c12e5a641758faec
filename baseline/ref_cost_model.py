"""Cost model of the REFERENCE call graph: walks zkFC::prove / zkReLU::prove / Commitment::commit
as written in /root/reference and counts kernel launches, cudaMallocs, HBM bytes, Fr muls and G1 scalar-muls.
Analysis aid only (no arithmetic is performed); run: python baseline/ref_cost_model.py"""
from math import ceil, log2
def cl2(n):
    return 0 if n==0 else (n-1).bit_length()
class C:
    def __init__(s): s.k=0; s.m=0; s.bytes=0; s.frmul=0; s.g1mul=0; s.g1add=0; s.sync=0; s.seqg1=0
    def add(s,o):
        for a in vars(s): setattr(s,a,getattr(s,a)+getattr(o,a))
    def __repr__(s): return f"launch={s.k} malloc={s.m} HBM={s.bytes/1e9:.3f}GB Frmul={s.frmul/1e6:.2f}M G1scalarmul={s.g1mul} G1add={s.g1add} seqG1mul_depth={s.seqg1}"
FR=32; G1=144
def fr_sum(c,n):
    c.m+=2; c.bytes+=2*n*FR
    while n>1:
        g=(n+255)//256; c.k+=1; c.bytes+=n*FR+g*FR; n=g
def g1_sum(c,n):
    c.m+=2; c.bytes+=2*n*G1
    while n>1:
        g=(n+63)//64; c.k+=1; c.g1add+=n; c.bytes+=n*G1+g*G1; n=g
def fr_me(c,n,rounds):
    # Fr_me: allocs t_new each level (even at base)
    for r in range(rounds):
        o=(n+1)//2; c.m+=1; c.k+=1; c.bytes+=n*FR+o*FR; c.frmul+=o; n=o
    c.m+=1
def partial_me(c,n,rounds,w):
    for r in range(rounds):
        nw=(n+2*w-1)//(2*w); o=w*nw; c.m+=1; c.k+=1; c.bytes+=n*FR+o*FR; c.frmul+=o; n=o
    c.m+=1; c.bytes+=2*n*FR
    return n
def ip_sc(c,n,rounds):
    for r in range(rounds):
        o=(n+1)//2; c.m+=3; c.k+=1; c.bytes+=2*n*FR+3*o*FR; c.frmul+=4*o
        for _ in range(3): fr_sum(c,o)
        c.m+=2; c.k+=2; c.bytes+=2*(n*FR+o*FR); c.frmul+=2*o; n=o
def hp_sc(c,n,rounds):
    for r in range(rounds):
        o=(n+1)//2; c.m+=3; c.k+=1; c.bytes+=2*n*FR+3*o*FR; c.frmul+=4*o
        for _ in range(3): fr_me(c,o,rounds-r-1)
        c.m+=2; c.k+=2; c.bytes+=2*(n*FR+o*FR); c.frmul+=2*o; n=o
def bin_sc(c,n,rounds):
    for r in range(rounds):
        o=(n+1)//2; c.m+=3; c.k+=1; c.bytes+=n*FR+3*o*FR; c.frmul+=3*o
        for _ in range(3): fr_me(c,o,rounds-r-1)
        c.m+=1; c.k+=1; c.bytes+=(n*FR+o*FR); c.frmul+=o; n=o
def g1_me(c,n,rounds):
    for r in range(rounds):
        o=(n+1)//2; c.m+=1; c.k+=1; c.g1mul+=o; c.g1add+=2*o; c.bytes+=n*G1+o*G1; c.seqg1+=1; n=o
    c.m+=1
def me_open(c,n,rounds):
    for r in range(rounds):
        o=n//2; c.m+=5; c.k+=1; c.g1mul+=5*o; c.g1add+=3*o; c.frmul+=2*o; c.seqg1+=5
        c.bytes+=n*(FR+G1)+o*(FR+4*G1)
        for _ in range(3): g1_sum(c,o)
        n=o
def fc_prove(B,I,O,gsize):
    c=C()
    partial_me(c,B*I,cl2(B),I)
    partial_me(c,I*O,cl2(O),1)
    ip_sc(c,I,cl2(I))
    fr_me(c,B*O,cl2(B)+cl2(O))
    m=I*O//gsize
    g1_me(c,m,cl2(m))
    partial_me(c,I*O,cl2(m),gsize)
    me_open(c,gsize,cl2(gsize))
    return c
def relu_prove(n):
    c=C(); L=cl2(n)
    bin_sc(c,n*32,L+5); partial_me(c,n*32,L,32)
    bin_sc(c,n*16,L+4); partial_me(c,n*16,L,16)
    hp_sc(c,n,L)
    return c
def commit(I,O,gsize):
    c=C(); n=I*O; c.m+=3; c.k+=3; c.g1mul+=n; c.bytes+=n*(2*FR+G1)+n*G1; c.seqg1=1
    return c
def mlp(dims,B):
    r=lambda x:1<<cl2(x)
    Br=r(B); tot=C(); totc=C(); params=0
    for l in range(len(dims)-1):
        i,o=dims[l],dims[l+1]; params+=i*o
        gs=1<<((cl2(i*o)+1)//2)
        I,O=r(i),r(o)
        f=fc_prove(Br,I,O,gs); tot.add(f)
        totc.add(commit(I,O,gs))
        if l<len(dims)-2:
            rl=relu_prove(Br*O); tot.add(rl)
        print(f"  L{l}: {i}x{o} -> {I}x{O} gens={gs} com.size={I*O//gs} | FC {f}")
        if l<len(dims)-2: print(f"       ReLU n={Br*O}: {rl}  aux bytes={(Br*O*49*32)/1e9:.2f}GB")
    print(" params",params); print(" PROVE TOTAL",tot); print(" COMMIT(setup, untimed)",totc)
print("== demo MLP batch 256"); mlp([784,1000,1773,1773,1773,1773,1773,1124,1000],256)
print("== demo MLP batch 1"); mlp([784,1000,1773,1773,1773,1773,1773,1124,1000],1)
print("== 4096x4096 b256 single FC"); print(fc_prove(256,4096,4096,4096)); print(commit(4096,4096,4096))
print("== deep narrow 32x1024 batch 1"); mlp([1024]*33,1)
print("== deep narrow 32x1024 batch 4096"); mlp([1024]*33,4096)
