import random, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.py import curves, ntt as pyntt
c=curves.BLS12_381; r_=c.r
def plan(logn):
    if logn==0: return []
    npass=(logn+8)//9; base=logn//npass; extra=logn%npass
    return [base+1 if i<extra else base for i in range(npass)]
def run_pass(inp, L, lr, Ns, logroot, w):
    R=1<<lr; Q=L//R
    out=[None]*L
    if Ns==1:
        Q0,Q1=Q,1; in_s=(1,0,0,L//R); out_s=(R,0,0,1); tw_sel=-1; tw_scale=0
    else:
        Q0,Q1=Ns,L//(R*Ns); in_s=(1,Ns,0,L//R); out_s=(1,Ns*R,0,Ns); tw_sel=0; tw_scale=(1<<logroot)//(Ns*R)
    TR=[pow(w,e<<(logroot-lr),r_) for e in range(R)]
    for q in range(Q):
        q0=q%Q0; qr=q//Q0; q1=qr%Q1; q2=qr//Q1
        x=[]
        for r in range(R):
            v=inp[q0*in_s[0]+q1*in_s[1]+q2*in_s[2]+r*in_s[3]]
            if tw_sel>=0 and r:
                tq=(q0,q1,q2)[tw_sel]
                v=v*pow(w,r*tq*tw_scale,r_)%r_
            x.append(v)
        # DIF rounds
        blk=R; rem=lr
        def rnd(B):
            nonlocal blk
            rho=1<<B; sub=blk//rho; tws=R//blk
            for grp in range(R//rho):
                b0=grp//sub; u=grp%sub; base=b0*blk+u
                xs=[x[base+i*sub] for i in range(rho)]
                span=rho//2
                while span>=1:
                    for h in range(0,rho,2*span):
                        for i in range(span):
                            a=(xs[h+i]+xs[h+i+span])%r_; d=(xs[h+i]-xs[h+i+span])%r_
                            if i: d=d*TR[i*(R//(2*span))]%r_
                            xs[h+i]=a; xs[h+i+span]=d
                    span//=2
                for i in range(rho):
                    m=0
                    for bb in range(B): m|=((i>>bb)&1)<<(B-1-bb)
                    y=xs[i]
                    if m and u and sub>1: y=y*TR[u*m*tws]%r_
                    x[base+m*sub]=y
            blk//=rho
        while rem>=3: rnd(3); rem-=3
        if rem==2: rnd(2)
        if rem==1: rnd(1)
        n8=lr//3; last=lr%3
        for k in range(R):
            p=0; kk=k; sub=R
            for t in range(n8): sub>>=3; p+=(kk&7)*sub; kk>>=3
            if last: sub>>=last; p+=(kk&((1<<last)-1))*sub
            out[q0*out_s[0]+q1*out_s[1]+q2*out_s[2]+k*out_s[3]]=x[p]
    return out
def ntt_model(a):
    L=len(a); logn=L.bit_length()-1
    w=curves.root_of_unity(c,logn)
    Ns=1; cur=list(a)
    for lr in plan(logn):
        cur=run_pass(cur,L,lr,Ns,logn,w); Ns<<=lr
    return cur
rng=random.Random(1)
for logn in (1,2,3,4,5,6,7,9,10,11,13):
    a=[rng.randrange(r_) for _ in range(1<<logn)]
    assert ntt_model(a)==pyntt.ntt(c,a),logn
    print(logn,plan(logn),'ok')
