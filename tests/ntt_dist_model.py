import random, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.py import curves, ntt as pyntt
c=curves.BLS12_381; r_=c.r

def dft_pass(P, mem_in, mem_out_sel, w, logroot):
    """Model of k_ntt_pass with the generalized params (dict P)."""
    R=1<<P['lr']; TR=[pow(w,e<<(logroot-P['lr']),r_) for e in range(R)]
    for q in range(P['Q']):
        q0=q&(P['Q0']-1); qr=q>>P['lq0']; q1=qr&(P['Q1']-1); q2=qr>>P['lq1']
        qs=(q0,q1,q2)
        x=[]
        for r in range(R):
            v=mem_in[q0*P['in_s0']+q1*P['in_s1']+q2*P['in_s2']+r*P['in_sr']]
            if P['tw_sel']>=0 and r and qs[P['tw_sel']]:
                v=v*pow(w,r*qs[P['tw_sel']]*P['tw_scale'],r_)%r_
            x.append(v)
        lr=P['lr']; blk=R; rem=lr
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
            v=x[p]
            if P.get('otw',0):
                row=qs[P['otw_rsel']]*P['otw_ra']+k*P['otw_rb']; col=P['otw_base']+q0
                e=(row*col)%(1<<logroot)*P['otw_scale']
                v=v*pow(w,e,r_)%r_
            off=q0*P['out_s0']+q1*P['out_s1']+q2*P['out_s2']
            if P.get('peer_k',0):
                h=k//P['peer_k']; mem_out_sel(h)[off+(k%P['peer_k'])*P['out_sr']]=v
            else:
                mem_out_sel(0)[off+k*P['out_sr']]=v

def lg(x):
    l=0
    while (1<<l)<x: l+=1
    return l
MAXLR=9
def split(logl):
    npass=(logl+MAXLR-1)//MAXLR; base=logl//npass; extra=logl%npass
    return [base+1 if i<extra else base for i in range(npass)]

def dist_ntt(a, G):
    N=len(a); logn=lg(N); w=curves.root_of_unity(c,logn)
    l1=logn//2; l2=logn-l1; N1=1<<l1; N2=1<<l2; C=N2//G; T=N1//G
    # slabs
    A=[[a[j1*N2+g*C+cc] for j1 in range(N1) for cc in range(C)] for g in range(G)]
    B=[[None]*(T*N2) for _ in range(G)]
    for g in range(G):
        cur=A[g]; Ns=1; rad=split(l1)
        for pi,lr in enumerate(rad):
            R=1<<lr; lastp=(pi==len(rad)-1)
            out=[None]*(N1*C)
            P=dict(lr=lr,Q=(N1//R)*C,Q0=C,lq0=lg(C),tw_sel=-1,tw_scale=0,in_s0=1,out_s0=1)
            if Ns==1:
                P.update(Q1=1,lq1=0,in_s1=0,in_s2=C,in_sr=(N1//R)*C,out_s1=0,out_s2=R*C,out_sr=C)
            else:
                P.update(Q1=Ns,lq1=lg(Ns),in_s1=C,in_s2=Ns*C,in_sr=(N1//R)*C,out_s1=C,out_s2=Ns*R*C,out_sr=Ns*C,tw_sel=1,tw_scale=N//(Ns*R))
            if lastp:
                # exchange: k1 = q1 + k*Ns ; peer h = k // (R/G) ; local row = (k % (R/G))*Ns + q1
                assert R%G==0
                P.update(otw=1,otw_rsel=1,otw_ra=1,otw_rb=Ns,otw_base=g*C,otw_scale=1,peer_k=R//G,out_s0=1,out_s1=N2,out_s2=0,out_sr=Ns*N2)
                base=g*C
                class View:
                    def __init__(s,arr,base): s.a=arr; s.b=base
                    def __setitem__(s,i,v): s.a[s.b+i]=v
                dft_pass(P,cur,lambda h:View(B[h],base),w,logn)
            else:
                dft_pass(P,cur,lambda h:out,w,logn)
                cur=out
            Ns*=R
    # step 3
    O=[[None]*(N2*T) for _ in range(G)]
    for h in range(G):
        rad=split(l2); cur=B[h]; Ns=1
        for pi,lr in enumerate(rad):
            R=1<<lr; lastp=(pi==len(rad)-1)
            out=[None]*(T*N2)
            if pi==0 and not lastp:
                # rows contiguous in; write transposed I[pos][t]
                P=dict(lr=lr,Q=(N2//R)*T,Q0=N2//R,lq0=lg(N2//R),Q1=T,lq1=lg(T),tw_sel=-1,tw_scale=0,
                       in_s0=1,in_s1=N2,in_s2=0,in_sr=N2//R,out_s0=R*T,out_s1=1,out_s2=0,out_sr=T)
            elif pi==0 and lastp:
                # single pass: rows in, out[k2][t]
                P=dict(lr=lr,Q=T,Q0=T,lq0=lg(T),Q1=1,lq1=0,tw_sel=-1,tw_scale=0,in_s0=N2,in_s1=0,in_s2=0,in_sr=1,out_s0=1,out_s1=0,out_s2=0,out_sr=T)
            else:
                # lanes along t; I[pos][t]; q1 = jb (<Ns), q2 = ja
                P=dict(lr=lr,Q=(N2//R)*T,Q0=T,lq0=lg(T),Q1=Ns,lq1=lg(Ns),tw_sel=1,tw_scale=N//N2*(N2//(Ns*R)),
                       in_s0=1,in_s1=T,in_s2=Ns*T,in_sr=(N2//R)*T,out_s0=1,out_s1=T,out_s2=Ns*R*T,out_sr=Ns*T)
            dft_pass(P,cur,lambda hh:out,w,logn)
            cur=out; Ns*=R
        O[h]=cur
    # gather natural order: X[k1 + N1*k2], GPU h has k1 in [h*T,(h+1)*T), layout O[h][k2*T + t]
    X=[None]*N
    for h in range(G):
        for k2 in range(N2):
            for t in range(T):
                X[(h*T+t)+N1*k2]=O[h][k2*T+t]
    return X
rng=random.Random(2)
import itertools
for MAXLR_,logn,G in ((9,6,2),(9,8,4),(2,8,2),(2,9,2),(3,10,4),(2,11,2),(3,12,4)):
    globals()['MAXLR']=MAXLR_
    a=[rng.randrange(r_) for _ in range(1<<logn)]
    assert dist_ntt(a,G)==pyntt.ntt(c,a),(logn,G)
    print(MAXLR_,logn,G,split(logn//2),split(logn-logn//2),'ok')
