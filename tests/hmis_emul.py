"""Plain-Python replay of what the DEVICE does for HMIS (hypre_ve_b200/csrc/b200_hmis.cu + the CF_init 1 branch of the PMIS
kernels in b200_setup.cu): the sequential Ruge-Stueben first pass over per-measure FIFO buckets, then PMIS with Jacobi-style
(read old / write new) sweeps and the special first sweep that honours the reference's in-place update order.  Test
infrastructure: lets the algorithm be compared with the reference's CF markers on a machine without a GPU."""
import numpy as np


def transpose(I,J,n):
    cnt=np.zeros(n+1,int)
    for c in J: cnt[c+1]+=1
    TI=np.cumsum(cnt); nxt=TI.copy(); TJ=np.zeros(len(J),int)
    for i in range(n):
        for k in range(I[i],I[i+1]):
            TJ[nxt[J[k]]]=i; nxt[J[k]]+=1
    return TI,TJ

class L:
    def __init__(s,n,nb): s.head=[-1]*nb; s.tail=[-1]*nb; s.next=[-1]*n; s.prev=[-1]*n; s.maxm=0
    def enter(s,m,i):
        s.next[i]=-1; s.prev[i]=s.tail[m]
        if s.tail[m]>=0: s.next[s.tail[m]]=i
        else: s.head[m]=i
        s.tail[m]=i
        if m>s.maxm: s.maxm=m
    def remove(s,m,i):
        p,q=s.prev[i],s.next[i]
        if p>=0: s.next[p]=q
        else: s.head[m]=q
        if q>=0: s.prev[q]=p
        else: s.tail[m]=p
        while s.maxm>0 and s.head[s.maxm]<0: s.maxm-=1

def ruge(I,J,n):
    TI,TJ=transpose(I,J,n)
    meas=[int(TI[j+1]-TI[j]) for j in range(n)]; cf=[0]*n; left=0
    nb=2*max(meas+[0])+4; Ls=L(n,nb)
    for j in range(n):
        if I[j+1]-I[j]==0: cf[j]=-3; meas[j]=0
        else: left+=1
    for j in range(n):
        if cf[j]==-3: continue
        if meas[j]>0: Ls.enter(meas[j],j); continue
        cf[j]=-2
        for k in range(I[j],I[j+1]):
            nb_=J[k]
            if cf[nb_]==-3: continue
            if nb_<j:
                if meas[nb_]>0: Ls.remove(meas[nb_],nb_)
                meas[nb_]+=1; Ls.enter(meas[nb_],nb_)
            else: meas[nb_]+=1
        left-=1
    while left>0:
        idx=Ls.head[Ls.maxm]; m=meas[idx]; cf[idx]=1; meas[idx]=0; left-=1; Ls.remove(m,idx)
        for j in range(TI[idx],TI[idx+1]):
            nb_=TJ[j]
            if cf[nb_]!=0: continue
            cf[nb_]=-1; Ls.remove(meas[nb_],nb_); left-=1
            for k in range(I[nb_],I[nb_+1]):
                n2=J[k]
                if cf[n2]!=0: continue
                Ls.remove(meas[n2],n2); meas[n2]+=1; Ls.enter(meas[n2],n2)
        for j in range(I[idx],I[idx+1]):
            nb_=J[j]
            if cf[nb_]!=0: continue
            m2=meas[nb_]; Ls.remove(m2,nb_); m2-=1; meas[nb_]=m2
            if m2>0: Ls.enter(m2,nb_); continue
            cf[nb_]=-1; left-=1
            for k in range(I[nb_],I[nb_+1]):
                n2=J[k]
                if cf[n2]!=0: continue
                Ls.remove(meas[n2],n2); meas[n2]+=1; Ls.enter(meas[n2],n2)
    return np.array(cf)

def rand_seq(n):
    seed=2747; out=[]
    for _ in range(n):
        high,low=seed//127773,seed%127773; t=16807*low-2836*high
        seed=t if t>0 else t+2147483647; out.append(seed/2147483647)
    return np.array(out)

def pmis_device_style(I,J,n,cf):
    colcnt=np.bincount(J,minlength=n).astype(float); m=colcnt+rand_seq(n)
    cf=cf.copy()
    rowlen=np.diff(I)
    # init kernel (cf_init 1)
    for i in range(n):
        c=cf[i]
        if c==-3: m[i]=0
        else:
            if c==-1: c=0
            if c==-2: c=0 if (m[i]>=1.0 or rowlen[i]>0) else -1
        cf[i]=c
    demoted=0
    # first sweep kernel (Jacobi style)
    cin=cf.copy(); cout=cf.copy()
    for i in range(n):
        c=cin[i]
        if c in (0,1):
            if m[i]<1:
                if c==1: demoted+=1
                c=-1
            if c>0: c=1
            else:
                for k in range(I[i],I[i+1]):
                    j=J[k]
                    if cin[j]>0 and not (m[j]<1 and j<i): c=-1
        cout[i]=c
    for i in range(n):
        if cin[i] in (0,1) and cout[i]!=0: m[i]=0
    cf=cout
    it=1
    while True:
        ing=(cf==0)
        if not ing.any(): break
        # mark
        for i in range(n):
            if cf[i]==0 and m[i]>1: cf[i]=1
        # remove (racy in CUDA but order independent)
        for i in range(n):
            if ing[i] and m[i]>1:
                for k in range(I[i],I[i+1]):
                    j=J[k]
                    if m[j]>1:
                        if m[i]>m[j]: cf[j]=0
                        elif m[j]>m[i]: cf[i]=0
        cin=cf.copy(); cout=cf.copy()
        for i in range(n):
            c=cin[i]
            if ing[i]:
                if m[i]<1: c=-1
                if c>0: c=1
                else:
                    for k in range(I[i],I[i+1]):
                        if cin[J[k]]>0: c=-1
            cout[i]=c
        for i in range(n):
            if ing[i] and cout[i]!=0: m[i]=0
        cf=cout; it+=1
    return cf,demoted
