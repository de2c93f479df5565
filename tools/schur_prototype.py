"""Prototype + accuracy check (vs LAPACK) of the Toeplitz solve implemented by wh_solve_kernel in
blah2_b200/csrc/wh.cu: generalized Schur recursion + Levinson accumulation with EVERY recurrence in
fraction-free form (no division / sqrt inside the n steps).     python tools/schur_prototype.py"""
import numpy as np, scipy.linalg as sla, math
def pow2_of(v):  # 2^(-e) with v = m 2^e, m in [0.5,1)
    return 2.0**(-math.frexp(v)[1])
def solve_ff(a,b):
    n=len(a); t=np.conj(a).astype(complex); t0=a[0].real
    al=t/t0; be=al.copy(); be[0]=0
    r=b.astype(complex).copy()
    phi=np.zeros(n,complex); phi[0]=1.0
    X=np.zeros(n,complex)
    p=1.0; sc=1.0; F=1.0; T=t0   # T = F^2/sigma
    mx=0
    for k in range(n-1):
        bk,rk=be[k+1],r[k]
        # scale u for r,phi: keep F in [0.5,1)
        u=pow2_of(F*p)
        ps=p*sc; bs=bk*sc; pu=p*u; bu=bk*u; rku=rk*u
        pnew=ps*p-(bs*np.conj(bk)).real
        if not pnew>0: return None
        c=pnew*u*u/sc
        # X update (lags nothing): X' = c (X + r_k conj(phi[k-i]))
        Xn=X.copy(); Xn[:k+1]+=rk*np.conj(phi[:k+1][::-1]); X=c*Xn
        # phi' = u (p phi - b conj(phi_m))
        ph=np.zeros(k+2,complex); ph[:k+1]=pu*phi[:k+1]; ph[1:]-=bu*np.conj(phi[:k+1][::-1]); phi[:k+2]=ph
        # r' = u (p r - a r_k), i>k
        r[k+1:]=pu*r[k+1:]-al[k+1:]*rku
        alt=np.zeros(n,complex); alt[k+1:]=al[k:n-1]
        al,be=ps*alt-np.conj(bs)*be, ps*be-bs*alt
        F=F*u*p; T=T*c
        p=pnew; sc=2.0**(1-math.frexp(p*p)[1])
        mx=max(mx,np.abs(r).max(),np.abs(X).max(),np.abs(phi).max())
    X[:]+=r[n-1]*np.conj(phi[::-1])
    return X/T, mx, (F,T,p)
rng=np.random.default_rng(0)
for n in (1,2,5,64,410,2048):
    N=40000
    xs=rng.standard_normal(N)+1j*rng.standard_normal(N); xs=np.convolve(xs,[1,0.9,0.5j,0.2])[:N]*1500
    ys=0.5*xs+0.2j*np.roll(xs,3)+75*(rng.standard_normal(N)+1j*rng.standard_normal(N))
    X_,Y_=np.fft.fft(xs),np.fft.fft(ys)
    a=np.conj(np.fft.ifft(X_*np.conj(X_))*N)[:n]; b=(np.fft.ifft(Y_*np.conj(X_))*N)[:n]
    ii,jj=np.meshgrid(np.arange(n),np.arange(n),indexing="ij")
    A=a[np.abs(ii-jj)]; A=np.where(ii>jj,np.conj(A),A)
    w_ref=sla.solve(A,b,assume_a="pos")
    w,mx,st=solve_ff(a,b)
    print(n,"rel err %.2e"%(np.abs(w-w_ref).max()/np.abs(w_ref).max()),"max mag %.2e"%mx, st)
