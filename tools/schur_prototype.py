"""Prototype + accuracy check (vs LAPACK) of the square-root-free Schur Toeplitz solve implemented by
wh_solve_kernel in blah2_b200/csrc/wh.cu.  python tools/schur_prototype.py"""
import numpy as np, scipy.linalg as sla, math
def schur_ff(a,b):
    n=len(a); t=np.conj(a).astype(complex); t0=a[0].real
    al=t/t0; be=al.copy(); be[0]=0      # scaled generators a_i^0 = t_i/t_0, p_0 = 1, G_0 = 1/t_0
    p=1.0; G=1.0/t0
    r=b.astype(complex).copy()
    Lff=np.zeros((n,n),complex); pv=np.zeros(n); u=np.zeros(n,complex)
    for k in range(n):
        Lff[k:,k]=al[k:]; pv[k]=p
        u[k]=r[k]*G/p                 # u_k = z_k * gamma_k = r_k G_k / p_k
        q=r[k]/p
        r[k+1:]-=al[k+1:]*q
        if k==n-1: break
        bk=be[k+1]
        # power-of-two scale from exponent of p^2
        e=math.frexp(p*p)[1]; s=2.0**(-e+1)
        pn=s*(p*p-abs(bk)**2)
        if not pn>0: return None
        alt=np.zeros(n,complex); alt[k+1:]=al[k:n-1]
        aln=s*(p*alt-np.conj(bk)*be); ben=s*(p*be-bk*alt)
        al=aln; be=ben
        G=G*s*pn; p=pn
    w=np.zeros(n,complex)
    for k in range(n-1,-1,-1):
        w[k]=u[k]/pv[k]
        u[:k]-=np.conj(Lff[k,:k])*w[k]
    return w, (p,G)
rng=np.random.default_rng(0)
for n in (5,64,410,2048):
    N=40000
    xs=rng.standard_normal(N)+1j*rng.standard_normal(N)
    xs=np.convolve(xs,[1,0.9,0.5,0.2])[:N]*1500
    ys=0.5*xs+0.2*np.roll(xs,3)+0.05*1500*(rng.standard_normal(N)+1j*rng.standard_normal(N))
    X=np.fft.fft(xs);Y=np.fft.fft(ys)
    a=np.conj(np.fft.ifft(X*np.conj(X))*N)[:n]/N*N; b=(np.fft.ifft(Y*np.conj(X))*N)[:n]/N*N
    ii,jj=np.meshgrid(np.arange(n),np.arange(n),indexing='ij')
    A=a[np.abs(ii-jj)]; A=np.where(ii>jj,np.conj(A),A)
    w_ref=sla.solve(A,b,assume_a='pos')
    w,(p,G)=schur_ff(a,b)
    print(n,np.linalg.cond(A), np.abs(w-w_ref).max()/np.abs(w_ref).max(), 'final p,G',p,G)
