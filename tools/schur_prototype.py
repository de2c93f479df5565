"""Prototype + accuracy check (vs LAPACK) of the Toeplitz solve implemented by wh_solve_kernel in
blah2_b200/csrc/wh.cu: fraction-free / square-root-free generalized Schur recursion (reflection
coefficients + forward-substitution innovations, no inner products) driving a Levinson-type
accumulation of the solution, all elementwise per step.     python tools/schur_prototype.py"""
import math

import numpy as np
import scipy.linalg as sla


def solve(a, b):
    """A(i,j) = a[j-i] (j>=i), conj(a[i-j]) (i>j);  returns w with A w = b, or None if not PD."""
    n = len(a)
    t = np.conj(a).astype(complex)
    t0 = a[0].real
    if not t0 > 0:
        return None
    al = t / t0                      # Schur generators (scaled): a_i, b_i ; pivot p = a_k
    be = al.copy(); be[0] = 0
    r = b.astype(complex).copy()     # forward-substitution residuals (innovations)
    phi = np.zeros(n, complex); phi[0] = 1.0   # monic predictor, f = phi * sigma solves T f = e_1
    x = np.zeros(n, complex)
    p, inv_p, sc, sigma = 1.0, 1.0, 1.0, 1.0 / t0
    for k in range(n - 1):
        bk, rk = be[k + 1], r[k]
        rho = bk * inv_p
        # solution: x_i += (r_k sigma) conj(phi[k-i]),  i <= k
        x[:k + 1] += (rk * sigma) * np.conj(phi[:k + 1][::-1])
        # predictor: phi'[i] = phi[i] (i<=k) - rho conj(phi[k+1-i]) (i>=1),  i <= k+1
        pn_ = np.zeros(k + 2, complex)
        pn_[:k + 1] = phi[:k + 1]
        pn_[1:] -= rho * np.conj(phi[:k + 1][::-1])
        phi[:k + 2] = pn_
        # Schur: forward substitution + generator update, i > k
        q = rk * inv_p
        r[k + 1:] -= al[k + 1:] * q
        ps, bs = p * sc, bk * sc
        pn = ps * p - (bs * np.conj(bk)).real
        if not pn > 0:
            return None
        alt = np.zeros(n, complex); alt[k + 1:] = al[k:n - 1]
        al, be = ps * alt - np.conj(bs) * be, ps * be - bs * alt
        inv_pn = 1.0 / pn
        sigma = sigma * sc * p * p * inv_pn      # sigma / (1 - |rho|^2)
        p, inv_p = pn, inv_pn
        sc = 2.0 ** (1 - math.frexp(p * p)[1])
    x += (r[n - 1] * sigma) * np.conj(phi[::-1])
    return x


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 64, 410, 2048):
        N = 40000
        xs = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        xs = np.convolve(xs, [1, 0.9, 0.5j, 0.2])[:N] * 1500
        ys = 0.5 * xs + 0.2j * np.roll(xs, 3) + 75 * (rng.standard_normal(N) + 1j * rng.standard_normal(N))
        X, Y = np.fft.fft(xs), np.fft.fft(ys)
        a = np.conj(np.fft.ifft(X * np.conj(X)) * N)[:n]
        b = (np.fft.ifft(Y * np.conj(X)) * N)[:n]
        ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        A = a[np.abs(ii - jj)]
        A = np.where(ii > jj, np.conj(A), A)
        w_ref = sla.solve(A, b, assume_a="pos")
        w = solve(a, b)
        print(n, "cond %.1f" % np.linalg.cond(A), "rel err %.2e" % (np.abs(w - w_ref).max() / np.abs(w_ref).max()))
    assert solve(np.array([1.0, 2.0 + 0j]), np.array([1.0, 1.0 + 0j])) is None     # not PD
