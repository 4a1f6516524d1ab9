"""Numpy prototype of the in-kernel eigen-solver (tridiagonalise -> multisection -> inverse iteration -> back-transform).
Kept as documentation of the algorithm the CUDA kernel in csrc/spectral.cu transcribes."""
import numpy as np

def tridiag(M):
    M = M.copy(); G = M.shape[0]
    d = np.zeros(G); e = np.zeros(G-1); tau = np.zeros(G-1)
    for k in range(G-2):
        x = M[k+1:, k].copy()
        alpha = x[0]
        xnorm2 = np.dot(x[1:], x[1:])
        if xnorm2 == 0.0:
            tau[k] = 0.0; e[k] = alpha; d[k] = M[k,k]
            M[k+1:, k] = 0; M[k+1,k] = 1.0
            continue
        beta = -np.copysign(np.sqrt(alpha*alpha + xnorm2), alpha)
        tau[k] = (beta - alpha)/beta
        scale = 1.0/(alpha - beta)
        v = x*scale; v[0] = 1.0
        # p = tau * A v
        A = M[k+1:, k+1:]
        p = tau[k] * (A @ v)
        w = p - (0.5*tau[k]*np.dot(p, v))*v
        A -= np.outer(v, w) + np.outer(w, v)
        d[k] = M[k,k]; e[k] = beta
        M[k+1:, k] = v
    d[G-2] = M[G-2,G-2]; d[G-1] = M[G-1,G-1]; e[G-2] = M[G-1,G-2]
    return d, e, tau, M

def sturm_count(d, e2, x):
    """# eigenvalues < x via the scaled multiplicative recurrence."""
    G = len(d)
    pm1 = 1.0; p = d[0]-x; cnt = 1 if p < 0 else 0
    sgn_prev = -1 if p < 0 else 1
    for i in range(1, G):
        pn = (d[i]-x)*p - e2[i-1]*pm1
        pm1, p = p, pn
        # rescale
        a = abs(p)
        if a > 1e100: p *= 1e-100; pm1 *= 1e-100
        elif a < 1e-100 and a > 0: p *= 1e100; pm1 *= 1e100
        if p == 0.0:
            s = -sgn_prev
        else:
            s = -1 if p < 0 else 1
        # eigenvalue count = number of sign agreements?  use: count increments when sign(p_i) != sign(p_{i-1}) ... see below
        if s != sgn_prev: cnt_change = 1
        else: cnt_change = 0
        # q_i = p_i/p_{i-1} negative <=> sign change
        cnt += 0
        sgn_prev_old = sgn_prev
        sgn_prev = s
        if s != sgn_prev_old: cnt += 1
    return cnt

def sturm_count2(d, e2, x):
    # reference: LDL pivots
    q = d[0]-x; cnt = int(q<0)
    for i in range(1,len(d)):
        if q == 0: q = 1e-300
        q = d[i]-x - e2[i-1]/q
        cnt += int(q<0)
    return cnt

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    G=64
    A = rng.standard_normal((G,G)); A=(A+A.T)/2
    d,e,tau,M = tridiag(A)
    T = np.diag(d)+np.diag(e,1)+np.diag(e,-1)
    w0 = np.linalg.eigvalsh(A); w1=np.linalg.eigvalsh(T)
    print("tridiag eig err", np.abs(w0-w1).max())
    e2=e*e
    for x in [-3,-1,0,0.5,2]:
        print(x, sturm_count(d,e2,x), sturm_count2(d,e2,x), (w0<x).sum())


def multisection(d, e, want, npts=64, rounds=12):
    """eigenvalue with 0-based index `want` (ascending) by repeated npts-way section on Sturm counts."""
    e2 = e*e
    r = np.abs(np.concatenate(([0], e))) + np.abs(np.concatenate((e, [0])))
    lo = (d - r).min(); hi = (d + r).max()
    span = hi - lo; lo -= 1e-3*span + 1e-300; hi += 1e-3*span + 1e-300
    for _ in range(rounds):
        xs = lo + (hi-lo)*(np.arange(1, npts+1)/(npts+1))
        cnts = np.array([sturm_count(d, e2, x) for x in xs])
        # largest x_j with count <= want  -> new lo ; smallest x_j with count > want -> new hi
        below = cnts <= want
        nb = below.sum()   # counts are monotone
        new_lo = xs[nb-1] if nb > 0 else lo
        new_hi = xs[nb] if nb < npts else hi
        lo, hi = new_lo, new_hi
        if hi - lo <= 4e-16*max(abs(lo), abs(hi)): break
    return 0.5*(lo+hi)

def tri_inverse_iteration(d, e, lam, prev=(), iters=3):
    G = len(d)
    # LU with partial pivoting of T - lam I (rows: a=sub, b=diag, c=super)
    norm = np.abs(d).max() + 2*np.abs(e).max()
    eps = 2.2e-16*norm
    a = np.concatenate(([0.0], e)); b = d - lam; c = np.concatenate((e, [0.0]))
    u0 = np.zeros(G); u1 = np.zeros(G); u2 = np.zeros(G); l = np.zeros(G); piv = np.zeros(G, dtype=bool)
    # current row i (to be eliminated against row i+1)
    r0, r1, r2 = b[0], c[0], 0.0
    for i in range(G-1):
        s0, s1, s2 = a[i+1], b[i+1], c[i+1]   # next row: entries at cols i, i+1, i+2
        if abs(s0) > abs(r0):
            piv[i] = True
            r0, r1, r2, s0, s1, s2 = s0, s1, s2, r0, r1, r2
        if r0 == 0.0: r0 = eps
        m = s0/r0
        u0[i], u1[i], u2[i], l[i] = r0, r1, r2, m
        r0, r1, r2 = s1 - m*r1, s2 - m*r2, 0.0
    if abs(r0) < eps: r0 = eps if r0 >= 0 else -eps
    u0[G-1] = r0
    z = np.array([((i*7919) % 13 - 6)/6.0 + 0.37 for i in range(G)])
    for it in range(iters):
        # forward: apply row ops
        y = z.copy()
        for i in range(G-1):
            if piv[i]: y[i], y[i+1] = y[i+1], y[i]
            y[i+1] -= l[i]*y[i]
        # back substitution
        x = np.zeros(G)
        for i in range(G-1, -1, -1):
            t = y[i]
            if i+1 < G: t -= u1[i]*x[i+1]
            if i+2 < G: t -= u2[i]*x[i+2]
            x[i] = t/u0[i]
        for p in prev:
            x -= np.dot(p, x)*p
        z = x/np.linalg.norm(x)
    return z

def back_transform(M, tau, z):
    G = len(z); x = z.copy()
    for k in range(G-3, -1, -1):
        v = M[k+1:, k]
        s = tau[k]*np.dot(v, x[k+1:])
        x[k+1:] -= s*v
    return x

if __name__ == "__main__":
    for G in (64, 128):
        A = rng.standard_normal((G,G)); A=(A+A.T)/2
        d,e,tau,M = tridiag(A)
        w0, V0 = np.linalg.eigh(A)
        prev=[]
        for s in range(4):
            lam = multisection(d, e, s)
            z = tri_inverse_iteration(d, e, lam, prev)
            prev.append(z)
            x = back_transform(M, tau, z)
            x *= np.sign(x[0]); v = V0[:,s]*np.sign(V0[0,s])
            print(G, s, "lam err", abs(lam-w0[s]), "vec err", np.abs(x-v).max(), "resid", np.linalg.norm(A@x-lam*x))
