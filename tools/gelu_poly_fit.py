"""Minimax fit and EXHAUSTIVE f16 / bf16 evaluation (all finite 16-bit inputs) of the polynomial GELU used by the tensor-core
decode kernels (gelu_poly_pair in csrc/nic_tc_common.cuh): gelu(x) = x * clamp(1/2 + x Q(x^2), 0, 1).  Variants B (f16) and E
(bf16) are the ones compiled in.  CPU only: python tools/gelu_poly_fit.py"""
import numpy as np, math
from scipy.special import erf
from scipy.optimize import minimize
bits = np.arange(65536, dtype=np.uint16)
x16 = bits.view(np.float16); x16 = x16[np.isfinite(x16)]
x = x16.astype(np.float64)
gelu = 0.5 * x * (1 + erf(x / math.sqrt(2)))
def rnd(v, fmt):
    v = np.asarray(v, dtype=np.float64)
    if fmt == 'f16':
        with np.errstate(over='ignore'): return v.astype(np.float16).astype(np.float64)
    # bf16 round-to-nearest-even via float32 bits
    f = v.astype(np.float32); u = f.view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    u2 = ((u + r) & 0xFFFF0000).astype(np.uint32)
    return u2.view(np.float32).astype(np.float64)
def sat(v): return np.clip(np.nan_to_num(v, nan=0.0), 0.0, 1.0)
def evalp(c, fmt, xin, L2=None):
    xx = rnd(xin, fmt); s = rnd(xx * xx, fmt)
    if L2 is not None: s = np.minimum(s, rnd(L2, fmt))
    with np.errstate(over='ignore', invalid='ignore'):
        q = rnd(c[-1], fmt) * np.ones_like(s)
        for k in range(len(c) - 2, -1, -1): q = rnd(q * s + rnd(c[k], fmt), fmt)
        phi = sat(rnd(xx * q + 0.5, fmt))
        return rnd(xx * phi, fmt)
def objective(c, L2, xs):
    # exact-arithmetic max error in gelu units over xs
    s = xs * xs
    if L2 is not None: s = np.minimum(s, L2)
    q = np.polyval(c[::-1], s)
    phi = np.clip(xs * q + 0.5, 0, 1)
    return np.max(np.abs(xs * phi - 0.5 * xs * (1 + erf(xs / math.sqrt(2)))))
xs = np.concatenate([np.linspace(-8, 8, 6401)])
res = {}
for name, d, clamp in (("A d3 clamp", 3, True), ("B d4 free", 4, False), ("C d2 free", 2, False), ("D d2 clamp", 2, True), ("E d4 clamp", 4, True)):
    best = None
    for L in (2.6, 2.8, 3.0, 3.2, 3.4, 3.6):
        # init by lstsq
        xf = np.linspace(1e-3, L, 2000); A = np.stack([xf ** (2 * k) for k in range(d + 1)], 1)
        c0, *_ = np.linalg.lstsq(A * (xf ** 2)[:, None], (0.5 * erf(xf / math.sqrt(2)) / xf) * xf ** 2, rcond=None)
        L2 = L * L if clamp else None
        scale = np.array([10.0 ** (-k) for k in range(d + 1)])
        f = lambda z: objective(z * scale, L2, xs)
        r = minimize(f, c0 / scale, method='Nelder-Mead', options={'xatol': 1e-9, 'fatol': 1e-9, 'maxiter': 20000, 'maxfev': 20000})
        c = r.x * scale
        if best is None or r.fun < best[0]: best = (r.fun, L, c)
    fun, L, c = best
    L2 = L * L if clamp else None
    m = np.abs(x) < 60000
    out = [name, "L=%.1f exact-max %.2e" % (L, fun)]
    for fmt in ('f16', 'bf16'):
        g = evalp(c, fmt, x, L2)
        # compare against gelu of the rounded input
        xi = rnd(x, fmt); ge = 0.5 * xi * (1 + erf(xi / math.sqrt(2)))
        err = np.abs(g - ge)
        sel = np.abs(x) < 4
        out.append("%s max %.2e (x=%.2f) rms|x|<4 %.2e, max|x|>4: %.2e" % (fmt, err[m].max(), x[m][err[m].argmax()], np.sqrt((err[sel] ** 2).mean()), (err[m & ~sel] / np.maximum(1, np.abs(x[m & ~sel]))).max()))
    print(" | ".join(out)); print("   c =", list(c), "L2 =", L2)
for fmt in ('f16', 'bf16'):
    xx = rnd(x, fmt); x2 = rnd(xx*xx, fmt); p = rnd(x2*0.0356774081+0.7978845608, fmt); u = rnd(p*xx, fmt); t = rnd(np.tanh(u), fmt); h = rnd(xx*t+xx, fmt)
    ge = 0.5 * xx * (1 + erf(xx / math.sqrt(2)))
    err = np.abs(0.5*h - ge); m = np.abs(x) < 60000; sel = np.abs(x) < 4
    print("tanh form %s: max %.2e at %.3f rms %.2e" % (fmt, err[m].max(), x[m][err[m].argmax()], np.sqrt((err[sel]**2).mean())))
