"""NumPy restatement of the DEVICE draw generator (vbmc_b200/csrc/philox.cu) — test infrastructure only.

The reference draws its entropy samples from MATLAB's global ``randn`` stream (ent/entmc_vbmc.m:53; mt19937ar +
ziggurat), which cannot be reproduced outside MATLAB.  The product therefore carries its own counter-based generator
(generator mode, ``rng=(seed, stream)``): Philox4x32-10 (Salmon et al., SC'11; Random123 known answers in
tests/test_gpu_parity.py) feeding a 1024-strip ziggurat for the standard normal (Marsaglia & Tsang 2000 in Doornik's
ZIGNOR formulation, the same family of algorithm MATLAB's randn uses).  This file restates that generator element by
element so that tests can (i) check its distribution on the CPU and (ii) pin the device output against it.

Element e of the flat draw array eps[K][Ns/2][D] belongs to Philox counter c = e >> 1 (words 0-1 for even e, 2-3 for
odd e); counter = (c_lo, c_hi, stream_lo, stream_hi), key = (seed_lo, seed_hi).  A 64-bit word w gives the strip index
i = w & 1023 and the symmetric uniform u = ((w >> 12) + 0.5) * 2^-51 - 1.  Rejected attempts (0.3 % of the elements; 1024 strips
keep a 32-lane warp on the one-multiply fast path 90 % of the time) take
their extra uniforms from further Philox calls with the same counter and the key perturbed by (round, element parity).
"""
import math

import numpy as np

ZIG_C = 1024
ZIG_R = 4.038849846109504522714     # right edge of the base strip (1024 strips); bisection of the closure condition, 40 digits
ZIG_V = 0.001226324646353088072885  # area of every strip

M32 = np.uint64(0xFFFFFFFF)


def zig_tables():
    """x[0..C] (x[0] virtual base-strip width, x[1] = R, decreasing to x[C] = 0), r[i] = x[i+1]/x[i], f[i] = exp(-x[i]^2/2)."""
    x = np.zeros(ZIG_C + 1)
    f = math.exp(-0.5 * ZIG_R * ZIG_R)
    x[0] = ZIG_V / f
    x[1] = ZIG_R
    for i in range(2, ZIG_C):
        x[i] = math.sqrt(-2.0 * math.log(ZIG_V / x[i - 1] + f))
        f = math.exp(-0.5 * x[i] * x[i])
    x[ZIG_C] = 0.0
    r = x[1:] / x[:-1]
    fz = np.array([math.exp(-0.5 * v * v) for v in x])
    return x, r, fz


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all arguments uint64 arrays/scalars holding 32-bit values."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & M32 for v in (c0, c1, c2, c3))
    k0 = np.asarray(k0, dtype=np.uint64) & M32
    k1 = np.asarray(k1, dtype=np.uint64) & M32
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & M32, p1 >> np.uint64(32), p1 & M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & M32, lo1, (hi0 ^ c3 ^ k1) & M32, lo0
        k0, k1 = (k0 + W0) & M32, (k1 + W1) & M32
    return c0, c1, c2, c3


def _retry_key(seed, rnd, which):
    """Key of the extra Philox calls of a rejected attempt (csrc/philox.cu retry_key)."""
    k0 = (np.uint64(seed & 0xFFFFFFFF) ^ ((np.uint64(0xA5A5A5A5) + np.uint64(0x9E3779B9) * np.asarray(rnd, dtype=np.uint64)) & M32)) & M32
    k1 = np.uint64((seed >> 32) & 0xFFFFFFFF) ^ np.where(np.asarray(which) != 0, np.uint64(0xC2B2AE35), np.uint64(0x27D4EB2F))
    return k0, k1 & M32


def _u_sym(w):
    return ((w >> np.uint64(12)).astype(np.float64) + 0.5) * 2.0 ** -51 - 1.0


def _u01(w):
    return ((w >> np.uint64(12)).astype(np.float64) + 0.5) * 2.0 ** -52


def normals(seed, stream, e_begin, e_end):
    """Elements [e_begin, e_end) of the flat draw array of (seed, stream)."""
    x, r, fz = zig_tables()
    e = np.arange(e_begin, e_end, dtype=np.uint64)
    c = e >> np.uint64(1)
    which = (e & np.uint64(1)).astype(np.int64)
    c_lo, c_hi = c & M32, c >> np.uint64(32)
    s_lo, s_hi = np.uint64(stream & 0xFFFFFFFF), np.uint64((stream >> 32) & 0xFFFFFFFF)
    q = philox4x32_10(c_lo, c_hi, s_lo, s_hi, np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF))
    w = np.where(which == 0, q[0] | (q[1] << np.uint64(32)), q[2] | (q[3] << np.uint64(32)))
    out = np.full(e.size, np.nan)
    todo = np.arange(e.size)
    rnd = np.zeros(e.size, dtype=np.uint64)
    while todo.size:
        wi = w[todo]
        i = (wi & np.uint64(ZIG_C - 1)).astype(np.int64)
        u = _u_sym(wi)
        fast = np.abs(u) < r[i]
        out[todo[fast]] = u[fast] * x[i[fast]]
        todo, i, u = todo[~fast], i[~fast], u[~fast]
        if not todo.size:
            break
        rnd[todo] += np.uint64(1)
        k0, k1 = _retry_key(seed, rnd[todo], which[todo])
        q = philox4x32_10(c_lo[todo], c_hi[todo], s_lo, s_hi, k0, k1)
        wa, wb = q[0] | (q[1] << np.uint64(32)), q[2] | (q[3] << np.uint64(32))
        # ---- base strip: sample the tail beyond R (Marsaglia 1964) ----
        tail = i == 0
        tidx, ta, tb, tu = todo[tail], wa[tail], wb[tail], u[tail]
        while tidx.size:
            xx = -np.log(_u01(ta)) / ZIG_R
            yy = -np.log(_u01(tb))
            ok = yy + yy >= xx * xx
            out[tidx[ok]] = np.where(tu[ok] < 0, -(ZIG_R + xx[ok]), ZIG_R + xx[ok])
            tidx, tu = tidx[~ok], tu[~ok]
            if not tidx.size:
                break
            rnd[tidx] += np.uint64(1)
            k0, k1 = _retry_key(seed, rnd[tidx], which[tidx])
            q = philox4x32_10(c_lo[tidx], c_hi[tidx], s_lo, s_hi, k0, k1)
            ta, tb = q[0] | (q[1] << np.uint64(32)), q[2] | (q[3] << np.uint64(32))
        # ---- wedge of strip i ----
        todo, i, u, wa, wb = todo[~tail], i[~tail], u[~tail], wa[~tail], wb[~tail]
        xv = u * x[i]
        ok = fz[i] + _u01(wb) * (fz[i + 1] - fz[i]) < np.exp(-0.5 * xv * xv)   # uniform height inside the strip vs the density
        out[todo[ok]] = xv[ok]
        todo = todo[~ok]
        w[todo] = wa[~ok]   # next attempt: a fresh (strip, u) word
    return out
