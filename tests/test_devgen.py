"""The device draw generator (vbmc_b200/csrc/philox.cu: Philox4x32-10 + 1024-strip ziggurat) and its NumPy restatement
oracle/devgen.py.  The reference draws from MATLAB's randn stream (ent/entmc_vbmc.m:53), which cannot be reproduced outside
MATLAB; what can be pinned is (i) the counter-based generator against the Random123 known answers, (ii) the normal
transform against the N(0,1) distribution, (iii) the device output against the restatement, element by element."""
import numpy as np
import pytest
from scipy import stats

from oracle import devgen


def test_philox_restatement_known_answers():
    """Random123 kat_vectors for philox4x32-10 (the same vectors the device kernel is checked against)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        q = devgen.philox4x32_10(*ctr, *key)
        assert tuple(int(v) for v in q) == out


def test_ziggurat_tables_close():
    """Equal-area strips: x[i-1] (f(x[i]) - f(x[i-1])) = V for every strip, and the recursion ends at f(0) = 1."""
    x, r, fz = devgen.zig_tables()
    f = np.exp(-0.5 * x * x)
    assert np.allclose(fz, f, rtol=1e-15)
    assert x[1] == devgen.ZIG_R and x[-1] == 0.0 and np.all(np.diff(x) < 0)
    assert np.allclose(x[1:-1] * (f[2:] - f[1:-1]), devgen.ZIG_V, rtol=1e-11)
    assert abs(x[0] * f[1] - devgen.ZIG_V) < 1e-15                       # base strip: rectangle R f(R) + tail = V
    tail = np.sqrt(np.pi / 2) * 2 * stats.norm.sf(devgen.ZIG_R)
    assert abs(devgen.ZIG_R * f[1] + tail - devgen.ZIG_V) < 1e-14
    assert np.all((r >= 0) & (r < 1))


def test_ziggurat_restatement_is_standard_normal():
    z = devgen.normals(12345, 7, 0, 2_000_000)
    assert not np.isnan(z).any()
    assert abs(z.mean()) < 3e-3 and abs(z.std() - 1) < 3e-3
    assert abs(stats.skew(z)) < 0.01 and abs(stats.kurtosis(z)) < 0.02
    assert stats.kstest(z, "norm").pvalue > 1e-3
    edges = stats.norm.ppf(np.linspace(0, 1, 201))
    cnt, _ = np.histogram(z, bins=edges)
    assert stats.chisquare(cnt).pvalue > 1e-3
    for a in (2.0, 3.0, devgen.ZIG_R, 4.0):                               # wedges and the tail beyond R
        p, n = 2 * stats.norm.sf(a), z.size
        assert abs((np.abs(z) > a).mean() - p) < 5 * np.sqrt(p / n)
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 5e-3 and abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 5e-3
    # a slice is a function of (seed, stream, element index) only; other streams / seeds give other draws
    assert np.array_equal(devgen.normals(12345, 7, 1000, 1100), z[1000:1100])
    assert not np.array_equal(devgen.normals(12345, 8, 0, 100), z[:100])
    assert not np.array_equal(devgen.normals(12346, 7, 0, 100), z[:100])


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(5, 8, 2048), (3, 3, 74), (10, 50, 4096)])
def test_device_generator_equals_restatement(gpu_ctx, shape):
    """Every element of the device draws equals the restatement (tails go through log(): 1-ulp libm differences allowed)."""
    D, K, Ns = shape
    seed, stream = 0x1234567890ABCDEF, 0xFEDCBA9876543210 - 5
    eps = gpu_ctx.eps_philox(D, K, Ns, seed, stream, readback=True)        # (K, Ns/2, D): flat element e = (k*Ns/2 + s)*D + d
    flat = eps.ravel()
    ref = devgen.normals(seed, stream, 0, flat.size)
    assert np.max(np.abs(flat - ref)) < 1e-14
    assert (flat == ref).mean() > 0.999
