"""Index arithmetic of the one-launch single-column triangular sweep (vbmc_b200/csrc/trsm.cu, trsv1_kernel), restated in Python:
P CTAs per sample, block J owned by CTA J mod P, per block step the owner solves its 64x64 diagonal tile and every CTA updates
the blocks it owns that are not solved yet, four per pass.  The model executes the CTAs' passes exactly as the kernel enumerates
them (Jfirst, block_of, the four-block passes) and checks (i) every unsolved block receives the step's update exactly once, from
its owner only, (ii) the result equals a dense triangular solve -- for sizes and CTA counts beyond those the GPU tests run."""
import numpy as np
import pytest

TT = 64


def _passes(b, me, P, nbN, backward):
    """blocks CTA `me` updates in step b, in the kernel's order (list of passes of up to 4 blocks)"""
    if backward:
        Jfirst = -1 if b == 0 else b - 1 - ((b - 1 - me) % P + P) % P
    else:
        Jfirst = b + 1 + ((me - (b + 1)) % P + P) % P
        if Jfirst >= nbN:
            Jfirst = -1

    def block_of(J0, q):
        if J0 < 0:
            return -1
        J = J0 - q * P if backward else J0 + q * P
        return -1 if (J < 0 or J >= nbN) else J

    out = []
    J0 = Jfirst
    while 0 <= J0 < nbN:
        out.append([block_of(J0, q) for q in range(4)])
        J0 += -4 * P if backward else 4 * P
    return out


def _sweep(R, z, P, backward):
    N = R.shape[0]
    nbN = (N + TT - 1) // TT
    z = z.copy()
    order = range(nbN - 1, -1, -1) if backward else range(nbN)
    for b in order:
        b0, b1 = b * TT, min(N, (b + 1) * TT)
        Rbb = R[b0:b1, b0:b1]
        x = np.linalg.solve(Rbb if backward else Rbb.T, z[b0:b1])       # the owner (b mod P) solves its diagonal tile
        z[b0:b1] = x
        touched = []
        for me in range(P):
            for blocks in _passes(b, me, P, nbN, backward):
                for J in blocks:
                    if J < 0:
                        continue
                    assert J % P == me, (b, me, J)
                    touched.append(J)
                    j0, j1 = J * TT, min(N, (J + 1) * TT)
                    if backward:
                        z[j0:j1] -= R[j0:j1, b0:b1] @ x                  # z_J -= R(J, b) x_b
                    else:
                        z[j0:j1] -= R[b0:b1, j0:j1].T @ x                # z_J -= R(b, J)' x_b
        expect = list(range(b)) if backward else list(range(b + 1, nbN))
        assert sorted(touched) == expect, (b, sorted(touched), expect)
    return z


@pytest.mark.parametrize("N", [1, 63, 64, 65, 200, 777, 2001])
@pytest.mark.parametrize("P", [1, 2, 3, 7, 8])
@pytest.mark.parametrize("backward", [False, True])
def test_every_block_is_updated_once_and_the_solution_is_right(N, P, backward):
    rs = np.random.default_rng(N * 31 + P)
    nbN = (N + TT - 1) // TT
    P = min(P, nbN)
    R = np.triu(0.1 * rs.standard_normal((N, N))) + np.diag(1.0 + rs.random(N))
    z = rs.standard_normal(N)
    got = _sweep(R, z, P, backward)
    ref = np.linalg.solve(R if backward else R.T, z)
    assert np.max(np.abs(got - ref)) <= 1e-9 * max(1.0, np.max(np.abs(ref)))


def test_owned_block_storage_index():
    """the kernel keeps block J of CTA J mod P at local index J // P: distinct and within ceil(nbN / P)"""
    for nbN in (1, 5, 32, 63):
        for P in (1, 2, 7, 8):
            P = min(P, nbN)
            cap = (nbN + P - 1) // P
            for me in range(P):
                idx = [J // P for J in range(me, nbN, P)]
                assert idx == list(range(len(idx))) and len(idx) <= cap
