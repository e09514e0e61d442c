"""CPU emulation of the index algebra in csrc/fft256w.cuh (256-point complex FFT on a whole warp,
256 = 8 x 4 x 8, natural order in and out: lane q holds element q + 32 j in register j) and of the
split / merge of two real frames riding in one complex transform.  Run before touching the kernel."""
import numpy as np

W = lambda n, e, inv=False: np.exp((2j if inv else -2j) * np.pi * e / n)


def wfft256(V, inv=False):
    """V[lane, reg]: element lane + 32 reg.  Returns the transform in the same layout."""
    V = V.copy()
    # stage 1: 8-point DFT over the register index (m -> k0), twiddle W256^(q k0)
    for q in range(32):
        V[q] = np.array([sum(V[q, m] * W(8, m * k0, inv) for m in range(8)) for k0 in range(8)])
        V[q] *= np.array([W(256, q * k0, inv) for k0 in range(8)])
    # exchange 1 (its own region): A[k0][q], pitch 34; lane (k0, h) = k0 + 8 h reads pairs at A[k0][2h + 8 q1]
    A = np.zeros(8 * 34, complex)
    for q in range(32):
        for k0 in range(8):
            A[k0 * 34 + q] = V[q, k0]
    U = np.zeros((32, 8), complex)
    for lam in range(32):
        k0, h = lam & 7, lam >> 3
        for q1 in range(4):
            U[lam, 2 * q1] = A[k0 * 34 + 2 * h + 8 * q1]
            U[lam, 2 * q1 + 1] = A[k0 * 34 + 2 * h + 8 * q1 + 1]
    # stage 2: 4-point DFT over q1 for e = 0, 1; twiddle W32^(q0 k1), q0 = 2h + e
    for lam in range(32):
        h = lam >> 3
        out = np.zeros(8, complex)
        for e in range(2):
            for k1 in range(4):
                out[2 * k1 + e] = sum(U[lam, 2 * q1 + e] * W(4, q1 * k1, inv) for q1 in range(4)) * W(32, (2 * h + e) * k1, inv)
        U[lam] = out
    # exchange 2: row k0 + 8 k1 (pitch 10), column h + 4 e holds q0 = 2 h + e (64-bit stores without bank conflicts)
    B = np.zeros(32 * 10, complex)
    for lam in range(32):
        k0, h = lam & 7, lam >> 3
        for k1 in range(4):
            B[(k0 + 8 * k1) * 10 + h] = U[lam, 2 * k1]
            B[(k0 + 8 * k1) * 10 + h + 4] = U[lam, 2 * k1 + 1]
    out = np.zeros((32, 8), complex)
    for l in range(32):
        row = B[l * 10:l * 10 + 8]
        v = np.zeros(8, complex)
        # the four 128-bit reads: columns (0,1) (2,3) (4,5) (6,7) = q0 (0,2) (4,6) (1,3) (5,7)
        v[0], v[2], v[4], v[6], v[1], v[3], v[5], v[7] = row
        out[l] = np.array([sum(v[q0] * W(8, q0 * k2, inv) for q0 in range(8)) for k2 in range(8)])
    return out


def to_lanes(x):
    return x.reshape(8, 32).T.copy()            # [lane, reg] = x[lane + 32 reg]


def from_lanes(V):
    return V.T.reshape(256).copy()


def split_planar(V):
    """Half spectra of the two real frames (a in re, b in im), 0.5 not applied: slots r = 0..4 per lane."""
    XA = np.zeros((32, 5), complex)
    XB = np.zeros((32, 5), complex)
    for q in range(32):
        src = (32 - q) & 31
        for r in range(5):
            got = V[src, 7 - r if r < 4 else 3]
            own = V[q, (8 - r) & 7]
            zp = own if q == 0 else got
            z = V[q, r]
            XA[q, r] = z + np.conj(zp)
            XB[q, r] = -1j * (z - np.conj(zp))
    return XA, XB


def merge_pair(L, Mi):
    V = np.zeros((32, 8), complex)
    for q in range(32):
        src = (32 - q) & 31
        for j in range(4):
            V[q, j] = L[q, j]
        for j in range(4, 8):
            got = Mi[src, 7 - j]
            own = L[q, 4] if j == 4 else Mi[q, 8 - j]
            V[q, j] = own if q == 0 else got
    return V


def main():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(256) + 1j * rng.standard_normal(256)
    X = from_lanes(wfft256(to_lanes(x)))
    print("forward err", np.abs(X - np.fft.fft(x)).max())
    xb = from_lanes(wfft256(to_lanes(X), inv=True)) / 256
    print("round trip err", np.abs(xb - x).max())
    a, b = rng.standard_normal(256), rng.standard_normal(256)
    V = wfft256(to_lanes(0.5 * (a + 1j * b)))
    XA, XB = split_planar(V)
    FA, FB = np.fft.rfft(a), np.fft.rfft(b)
    err = 0.0
    for q in range(32):
        for r in range(5):
            k = q + 32 * r
            if k <= 128 and (r < 4 or q == 0):
                err = max(err, abs(XA[q, r] - FA[k]), abs(XB[q, r] - FB[k]))
    print("split err", err)
    # inverse of (P, Q) = masked spectra: L = P + iQ, Mi = conj P + i conj Q
    L = XA + 1j * XB
    Mi = np.conj(XA) + 1j * np.conj(XB)
    y = from_lanes(wfft256(merge_pair(L, Mi), inv=True)) / 256
    print("merge err", np.abs(y.real - a).max(), np.abs(y.imag - b).max())


if __name__ == "__main__":
    main()
