"""CPU emulation of the index math used by csrc/fft.cuh (developer tool, not product).

Checks, against numpy.fft, the formulas the CUDA kernels rely on:
  * warp Stockham autosort FFT (radix-4 stages + one radix-2 stage)
  * real FFT of n = 2M points through one M-point complex FFT (+ post twiddle)
  * its inverse (pre twiddle + M-point inverse FFT), DC/Nyquist imag ignored
  * 256-point complex FFT as 16 x 16 with one transpose, natural-order input,
    "lane k1 holds bins k1 + 16 k2" output, and the mirrored inverse
  * two real signals in one complex FFT (pair trick) and its inverse
"""
import numpy as np


def stockham(x, inverse=False):
    M = len(x)
    tw = np.exp(-2j * np.pi * np.arange(M) / M)
    if inverse:
        tw = tw.conj()
    a = np.array(x, dtype=np.complex128)
    b = np.empty_like(a)
    Ns = 1
    while Ns * 4 <= M:
        q = M // 4
        tstep = M // (Ns * 4)
        for j in range(q):
            k = j & (Ns - 1)
            v = [a[j + r * q] for r in range(4)]
            for r in range(1, 4):
                v[r] = v[r] * tw[r * k * tstep]
            a0, a1 = v[0] + v[2], v[0] - v[2]
            a2 = v[1] + v[3]
            a3 = (v[1] - v[3]) * (1j if inverse else -1j)
            d = ((j - k) << 2) + k
            b[d], b[d + Ns], b[d + 2 * Ns], b[d + 3 * Ns] = a0 + a2, a1 + a3, a0 - a2, a1 - a3
        a, b = b, a
        Ns *= 4
    if Ns < M:
        q = M // 2
        for j in range(q):
            k = j & (Ns - 1)
            w = tw[k * (M // (Ns * 2))]
            v0, v1 = a[j], a[j + q] * w
            d = ((j - k) << 1) + k
            b[d], b[d + Ns] = v0 + v1, v0 - v1
        a, b = b, a
    return a


def rfft_half(x, win):
    n = len(x); M = n // 2
    xw = x * win * 0.5
    Z = stockham(xw[0::2] + 1j * xw[1::2])
    tw = np.exp(-2j * np.pi * np.arange(M + 1) / n)
    X = np.empty(M + 1, complex)
    for k in range(M + 1):
        zk = Z[k % M]; zc = np.conj(Z[(M - k) % M])
        X[k] = (zk + zc) - 1j * tw[k] * (zk - zc)
    return X


def irfft_half(X):
    M = len(X) - 1; n = 2 * M
    X = X.copy(); X[0] = X[0].real; X[M] = X[M].real
    tw = np.exp(+2j * np.pi * np.arange(M) / n)
    Z = np.empty(M, complex)
    for k in range(M):
        a = X[k]; b = np.conj(X[M - k])
        Z[k] = (a + b) + 1j * tw[k] * (a - b)
    z = stockham(Z, inverse=True) / n
    out = np.empty(n)
    out[0::2] = z.real; out[1::2] = z.imag
    return out


def fft16(v, inverse=False):
    return np.fft.ifft(v) * 16 if inverse else np.fft.fft(v)


def fft256_16x16(x):
    """lane p holds x[p + 16 m]; result: lane k1 holds X[k1 + 16 k2]."""
    W = np.exp(-2j * np.pi / 256)
    A = np.empty((16, 16), complex)           # A[p][k1]
    for p in range(16):
        A[p] = fft16(x[p::16]) * W ** (p * np.arange(16))
    out = np.empty((16, 16), complex)         # out[k1][k2]
    for k1 in range(16):
        out[k1] = fft16(A[:, k1])
    return out


def ifft256_16x16(Y):
    """Y[k1][k2] = bin k1 + 16 k2; result lane p holds y[p + 16 m] (unnormalised)."""
    W = np.exp(+2j * np.pi / 256)
    Bm = np.empty((16, 16), complex)          # B[k1][p]
    for k1 in range(16):
        Bm[k1] = fft16(Y[k1], inverse=True) * W ** (k1 * np.arange(16))
    y = np.empty(256, complex)
    for p in range(16):
        y[p::16] = fft16(Bm[:, p], inverse=True)
    return y


def fft16_radix4(v, inverse=False):
    """The in-register 16-point FFT of fft256.cuh (n = 4a + b, k = c + 4d)."""
    v = list(v)
    sgn = 1 if inverse else -1

    def fft4(x0, x1, x2, x3):
        t0, t1, t2, t3 = x0 + x2, x0 - x2, x1 + x3, x1 - x3
        r = t3 * (1j if inverse else -1j)
        return t0 + t2, t1 + r, t0 - t2, t1 - r

    for b in range(4):
        v[b], v[4 + b], v[8 + b], v[12 + b] = fft4(v[b], v[4 + b], v[8 + b], v[12 + b])
    for c in range(1, 4):
        for b in range(1, 4):
            v[4 * c + b] *= np.exp(sgn * 2j * np.pi * b * c / 16)
    for c in range(4):
        v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3] = fft4(*v[4 * c:4 * c + 4])
    out = [0] * 16
    for c in range(4):
        for d in range(4):
            out[c + 4 * d] = v[4 * c + d]
    return np.array(out)


def lanes_split_pair(V):
    """V[lane][k2] = Z[lane + 16 k2] -> A[lane][r], B[lane][r] (r = 0..8)."""
    A = np.zeros((16, 9), complex); B = np.zeros((16, 9), complex)
    for l in range(16):
        src = (16 - l) & 15
        for r in range(9):
            got = V[src][15 - r if r < 8 else 7]
            own = V[l][(16 - r) & 15]
            zp = own if l == 0 else got
            z = V[l][r]
            A[l, r] = complex(z.real + zp.real, z.imag - zp.imag)
            B[l, r] = complex(z.imag + zp.imag, zp.real - z.real)
    return A, B


def lanes_merge_pair(L, Mi):
    V = np.zeros((16, 16), complex)
    for l in range(16):
        src = (16 - l) & 15
        for r in range(8):
            V[l, r] = L[l, r]
        for k2 in range(8, 16):
            got = Mi[src, 15 - k2]
            own = L[l, 8] if k2 == 8 else Mi[l, (16 - k2) & 7]
            V[l, k2] = own if l == 0 else got
    return V


def check_lane_algebra(rng):
    for inv in (False, True):
        x = rng.standard_normal(16) + 1j * rng.standard_normal(16)
        assert np.allclose(fft16_radix4(x, inv), np.fft.ifft(x) * 16 if inv else np.fft.fft(x))
    a = rng.standard_normal(256); b = rng.standard_normal(256)
    V = fft256_16x16(0.5 * (a + 1j * b))              # the 1/2 lives in the window
    A, B = lanes_split_pair(V)
    Fa, Fb = np.fft.rfft(a), np.fft.rfft(b)
    for l in range(16):
        for r in range(9 if l == 0 else 8):
            assert np.allclose(A[l, r], Fa[l + 16 * r]) and np.allclose(B[l, r], Fb[l + 16 * r])
    # inverse: P, Q half spectra (masked) -> frames p (re), q (im)
    m1 = rng.random(129); m2 = rng.random(129)
    P, Q = m1 * Fa, m2 * Fb
    L = np.zeros((16, 9), complex); Mi = np.zeros((16, 9), complex)
    for l in range(16):
        for r in range(9):
            k = l + 16 * r
            if k <= 128:
                L[l, r] = P[k] + 1j * Q[k]
                Mi[l, r] = np.conj(P[k]) + 1j * np.conj(Q[k])
    y = ifft256_16x16(lanes_merge_pair(L, Mi)) / 256
    assert np.allclose(y.real, np.fft.irfft(P)) and np.allclose(y.imag, np.fft.irfft(Q))


def main():
    rng = np.random.default_rng(0)
    for M in (16, 32, 64, 128, 256, 512, 1024):
        x = rng.standard_normal(M) + 1j * rng.standard_normal(M)
        assert np.allclose(stockham(x), np.fft.fft(x)), M
        assert np.allclose(stockham(x, True), np.fft.ifft(x) * M), M
    for n in (64, 256, 512):
        x = rng.standard_normal(n); w = np.blackman(n)
        assert np.allclose(rfft_half(x, w), np.fft.rfft(x * w))
        X = rng.standard_normal(n // 2 + 1) + 1j * rng.standard_normal(n // 2 + 1)
        assert np.allclose(irfft_half(X), np.fft.irfft(X))
    x = rng.standard_normal(256) + 1j * rng.standard_normal(256)
    out = fft256_16x16(x); ref = np.fft.fft(x)
    for k1 in range(16):
        assert np.allclose(out[k1], ref[k1::16])
    assert np.allclose(ifft256_16x16(out) / 256, x)
    # pair trick: a, b real -> A[k] = (Z[k] + conj Z[N-k]) / 2, B[k] = (Z[k] - conj Z[N-k]) / 2i
    a = rng.standard_normal(256); b = rng.standard_normal(256)
    Z = np.fft.fft(a + 1j * b); Zr = np.conj(Z[(-np.arange(256)) % 256])
    assert np.allclose((Z + Zr) / 2, np.fft.fft(a)) and np.allclose((Z - Zr) / 2j, np.fft.fft(b))
    # inverse pair: Y1, Y2 hermitian half spectra (imag of DC / Nyquist dropped) -> y1 + i y2
    Y1 = np.fft.rfft(a); Y2 = np.fft.rfft(b)
    full = lambda H: np.concatenate([H, np.conj(H[-2:0:-1])])
    z = np.fft.ifft(full(Y1) + 1j * full(Y2))
    assert np.allclose(z.real, a) and np.allclose(z.imag, b)
    check_lane_algebra(rng)
    print("fft emulation OK")


if __name__ == "__main__":
    main()
