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
    print("fft emulation OK")


if __name__ == "__main__":
    main()
