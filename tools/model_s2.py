"""numpy model of the packed ("S2") 512-point real FFT dataflow used by fbank_quad (fbank_tile.cuh): per 16-thread
group, thread j owns samples n = 16*row + j, runs a real 32-point FFT (half-size complex FFT + split), one transpose,
then a complex 16-point FFT on its column.  Checks the index maps, the rotation of the second group and the scales
against numpy's rfft.  Development aid only (not imported by the package or the tests)."""
import numpy as np

W = lambda n, k: np.exp(-2j * np.pi * k / n)


def group_power(x, g):
    """x: [400] windowed frame (one SIMD lane).  Returns 4*|rfft(x, 512)|^2 for bins 0..255 as the kernel stores it."""
    L = 400
    xg = np.zeros((17, 16), complex)       # rows 1..16: Y2[k1] of thread j ; row 0: real Y[0]
    for j in range(16):
        y = np.zeros(32)
        for i in range(32):
            row = i - g
            if 0 <= row < 25:
                y[i] = x[16 * row + j]
        z = y[0::2] + 1j * y[1::2]
        Z = np.fft.fft(z)                  # 16-point
        for k in range(1, 8):
            A, B = Z[k], Z[16 - k]
            S = A + np.conj(B)
            O2 = complex(A.imag + B.imag, B.real - A.real)
            T = W(32, k) * O2
            xg[k, j] = S + T
            xg[16 - k, j] = np.conj(S - T)
        xg[8, j] = np.conj(Z[8])
        xg[16, j] = Z[0].real - Z[0].imag
        xg[0, j] = Z[0].real + Z[0].imag
    tw = np.zeros((17, 16), complex)
    for k1 in range(1, 17):
        for c in range(16):
            # one table for both groups: the rotated group's column k1 is off by the unit-modulus factor W32^k1,
            # which the power spectrum does not see
            tw[k1, c] = W(512, c * k1) * (2.0 if k1 in (8, 16) else 1.0)
    P = np.full(257, np.nan)
    for j in range(16):
        col = 16 if j == 0 else j
        z2 = xg[col, :] * tw[col, :]
        X = np.fft.fft(z2)
        for k2 in range(16):
            if k2 < 8:
                P[col + 32 * k2] = abs(X[k2]) ** 2
            elif j != 0:
                P[(32 - col) + 32 * (15 - k2)] = abs(X[k2]) ** 2
    # column 0: a 16-point FFT of the real values Y_j[0] ACROSS the 16 lanes of the group (B200FE_C0_SHFL, quad_stage2):
    # decimation in frequency over the lane index with xor distances 8, 4, 2, 1 - a lane whose bit D is clear keeps
    # a + b, its partner (a - b) * W_2D^(j mod D); stage 0 carries the factor 2, stage 3 has no twiddle.  Lane j ends up
    # with bin 32 * bitrev4(j); the even lanes hold bins 0 .. 224.
    z = xg[0, :].real.astype(complex)
    for s, D in enumerate((8, 4, 2, 1)):
        nz = np.empty(16, complex)
        for j in range(16):
            p = z[j ^ D]
            upper = (j & D) != 0
            d = p - z[j] if upper else p + z[j]
            w = W(2 * D, j & (D - 1)) if (upper and D > 1) else 1.0
            nz[j] = d * w * (2.0 if s == 0 else 1.0)
        z = nz
    for j in range(0, 16, 2):
        t = int(format(j, "04b")[::-1], 2)
        P[32 * t] = abs(z[j]) ** 2
    return P[:256]


rng = np.random.default_rng(0)
for g in (0, 1):
    x = rng.standard_normal(400)
    ref = 4.0 * np.abs(np.fft.rfft(x, 512)) ** 2
    got = group_power(x, g)
    assert not np.isnan(got).any()
    err = np.max(np.abs(got - ref[:256]) / ref[:256])
    print("group", g, "max rel err", err)
    assert err < 1e-10
print("ok")
