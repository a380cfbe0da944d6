"""numpy emulation of the warp-level FFT dataflow used by pfann_b200/csrc/mel.cu.

1024-point real FFT per warp = 512-point complex FFT of z[n] = x[2n] + i x[2n+1]:
  lane l, register r hold n = 32 r + l
  (1) 16-point DFT over r inside each lane                    -> Y_l[k1]
  (2) twiddle by W_512^(l*k1)
  (3) 32-point radix-2 DIF across lanes (shfl_xor), output lane l holds k2 = bitrev5(l)
  (4) Z[k1 + 16 k2]; real-FFT post-processing X[k] = E + W_1024^k O
This script checks the index algebra against numpy.fft (run on CPU, no GPU needed)."""
import numpy as np


def bitrev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


def fft16_dif_inplace(a):
    """radix-2 DIF on axis 0 (16 regs), returns natural-order output via compile-time bit reversal."""
    a = a.copy()
    n = 16
    half = 8
    while half >= 1:
        for base in range(0, n, 2 * half):
            for j in range(half):
                w = np.exp(-2j * np.pi * j / (2 * half))
                u, v = a[base + j].copy(), a[base + j + half].copy()
                a[base + j] = u + v
                a[base + j + half] = (u - v) * w
        half //= 2
    out = np.empty_like(a)
    for i in range(n):
        out[bitrev(i, 4)] = a[i]
    return out


def warp_fft_1024_real(x):
    z = x[0::2] + 1j * x[1::2]                 # 512 complex
    lanes = np.arange(32)
    a = np.empty((16, 32), complex)            # a[r][l] = z[32 r + l]
    for r in range(16):
        a[r] = z[32 * r + lanes]
    Y = fft16_dif_inplace(a)                   # Y[k1][l]
    for k1 in range(16):
        Y[k1] *= np.exp(-2j * np.pi * lanes * k1 / 512)
    half = 16
    while half >= 1:                           # cross-lane DIF via shfl_xor
        partner = Y[:, lanes ^ half]
        upper = (lanes & half) != 0
        tw = np.exp(-2j * np.pi * (lanes & (half - 1)) / (2 * half))
        Y = np.where(upper[None, :], (partner - Y) * tw[None, :], Y + partner)
        half //= 2
    Z = np.empty(512, complex)
    for l in range(32):
        k2 = bitrev(l, 5)
        for k1 in range(16):
            Z[k1 + 16 * k2] = Y[k1, l]
    X = np.empty(513, complex)
    for k in range(513):
        A = Z[k % 512]
        B = np.conj(Z[(512 - k) % 512])
        E = 0.5 * (A + B)
        O = -0.5j * (A - B)
        X[k] = E + np.exp(-2j * np.pi * k / 1024) * O
    return X


def warp_fft_1024_real_v2(x):
    """Round-2 dataflow (mel_fast_kernel): the 32-point cross-lane part is ONE shared-memory transpose, a second
    in-lane 16-point DFT and a single shfl_xor(1) butterfly; the recombination forms bins k and 512-k from one pair."""
    z = x[0::2] + 1j * x[1::2]
    lanes = np.arange(32)
    a = np.empty((16, 32), complex)
    for r in range(16):
        a[r] = z[32 * r + lanes]
    Y = fft16_dif_inplace(a)                   # Y[k1][n2 = lane]
    for k1 in range(16):
        Y[k1] *= np.exp(-2j * np.pi * lanes * k1 / 512)
    # transpose through shared memory: buf[k1 * 34 + n2]; lane L reads k1 = L >> 1, n2 = (L & 1) + 2 j
    buf = np.zeros(16 * 34, complex)
    for k1 in range(16):
        buf[k1 * 34 + lanes] = Y[k1]
    u = np.empty((16, 32), complex)
    for j in range(16):
        u[j] = buf[(lanes >> 1) * 34 + (lanes & 1) + 2 * j]
    F = fft16_dif_inplace(u)                   # F[q][lane]: E (even lanes) / O (odd lanes) of the 32-point DFT
    b = lanes & 1
    Zbuf = np.zeros(552, complex)
    idx = lambda k: k + (k >> 4) + (k >> 8) * 8
    for q in range(16):
        w = np.where(b == 1, F[q] * np.exp(-2j * np.pi * q / 32), F[q])
        p = w[lanes ^ 1]
        res = np.where(b == 1, p - w, w + p)
        k = (lanes >> 1) + 16 * (q + 16 * b)
        Zbuf[idx(k)] = res
    P = np.zeros(513)
    for it in range(8):
        k = lanes + 32 * it
        A = Zbuf[idx(k)]
        B = np.conj(Zbuf[idx((512 - k) & 511)])
        E = 0.5 * (A + B)
        O = -0.5j * (A - B)
        T = np.exp(-2j * np.pi * k / 1024) * O
        P[k] = np.abs(E + T) ** 2
        P[512 - k] = np.abs(E - T) ** 2
    A = Zbuf[idx(256)]
    B = np.conj(A)
    P[256] = np.abs(0.5 * (A + B) + np.exp(-2j * np.pi * 256 / 1024) * (-0.5j) * (A - B)) ** 2
    return P


def transpose_read_banks():
    """Banks touched by the transpose reads of one half-warp (64-bit accesses are served per half-warp)."""
    out = []
    for half in (0, 16):
        L = np.arange(half, half + 16)
        words = 2 * ((L >> 1) * 34 + (L & 1))
        out.append(sorted(set(np.concatenate([words % 32, (words + 1) % 32]).tolist())))
    return out


if __name__ == '__main__':
    rng = np.random.default_rng(0)
    x = rng.standard_normal(1024)
    err = np.abs(np.fft.rfft(x) - warp_fft_1024_real(x)).max()
    print('v1 max err', err)
    assert err < 1e-10
    ref = np.abs(np.fft.rfft(x)) ** 2
    err = np.abs(ref - warp_fft_1024_real_v2(x)).max() / ref.max()
    print('v2 max rel err', err)
    assert err < 1e-12
    assert all(len(b) == 32 for b in transpose_read_banks())
