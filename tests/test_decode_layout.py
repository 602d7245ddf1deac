"""Index arithmetic of the one-kernel decode step's skinny GEMM (csrc/decode_step.cu::phase_gemm), emulated lane by lane
on the CPU: fragments are filled from 16-byte row pieces with the k index of a 32-wide block permuted identically for the
activations (A) and the weights (B), fed to mma.m16n8k16 (PTX ISA fragment layout, restated below), the 16 warps' partial
tiles are reduced and stored through the kernel's (tile, element) -> (row, column) map.  The result must be A @ W^T."""
import numpy as np

WARPS, NKB = 16, 2


def mma_m16n8k16(a, b):
    """a[lane][4 regs][2 halves], b[lane][2][2] -> c[lane][4]; PTX: row = g (+8 for a1, a3 / c2, c3), k = 2q + half
    (+8 for a2, a3 / b1), n = g for B, column = 2q + (j & 1) for C; g = lane / 4, q = lane % 4."""
    A = np.zeros((16, 16)); B = np.zeros((16, 8))
    for lane in range(32):
        g, q = lane // 4, lane % 4
        for r in range(4):
            for h in range(2):
                A[g + 8 * (r & 1), 2 * q + h + 8 * (r >> 1)] = a[lane][r][h]
        for r in range(2):
            for h in range(2):
                B[2 * q + h + 8 * r, g] = b[lane][r][h]
    C = A @ B
    c = np.zeros((32, 4))
    for lane in range(32):
        g, q = lane // 4, lane % 4
        for j in range(4):
            c[lane][j] = C[g + 8 * (j >> 1), 2 * q + (j & 1)]
    return c


def emulate_tile(Amat, W, n0):
    """one 8-column unit: returns out[32, 8] the way the kernel's reduction thread e = mt*128 + lane*4 + j stores it"""
    K = Amat.shape[1]
    red = np.zeros((WARPS, 256))
    for warp in range(WARPS):
        acc = np.zeros((2, 32, 4))
        for kb in range(NKB):
            kblk = warp * NKB + kb
            if kblk * 32 >= K:
                continue
            for mt in range(2):
                for j in range(2):                      # the two mma of a 32-wide k block
                    a = np.zeros((32, 4, 2)); b = np.zeros((32, 2, 2))
                    for lane in range(32):
                        g, q = lane // 4, lane % 4
                        lo = Amat[mt * 16 + g, kblk * 32 + q * 8: kblk * 32 + q * 8 + 8]
                        hi = Amat[mt * 16 + g + 8, kblk * 32 + q * 8: kblk * 32 + q * 8 + 8]
                        w = W[n0 + g, kblk * 32 + q * 8: kblk * 32 + q * 8 + 8]
                        # af[kb][mt][j] = {lo pair 2j, hi pair 2j, lo pair 2j+1, hi pair 2j+1}; b = {w pair 2j, w pair 2j+1}
                        a[lane] = [lo[4 * j: 4 * j + 2], hi[4 * j: 4 * j + 2], lo[4 * j + 2: 4 * j + 4], hi[4 * j + 2: 4 * j + 4]]
                        b[lane] = [w[4 * j: 4 * j + 2], w[4 * j + 2: 4 * j + 4]]
                    acc[mt] += mma_m16n8k16(a, b)
        for mt in range(2):
            for lane in range(32):
                red[warp, mt * 128 + lane * 4: mt * 128 + lane * 4 + 4] = acc[mt, lane]
    out = np.zeros((32, 8))
    for idx in range(256):
        mt, ln, j = idx >> 7, (idx >> 2) & 31, idx & 3
        row, col = mt * 16 + (ln >> 2) + ((j >> 1) << 3), (ln & 3) * 2 + (j & 1)
        out[row, col] = red[:, idx].sum()
    return out


def test_fragment_permutation_and_output_map_give_the_plain_product():
    rng = np.random.default_rng(0)
    for K in (128, 768, 1024):
        A = rng.standard_normal((32, K)); W = rng.standard_normal((24, K))
        for n0 in (0, 16):
            np.testing.assert_allclose(emulate_tile(A, W, n0), A @ W[n0:n0 + 8].T, rtol=1e-10, atol=1e-10)
