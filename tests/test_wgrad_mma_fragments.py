"""Fragment-level statement of csrc/lora_wgrad_mma.cu on CPU: the ldmatrix.x4.trans address assignment and the m16n8k16 A / B / D
fragment ownership the kernel relies on (PTX ISA: matrix fragments for mma.m16n8k16 with 16-bit types; ldmatrix .trans), written
out lane by lane in numpy and checked to give out = P^T Q for one staged chunk.  It pins the index arithmetic of the kernel
(arow / acol / brow / bcol and the D write-out); the instruction semantics themselves are the ones the GPU-validated decode
attention kernel (attention.cu: ldsm_x4_trans + mma16816) already depends on.  TEST INFRASTRUCTURE ONLY."""
import numpy as np

TOK, FEAT, RANK = 64, 64, 16


def _ldsm_x4_trans(tile, row_of_lane, col_of_lane):
    """Lane l supplies the address of row l % 8 of matrix l // 8 (8 b16 elements); with .trans thread t receives, from matrix
    i, the elements [2 (t % 4)][t / 4] and [2 (t % 4) + 1][t / 4] in register i."""
    mats = [np.stack([tile[row_of_lane[8 * i + r], col_of_lane[8 * i + r]: col_of_lane[8 * i + r] + 8] for r in range(8)]) for i in range(4)]
    regs = np.zeros((32, 4, 2))
    for t in range(32):
        for i in range(4):
            regs[t, i, 0], regs[t, i, 1] = mats[i][2 * (t % 4)][t // 4], mats[i][2 * (t % 4) + 1][t // 4]
    return regs


def _mma(a, b0, b1, c):
    A, B = np.zeros((16, 16)), np.zeros((16, 8))
    for t in range(32):
        g, q = t // 4, t % 4
        A[g, 2 * q: 2 * q + 2], A[g + 8, 2 * q: 2 * q + 2] = a[t, 0], a[t, 1]
        A[g, 2 * q + 8: 2 * q + 10], A[g + 8, 2 * q + 8: 2 * q + 10] = a[t, 2], a[t, 3]
        B[2 * q: 2 * q + 2, g], B[2 * q + 8: 2 * q + 10, g] = b0[t], b1[t]
    D = A @ B
    for t in range(32):
        g, q = t // 4, t % 4
        c[t] += [D[g, 2 * q], D[g, 2 * q + 1], D[g + 8, 2 * q], D[g + 8, 2 * q + 1]]


def test_one_chunk_gives_p_transposed_times_q():
    rng = np.random.default_rng(0)
    Ps, Qs = rng.standard_normal((TOK, FEAT)), rng.standard_normal((TOK, RANK))
    out = np.zeros((FEAT, RANK))
    lanes = np.arange(32)
    for warp in range(4):
        acc = np.zeros((2, 32, 4))
        for ks in range(TOK // 16):
            arow, acol = ks * 16 + (lanes & 7) + 8 * (lanes >> 4), warp * 16 + 8 * ((lanes >> 3) & 1)      # as in the kernel
            brow, bcol = ks * 16 + (lanes & 7) + 8 * ((lanes >> 3) & 1), 8 * (lanes >> 4)
            a, b = _ldsm_x4_trans(Ps, arow, acol), _ldsm_x4_trans(Qs, brow, bcol)
            _mma(a, b[:, 0], b[:, 1], acc[0])
            _mma(a, b[:, 2], b[:, 3], acc[1])
        for nb in range(2):
            for t in range(32):
                g, q = t // 4, t % 4
                for e in range(4):
                    out[warp * 16 + g + (8 if e >= 2 else 0), nb * 8 + q * 2 + (e & 1)] += acc[nb, t, e]
    assert np.abs(out - Ps.T @ Qs).max() < 1e-12
