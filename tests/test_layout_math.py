"""Host-side checks of index algebra the CUDA kernels rely on (no GPU, no library call): the dense centred-DFT formula of
``csrc/fftprox_any.cuh`` and the XOR-swizzled exchange layouts of the 256x256 cluster kernel (``csrc/fftprox_cl.cuh``:
``cl_q_idx``, ``cl_a_idx``, ``fft256_row_swz``), restated in Python exactly as the kernels index them."""
import numpy as np
import pytest


@pytest.mark.parametrize("N", [2, 3, 5, 16, 45, 130, 136])
def test_dense_centred_dft_formula_equals_shifted_fft(N):
    """fft_c(w)[k] = N^-1/2 sum_m w[m] omega^((m - h)(k - h)), h = N // 2, for even AND odd N (transformations.py:6-19)."""
    rng = np.random.default_rng(N)
    w = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    h = N // 2
    m = np.arange(N)
    for sign, ref in ((-1, np.fft.fftshift(np.fft.fft(np.fft.ifftshift(w), norm="ortho"))),
                      (+1, np.fft.fftshift(np.fft.ifft(np.fft.ifftshift(w), norm="ortho")))):
        # the kernel walks the exponent incrementally modulo N from ((N - h) % N * step) % N with step = (k - h) mod N
        out = np.empty(N, dtype=complex)
        for k in range(N):
            step = (k - h) % N
            ex = ((N - h) % N * step) % N
            e = (ex + step * m) % N
            assert np.array_equal(e, ((m - h) * (k - h)) % N)
            out[k] = np.sum(w * np.exp(sign * 2j * np.pi * e / N)) / np.sqrt(N)
        assert np.abs(out - ref).max() < 1e-12


def cl_q_idx(i, c):
    return 16 * i + (c ^ (i & 15) ^ (i >> 4))


def cl_a_idx(rho, s, cc):
    return 256 * s + 16 * cc + (rho ^ cc)


def _bank_pairs(addrs):
    """float2 elements: 16 consecutive elements cover the 32 banks once."""
    return [a % 16 for a in addrs]


def test_swizzled_q_layout_is_a_bijection_and_conflict_free():
    idx = {cl_q_idx(i, c) for i in range(256) for c in range(16)}
    assert idx == set(range(4096))
    for c in range(16):
        for r in range(16):
            # pass 1 / 3: lanes jc read image rows jc + 16 r; pass 2: rows r + 16 jc
            assert sorted(_bank_pairs([cl_q_idx(jc + 16 * r, c) for jc in range(16)])) == list(range(16))
            assert sorted(_bank_pairs([cl_q_idx(r + 16 * jc, c) for jc in range(16)])) == list(range(16))
    # exchange 1: the 16 lanes j of a row write one 128-byte segment (16 consecutive elements, permuted)
    for i in range(256):
        seg = sorted(cl_q_idx(i, j) for j in range(16))
        assert seg == list(range(16 * i, 16 * i + 16))


def test_swizzled_a_layout_sends_loads_and_row_exchange():
    idx = {cl_a_idx(rho, s, cc) for rho in range(16) for s in range(16) for cc in range(16)}
    assert idx == set(range(4096))
    for s in range(16):
        for cc in range(16):
            # exchange 2: lanes jc (= local row) of a half-warp write one permuted 128-byte segment
            seg = sorted(cl_a_idx(jc, s, cc) for jc in range(16))
            assert seg == list(range(256 * s + 16 * cc, 256 * s + 16 * cc + 16))
    for hw in range(16):
        own = {cl_a_idx(hw, s, cc) for s in range(16) for cc in range(16)}      # the row's 256 slots
        for r in range(16):
            # inverse rows: lanes j load element (row hw, col j + 16 r) = block r, cc = j
            assert sorted(_bank_pairs([cl_a_idx(hw, r, j) for j in range(16)])) == list(range(16))
        # radix-16 exchange inside the row's own slots: lane j writes V_j[q] to (s = q, cc = j ^ q), lane j reads V_r[j]
        # from (s = j, cc = r ^ j)
        written = {}
        for j in range(16):
            for q in range(16):
                a = cl_a_idx(hw, q, j ^ q)
                assert a in own and a not in written
                written[a] = (j, q)
        for q in range(16):
            assert sorted(_bank_pairs([cl_a_idx(hw, q, j ^ q) for j in range(16)])) == list(range(16))
        for j in range(16):
            for r in range(16):
                assert written[cl_a_idx(hw, j, r ^ j)] == (r, j)                 # reader j gets V_r[j]
        for r in range(16):
            assert sorted(_bank_pairs([cl_a_idx(hw, j, r ^ j) for j in range(16)])) == list(range(16))
