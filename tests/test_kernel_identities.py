"""Exactness of the integer identities the CUDA kernels lean on, checked over their whole input range (CPU, numpy).

Kernel A1 (csrc/smooth_sobel.cu) replaces the reference's multiply-high thirds / ninths ((s * 21846) >> 16 of
filter.hpp:304,332-371, (s * 7282) >> 16 of filter.hpp:416,466-471) by BYTE 2 of a plain 32-bit product, and the Sobel response
|A-B|^2 + |C-D|^2 by a byte-wise absolute difference followed by the dot product of the word with itself.
Kernel A2 (csrc/hash_tiles.cu) moves a test's four flags from bit 7 of their bytes to bit p with a multiply-high by
2^(25+p).  These tests pin the arithmetic; the kernels themselves are pinned by the GPU parity tests."""
import numpy as np


def test_third_is_byte2_of_plain_product():
    s = np.arange(0, 766, dtype=np.uint64)                 # three bytes summed
    prod = (s * 21846) & 0xFFFFFFFF
    assert np.all(prod < (1 << 24))                        # byte 3 = 0: the PRMT that picks byte 3 reads a zero
    assert np.array_equal((prod >> 16) & 0xFF, (s * 21846) >> 16)
    assert np.all(((s * 21846) >> 16) <= 255)


def test_ninth_is_byte2_of_plain_product():
    s = np.arange(0, 1021, dtype=np.uint64)                # [1 2 1] sums of bytes
    prod = (s * 7282) & 0xFFFFFFFF
    assert np.all(prod < (1 << 24))
    assert np.array_equal((prod >> 16) & 0xFF, (s * 7282) >> 16)
    assert ((1020 * 7282) >> 16) == 113                    # so 2 * 113^2 = 25538 fits the int compare


def test_sobel_response_from_packed_bytes():
    rng = np.random.default_rng(7)
    a, b, c, d = (rng.integers(0, 114, 200000, dtype=np.int64) for _ in range(4))
    # z = VABSDIFF4((a, c, 0, 0), (b, d, 0, 0)); dp4a(z, z) = sum of the squared bytes
    z0, z1 = np.abs(a - b), np.abs(c - d)
    assert np.array_equal(z0 * z0 + z1 * z1, (a - b) ** 2 + (c - d) ** 2)
    assert (z0 * z0 + z1 * z1).max() <= 25538


def test_flag_shift_by_multiply_high():
    rng = np.random.default_rng(11)
    flags = rng.integers(0, 16, 4096, dtype=np.uint64)     # which of the four pixels pass the test
    word = sum(((flags >> j) & 1) << (8 * j + 7) for j in range(4))
    for p in range(7):                                     # p = 7 is a plain addition in the kernel
        hi = (word * (1 << (25 + p))) >> 32                # IMAD.HI by the constant-bank multiplier pmul[t]
        want = sum(((flags >> j) & 1) << (8 * j + p) for j in range(4))
        assert np.array_equal(hi, want)
    # eight tests of one state byte never carry into each other: the accumulated byte is just their flags
    acc = np.zeros_like(word)
    bits = rng.integers(0, 16, (8, 4096), dtype=np.uint64)
    for p in range(8):
        w = sum(((bits[p] >> j) & 1) << (8 * j + 7) for j in range(4))
        acc += (w * (1 << (25 + p))) >> 32 if p < 7 else w
    for j in range(4):
        byte = (acc >> (8 * j)) & 0xFF
        want = sum(((bits[p] >> j) & 1) << p for p in range(8))
        assert np.array_equal(byte, want)
