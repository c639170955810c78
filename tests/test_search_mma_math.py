"""The algebra of the tensor-core search engine (libbicos_b200/csrc/search_mma.cu), restated in numpy.

No GPU: these tests pin WHY the int8 GEMM + min epilogue is exact, step by step as the kernel does it:
operand encoding (expand_word), accumulator identity acc = 128 * (ham - popc(left)), the 16-bit range for
128-bit descriptors, per-tile keys acc + u / acc + 127 - u, widen_key and the tile merge, and the final keys
cost << 16 | column (first minimum) and cost << 16 | 65535 - column (last minimum) that the reference's
bicos_search (include/impl/cpu/bicos.hpp:50-76) implies. The GPU tests compare the kernel itself with the
popcount engine and the oracle.
"""

import numpy as np
import pytest

TN = 128
COL_BITS = 13
COL_MAX = (1 << COL_BITS) - 1


def expand_right(w: np.ndarray, s: int) -> np.ndarray:
    """four unsigned bytes per descriptor word: bit 8i+s -> b * 2^s (s = 0: b * 128)"""
    w = w.astype(np.uint32)
    x = ((w << np.uint32(7)) & np.uint32(0x80808080)) if s == 0 else (w & np.uint32(0x01010101 << s))
    return x[..., None].view(np.uint8)  # little endian: byte i of the word


def expand_left(w: np.ndarray, s: int) -> np.ndarray:
    """four signed bytes per descriptor word: bit 8i+s -> (1 - 2a) * 2^(7-s) (s = 0: +-1)"""
    p = 0 if s == 0 else 7 - s
    mag = np.uint32(0x01010101 << p)
    high = np.uint32(0x01010101 * ((0xFF << (p + 1)) & 0xFF))
    neg = ((w.astype(np.uint32) >> np.uint32(s)) & np.uint32(0x01010101)) * np.uint32(0xFF)
    return ((neg & high) | mag)[..., None].view(np.int8)


def operands(desc: np.ndarray, left: bool) -> np.ndarray:
    """[pixels, K] uint32 -> [pixels, 32 K] operand bytes in the kernel's k order (word, s, byte)"""
    f = expand_left if left else expand_right
    parts = [f(desc, s) for s in range(8)]  # each [pixels, K, 4]
    return np.stack(parts, axis=2).reshape(desc.shape[0], -1).astype(np.int32)


def popcount(a: np.ndarray) -> np.ndarray:
    return np.unpackbits(np.ascontiguousarray(a).view(np.uint8), axis=-1).sum(axis=-1).astype(np.int64)


def widen_key(t: np.ndarray) -> np.ndarray:
    t = t.astype(np.int64)
    return ((t & ~np.int64(127)) << 6) + (t & 127)


@pytest.mark.parametrize("k", [4, 8, 12, 16])
def test_accumulator_is_128_times_hamming_minus_popcount(k):
    rng = np.random.default_rng(k)
    left = rng.integers(0, 2**32, size=(96, k), dtype=np.uint64).astype(np.uint32)
    right = rng.integers(0, 2**32, size=(160, k), dtype=np.uint64).astype(np.uint32)
    # extremes: all zero / all one descriptors on both sides
    left[0], left[1], right[0], right[1] = 0, 0xFFFFFFFF, 0, 0xFFFFFFFF
    a, b = operands(left, True), operands(right, False)
    assert a.min() >= -128 and a.max() <= 127 and b.min() >= 0 and b.max() <= 255  # s8 x u8
    acc = a @ b.T
    ham = popcount(left[:, None, :] ^ right[None, :, :])
    assert np.array_equal(acc, 128 * (ham - popcount(left)[:, None]))
    if k == 4:  # the packed 16-bit epilogue: acc + u and acc + 127 - u fit a signed half word
        assert np.abs(acc).max() + 127 < 2**15


@pytest.mark.parametrize("cols", [1, 97, 128, 300, 1000])
def test_tile_keys_merge_to_first_and_last_minimum(cols):
    rng = np.random.default_rng(cols)
    k = 4
    pool = rng.integers(0, 2**32, size=(12, k), dtype=np.uint64).astype(np.uint32)  # few distinct values: many ties
    left = pool[rng.integers(0, len(pool), size=64)]
    right = pool[rng.integers(0, len(pool), size=cols)]
    right ^= (rng.integers(0, 3, size=(cols, 1)) == 0) * np.uint32(1 << 5)
    ham = popcount(left[:, None, :] ^ right[None, :, :]).astype(np.int64)
    pa = popcount(left).astype(np.int64)
    acc = 128 * (ham - pa[:, None])

    m_first = np.full(len(left), np.iinfo(np.int32).max, dtype=np.int64)
    m_last = m_first.copy()
    for tile0 in range(0, cols, TN):
        u = np.arange(min(TN, cols - tile0))
        a = acc[:, tile0:tile0 + len(u)]
        m_first = np.minimum(m_first, widen_key((a + u).min(axis=1)) + tile0)
        m_last = np.minimum(m_last, widen_key((a + 127 - u).min(axis=1)) + (COL_MAX - 127 - tile0))
    key_first = ((pa + (m_first >> COL_BITS)) << 16) | (m_first & COL_MAX)
    key_last = ((pa + (m_last >> COL_BITS)) << 16) | ((65535 - COL_MAX) + (m_last & COL_MAX))

    best = ham.min(axis=1)
    first = (ham == best[:, None]).argmax(axis=1)
    last = cols - 1 - (ham[:, ::-1] == best[:, None]).argmax(axis=1)
    assert np.array_equal(key_first, (best << 16) | first)
    assert np.array_equal(key_last, (best << 16) | (65535 - last))
