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


def _used_bits(n: int, full: bool) -> int:
    """descriptor bits the transform can set (LIMITED 4n-6 for n >= 4, 4 and 7 bits for n = 2, 3)"""
    if full:
        return n * n - 2 * n + 3
    return {2: 4, 3: 7}.get(n, 4 * n - 6)


def test_top_descriptor_bit_is_never_used():
    """What `top_bit_free` (bicos_b200_match -> launch_search) promises: for every stack size the reference accepts,
    and for the wide-descriptor extension, the descriptor never fills its last word completely."""
    import oracle

    for full, sizes, wide in ((False, range(2, 66), False), (True, range(2, 17), False), (True, range(17, 24), True)):
        for n in sizes:
            k = oracle.words_per_descriptor(n, full, wide)
            assert _used_bits(n, full) <= 32 * k - 1, (n, full, k)


@pytest.mark.parametrize("n,dtype,full,wide", [(33, np.uint8, False, False), (65, np.uint8, False, False),
                                               (12, np.uint8, True, False), (16, np.uint16, True, False),
                                               (20, np.uint16, True, True)])
def test_oracle_descriptors_leave_the_top_bit_clear(oracles, n, dtype, full, wide):
    from libbicos_b200 import synth

    left, right, _ = synth.make_stacks(n, 64, 96, dtype, seed=3 * n)
    for stack in (left, right):
        d = oracles.port.descriptors(stack, full, wide)
        assert d.shape[-1] in (4, 8, 12, 16)
        assert not (d[..., -1] >> 31).any()


@pytest.mark.parametrize("k", [4, 8])
def test_column_term_rides_in_the_top_bits_byte(k):
    """With bit 32K-1 clear on both sides the left operand's byte for it is +1; putting the tile column u into that
    byte of the right operand makes the GEMM deliver acc + u (fold32<CT> / fold64_packed<CT> in search_mma.cu)."""
    rng = np.random.default_rng(100 + k)
    left = rng.integers(0, 2**32, size=(64, k), dtype=np.uint64).astype(np.uint32)
    right = rng.integers(0, 2**32, size=(TN, k), dtype=np.uint64).astype(np.uint32)
    left[:, -1] &= 0x7FFFFFFF
    right[:, -1] &= 0x7FFFFFFF
    a, b = operands(left, True), operands(right, False)
    top = 32 * k - 1  # k position of (last word, s = 7, byte 3) in the (word, s, byte) order
    assert np.all(a[:, top] == 1) and np.all(b[:, top] == 0)
    b[:, top] = np.arange(TN)  # what the producers OR in: the row's tile column, < 128, an unsigned byte
    acc = a @ b.T
    ham = popcount(left[:, None, :] ^ right[None, :, :])
    assert np.array_equal(acc, 128 * (ham - popcount(left)[:, None]) + np.arange(TN)[None, :])
    if k == 4:
        assert np.abs(acc).max() + 127 < 2**15  # the last-minimum key adds 127 - 2u on top


# ---- the one-pass consistency kernel (search_mma3_kernel): both directions from one accumulator ----

def test_top_two_descriptor_bits_are_never_used():
    """What free_top_bits = 2 (bicos_b200_match -> launch_search) promises for the one-pass kernel: bits 32K-1 and
    32K-2 stay clear for every stack size the reference accepts."""
    import oracle

    for full, sizes in ((False, range(2, 66)), (True, range(2, 17))):
        for n in sizes:
            k = oracle.words_per_descriptor(n, full, False)
            assert _used_bits(n, full) <= 32 * k - 2, (n, full, k)


def _onepass_operands(left: np.ndarray, right_block: np.ndarray):
    """A = the streamed LEFT tile as unsigned bytes (expand_moving_to_tmem), B = the RIGHT block as signed bytes
    (expand_block_pixel); the bytes of the two unused top bits (word 3, byte 3, s = 6 / 7) carry 128 x popc(right) and
    1 x block column."""
    k = left.shape[1]
    a, b = operands(left, False), operands(right_block, True)  # unsigned / signed encodings
    pos6, pos7 = 32 * k - 5, 32 * k - 1  # (last word, s = 6, byte 3) and (last word, s = 7, byte 3) in (word, s, byte) order
    assert np.all(a[:, pos6] == 0) and np.all(a[:, pos7] == 0) and np.all(b[:, pos6] == 2) and np.all(b[:, pos7] == 1)
    a[:, pos6], a[:, pos7] = 128, 1
    b[:, pos6], b[:, pos7] = popcount(right_block), np.arange(len(right_block))
    assert a.min() >= 0 and a.max() <= 255 and b.min() >= -128 and b.max() <= 127  # u8 x s8
    return a, b


def test_onepass_accumulator_is_128_hamming_plus_block_column():
    rng = np.random.default_rng(5)
    left = rng.integers(0, 2**32, size=(128, 4), dtype=np.uint64).astype(np.uint32)
    right = rng.integers(0, 2**32, size=(TN, 4), dtype=np.uint64).astype(np.uint32)
    left[:, -1] &= 0x3FFFFFFF
    right[:, -1] &= 0x3FFFFFFF
    left[0], right[0] = 0, 0
    left[1, :3], right[1, :3] = 0xFFFFFFFF, 0xFFFFFFFF
    left[1, 3], right[2, 3] = 0x3FFFFFFF, 0x3FFFFFFF  # 126 set bits: the largest popcount byte
    a, b = _onepass_operands(left, right)
    acc = a @ b.T  # [left pixel = TMEM lane, right pixel = accumulator column]
    ham = popcount(left[:, None, :] ^ right[None, :, :])
    assert np.array_equal(acc, 128 * ham + np.arange(TN)[None, :])
    assert acc.min() >= 0 and acc.max() + 127 < 2**15  # + tile index: still a positive signed half word


@pytest.mark.parametrize("cols", [1, 97, 128, 300, 1000])
def test_onepass_folds_give_both_first_minima(cols):
    """Forward: per left pixel the minimum of 128 ham + block column over the columns of each block, merged over
    the blocks as cost << 16 | column (atomicMin). Reverse: per right pixel the ELEMENTWISE minimum of
    128 ham + column + tile over the left tiles (per TMEM lane), then the minimum over the lanes of
    (that << 16) + lane: decodes to cost << 16 | first left column."""
    rng = np.random.default_rng(1000 + cols)
    pool = rng.integers(0, 2**32, size=(10, 4), dtype=np.uint64).astype(np.uint32)
    pool[:, -1] &= 0x3FFFFFFF
    left = pool[rng.integers(0, len(pool), size=cols)]
    right = pool[rng.integers(0, len(pool), size=cols)] ^ ((rng.integers(0, 3, size=(cols, 1)) == 0) * np.uint32(1 << 5))
    ham = popcount(left[:, None, :] ^ right[None, :, :]).astype(np.int64)  # [left, right]
    tiles = (cols + TN - 1) // TN

    def padded(d, t):  # a ragged tile / block repeats its last pixel
        idx = np.minimum(np.arange(t * TN, (t + 1) * TN), cols - 1)
        return d[idx]

    fwd = np.full(cols, 0xFFFFFFFF, dtype=np.int64)
    rev = np.zeros(cols, dtype=np.int64)
    for nb in range(tiles):
        running = np.full((TN, TN), 0x7FFF, dtype=np.int64)  # [lane, block column], one row per epilogue thread
        for t in range(tiles):
            a, b = _onepass_operands(padded(left, t), padded(right, nb))
            acc = (a @ b.T).astype(np.int64)
            v = acc.min(axis=1)  # in-thread fold: 128 ham + first column at that cost
            i = t * TN + np.arange(TN)
            key = ((v >> 7) << 16) | (nb * TN + (v & 127))
            ok = i < cols
            fwd[i[ok]] = np.minimum(fwd[i[ok]], key[ok])
            running = np.minimum(running, acc + t)
        k = ((running << 16) + np.arange(TN)[:, None]).min(axis=0)  # over the lanes
        v = (k >> 16) - np.arange(TN)
        col = nb * TN + np.arange(TN)
        ok = col < cols
        rev[col[ok]] = (((v >> 7) << 16) | ((v & 127) * TN + (k & 0xFFFF)))[ok]

    best_f, best_r = ham.min(axis=1), ham.min(axis=0)
    assert np.array_equal(fwd, (best_f << 16) | (ham == best_f[:, None]).argmax(axis=1))
    assert np.array_equal(rev, (best_r << 16) | (ham == best_r[None, :]).argmax(axis=0))


def test_onepass_accumulator_256_bits_carries_the_popcount_with_a_bias():
    """256-bit descriptors: popc(right) reaches 254 and no longer fits the signed operand byte, so the byte carries
    popc - 128 and the accumulator is 128 ham + column - 16384: still a signed half word; adding 16384 + tile in the
    reverse fold keeps the running minima in [0, 32767]."""
    rng = np.random.default_rng(8)
    left = rng.integers(0, 2**32, size=(128, 8), dtype=np.uint64).astype(np.uint32)
    right = rng.integers(0, 2**32, size=(TN, 8), dtype=np.uint64).astype(np.uint32)
    left[:, -1] &= 0x3FFFFFFF
    right[:, -1] &= 0x3FFFFFFF
    left[0], right[0] = 0, 0
    left[1, :7], right[1, :7] = 0xFFFFFFFF, 0xFFFFFFFF
    left[1, 7], right[2, 7] = 0x3FFFFFFF, 0x3FFFFFFF  # 254 set bits on one side, none on the other: ham = 254
    right[2, :7] = 0xFFFFFFFF
    a, b = operands(left, False), operands(right, True)
    pos6, pos7 = 32 * 8 - 5, 32 * 8 - 1
    a[:, pos6], a[:, pos7] = 128, 1
    b[:, pos6], b[:, pos7] = popcount(right) - 128, np.arange(TN)
    assert b.min() >= -128 and b.max() <= 127
    acc = a @ b.T
    ham = popcount(left[:, None, :] ^ right[None, :, :])
    assert ham.max() == 254
    assert np.array_equal(acc, 128 * ham + np.arange(TN)[None, :] - 16384)
    assert acc.min() >= -(2**15) and acc.max() < 2**15
    assert (acc + 16384 + 63).max() < 2**15 and (acc + 16384).min() >= 0  # + tile index (rows of up to 8192 pixels)
