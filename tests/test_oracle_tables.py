"""CPU: code tables of SURVEY.md 8(a) rows a20-a24 for the MATLAB-level restatement
(the reference has no vectors for these; the tables were derived from the .m sources)."""
import numpy as np

import oracle as O


def bits(s):
    return [int(c) for c in s]


def test_eg0_table():
    # cabacBinarizer.m:56-69 -- EG0: 0->0, 1->100, 2->101, 3->11000, 7->1110000
    want = {0: "0", 1: "100", 2: "101", 3: "11000", 4: "11001", 5: "11010", 6: "11011", 7: "1110000"}
    for v, s in want.items():
        assert list(O.binarize(v, 8, O.BIN_EG0)) == bits(s)
        assert O.debinarize(bits(s), 8, O.BIN_EG0) == v


def test_egk_lengths_and_roundtrip():
    for k, m in ((0, O.BIN_EG0), (1, O.BIN_EG1), (2, O.BIN_EG2)):
        for v in list(range(0, 300)) + [1023, 1024, 65535, 2**31, 2**32 - 1]:
            b = O.binarize(v, 256, m)
            n_p = ((v >> k) + 1).bit_length()
            assert len(b) == 2 * n_p + k - 1
            assert O.debinarize(b, 256, m) == v


def test_tu():
    # cabacBinarizer.m:30-37: no terminating zero at v == Nq-1
    assert list(O.binarize(0, 4, O.BIN_TU)) == [0]
    assert list(O.binarize(2, 4, O.BIN_TU)) == [1, 1, 0]
    assert list(O.binarize(3, 4, O.BIN_TU)) == [1, 1, 1]
    for v in range(4):
        assert O.debinarize(O.binarize(v, 4, O.BIN_TU), 4, O.BIN_TU) == v


def test_fl32_and_tr():
    b = O.binarize(0xDEADBEEF, 0, O.BIN_FL32)
    assert len(b) == 32 and O.debinarize(b, 0, O.BIN_FL32) == 0xDEADBEEF
    # TR-k below the escape value: prefix floor(v/2^k) ones + 0, k-bit suffix (cabacBinarizer.m:39-54)
    assert list(O.binarize(5, 100, O.BIN_TR1)) == [1, 1, 0, 1]
    assert O.debinarize([1, 1, 0, 1], 100, O.BIN_TR1) == 5
    # at/above Nq-1 the suffix is left all-ones (upstream TODO)
    assert list(O.binarize(9, 10, O.BIN_TR2)) == [1, 1, 0, 1, 1]


def test_demo_ctx_rule():
    # cabacDemo.m:113-121
    assert O.select_ctx(O.PROFILE_DEMO, 1, [], []) == 0
    assert O.select_ctx(O.PROFILE_DEMO, 1, [], [1, 0]) == 1
    assert O.select_ctx(O.PROFILE_DEMO, 1, [], [0]) == 2
    assert O.select_ctx(O.PROFILE_DEMO, 2, [1], [0]) == 0


def test_iss_ctx_rule():
    # cabacContextSelection.m:24-67, Nlbp=3, default types; returned ids are MATLAB id-1
    T = O.CM_COND0 | O.CM_COND1 | O.CM_CONDS0 | O.CM_CONDS1
    N = 3
    sel = lambda n, g, up, t=T: O.select_ctx(O.PROFILE_ISS, n, g, up, N, t) + 1
    assert sel(1, [], []) == 1                      # no neighbour
    assert sel(1, [], [0]) == N + 1                 # up(1)==0, cond0
    assert sel(1, [], [1, 0, 1]) == 2 * N + 1       # up(1)==1, cond1
    assert sel(2, [1], [0]) == 2                    # up too short -> default (condbinlft off)
    assert sel(2, [1], [0], T | O.CM_CONDBINLFT) == 3 * N + 1
    assert sel(2, [1], [1, 0, 1]) == N + 2          # up(2)==0 in up's prefix
    assert sel(3, [1, 1], [1, 0, 1]) == 3           # position 3 is beyond up's prefix -> default
    assert sel(4, [1, 1, 1], [1, 1, 1, 0]) == 7 * N + 1   # rest prefix
    # suffix: own prefix ended at position 2 -> m = n-2
    assert sel(3, [1, 0], []) == 4 * N + 1
    assert sel(3, [1, 0], [1, 0, 0]) == 5 * N + 1   # up(3)==0 and beyond up's prefix
    assert sel(3, [1, 0], [1, 0, 1]) == 6 * N + 1
    assert sel(3, [1, 0], [1, 1, 0, 1, 1]) == 4 * N + 1   # position 3 still in up's prefix
    assert sel(6, [1, 0, 1, 1, 1], []) == 7 * N + 2       # m=4 > Nlbp
    assert O.num_ctx(O.PROFILE_ISS, 3) == 23


def test_flat_profiles():
    N = 3
    f = lambda n, g: O.select_ctx(O.PROFILE_FLAT, n, g, [], N)
    assert [f(1, []), f(2, [1]), f(3, [1, 1]), f(4, [1, 1, 1]), f(5, [1, 1, 1, 1])] == [0, 1, 2, 3, 3]
    assert [f(3, [1, 0]), f(4, [1, 0, 1]), f(5, [1, 0, 1, 1]), f(6, [1, 0, 1, 1, 0])] == [4, 5, 6, 7]
    e = lambda n, g: O.select_ctx(O.PROFILE_FLAT_EPSUF, n, g, [], N)
    assert e(2, [1]) == 1 and e(3, [1, 0]) == -1
    assert O.num_ctx(O.PROFILE_FLAT, 3) == 8 and O.num_ctx(O.PROFILE_FLAT_EPSUF, 3) == 4


def test_symbol_roundtrip_all_profiles():
    rng = np.random.default_rng(5)
    T = O.CM_COND0 | O.CM_COND1 | O.CM_CONDS0 | O.CM_CONDS1
    for prof, meth, Nq, rows in ((O.PROFILE_DEMO, O.BIN_TU, 4, 0), (O.PROFILE_DEMO, O.BIN_EG0, 4, 0),
                                 (O.PROFILE_ISS, O.BIN_EG0, 8, 37), (O.PROFILE_ISS, O.BIN_EG2, 64, 10),
                                 (O.PROFILE_FLAT, O.BIN_EG1, 16, 0), (O.PROFILE_FLAT_EPSUF, O.BIN_EG2, 256, 0),
                                 (O.PROFILE_ISS, O.BIN_TU, 8, 20)):
        cfg = O.make_cfg(prof, meth, Nq, 3, T, rows)
        counts = rng.integers(0, 200, size=12)
        counts[0] = 0
        off = np.zeros(13, dtype=np.uint64)
        np.cumsum(counts, out=off[1:])
        sym = rng.integers(0, Nq, size=int(off[-1])).astype(np.uint32)
        ci = np.zeros(O.num_ctx(prof, 3), dtype=np.uint8) + 1
        slab, lens, bits_after = O.encode_symbols(cfg, sym, off, ci, out_stride=2048, want_bits=True)
        payload, boff = O.compact(slab, lens)
        dec, ok = O.decode_symbols(cfg, payload, boff, off, ci)
        assert ok.all() and (dec == sym).all()
        assert lens[0] == 2 and bytes(slab[0, :2]) == bytes.fromhex("fe80")   # empty stream = K0
