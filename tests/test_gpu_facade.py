"""GPU: the single-stream surfaces -- MEX command dispatcher, cabacWrapper mirror, handle API --
against the transcript of the UNMODIFIED reference mexFunction (tests/golden/mex_session.json)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    assert torch.cuda.is_available()
    import isscabac_b200.matlab_api as M
    return M


def test_replay_reference_mex_transcript(M, golden_dir, tmp_path):
    """Same commands -> same outputs, same error texts, same file bytes as the reference MEX."""
    with open(os.path.join(golden_dir, "mex_session.json")) as f:
        log = json.load(f)
    fn = str(tmp_path / "mex.bin")
    handle = None
    n_checked = 0
    for e in log:
        if "file" in e:
            assert open(fn, "rb").read().hex() == e["file"]
            continue
        args = []
        for a in e["args"]:
            if a == "<fn>":
                args.append(fn)
            elif a == "<handle>":
                args.append([handle])
            elif isinstance(a, str):
                args.append(a)
            else:
                args.append(np.array(a, dtype=np.float64))
        try:
            if not args:
                rc_out = M.SimpleCABACMex.__wrapped__ if False else None
                # nrhs == 0: call the dispatcher directly
                import isscabac_b200._lib as L
                err = C.create_string_buffer(512)
                out_n = C.c_int(0)
                rc = L.lib().simplecabac_dispatch(0, None, 0, C.byref(out_n), 0, None, err, 512)
                got = (rc, [], err.value.decode())
            else:
                r = M.SimpleCABACMex(args[0], *args[1:], nargout=e["nlhs"])
                got = (0, [] if r is None else [r], "")
        except M.MexError as ex:
            got = (1, [], str(ex))
        assert got[0] == e["rc"], e
        assert got[2] == e["err"], e
        if e["out"] == ["<handle>"]:
            handle = got[1][0]
        elif e["rc"] == 0:
            assert got[1] == e["out"], e
        n_checked += 1
    assert n_checked > 500


def test_cabacwrapper_demo_roundtrip(M, tmp_path):
    """cabacDemo.m:81-186 in miniature through the cabacWrapper mirror (3-context rule, TU)."""
    rng = np.random.default_rng(0)
    x = np.abs(rng.standard_normal(300))
    x[1:] += 0.8 * x[:-1]
    Nq = 4
    delta = np.quantile(x, 0.99) / Nq
    symbols = np.minimum(np.floor(x / delta + 0.5), Nq - 1).astype(int)
    fn = str(tmp_path / "test.bin")
    ctxInit = O.matlab_uint8(0.5 * np.ones(3) * 255) / 255.0
    strings = [M.cabacBinarizer(v, Nq, "DEC2TU") for v in symbols]
    enc = M.cabacWrapper(ctxInit, fn)
    enc.encodeStart()
    for it, g in enumerate(strings):
        for n, b in enumerate(g, start=1):
            ctxID = 0
            if n == 1 and it > 0:
                ctxID = 1 if strings[it - 1][0] == 1 else 2
            enc.encodeBin(b, ctxID)
    enc.encodeFinish()
    # the oracle's demo-profile encoder writes the same file
    cfg = O.make_cfg(O.PROFILE_DEMO, O.BIN_TU, Nq)
    s_ref, l_ref = O.encode_symbols(cfg, symbols.astype(np.uint32), np.array([0, len(symbols)], dtype=np.uint64),
                                    O.ctx_from_p0(ctxInit), 1024)
    assert open(fn, "rb").read() == bytes(s_ref[0, :l_ref[0]])
    dec = M.cabacWrapper(ctxInit, fn)
    dec.decodeStart()
    out = []
    for it in range(len(symbols)):
        g, n, fin, n_s, n_p = [], 1, False, -1, 0
        while not fin:
            ctxID = 0
            if n == 1 and it > 0:
                ctxID = 1 if out[it - 1][0] == 1 else 2
            g.append(dec.decodeBin(ctxID))
            fin, n, n_p, n_s = M.cabacDecodeSymbolFinished(g, n, Nq, "DEC2TU", n_p, n_s)
            n += 1
        out.append(g)
    dec.decodeFinish()
    assert [M.cabacDebinarizer(g, Nq, "DEC2TU") for g in out] == list(symbols)
    enc.close(); dec.close()


def test_handle_api_engine_calls(golden_dir):
    """SimpleCABAC.cpp:48-177 (K1) through the handle API incl. encodeBinsEP / decodeBinsEP / TRM."""
    import isscabac_b200 as I
    L = I.lib()
    with open(os.path.join(golden_dir, "kat.json")) as f:
        kat = json.load(f)
    for name in ("K1", "X_trm", "X_binsep", "K6", "K8"):
        k = kat[name]
        h = C.c_void_p()
        I._lib.check(L.simplecabac_create(C.byref(h), None))
        tri = np.array([[i, c & 1, c >> 1] for i, c in enumerate(k["ctx"])], dtype=np.float64).reshape(-1)
        I._lib.check(L.simplecabac_init_by_state(h, tri.ctypes.data_as(C.POINTER(C.c_double)), len(k["ctx"])))
        I._lib.check(L.simplecabac_encode_start(h))
        for kind, a, b in k["script"]:
            if kind == 0:
                rc = L.simplecabac_encode_bin(h, a, b)
            elif kind == 1:
                rc = L.simplecabac_encode_bin_ep(h, a)
            elif kind == 2:
                rc = L.simplecabac_encode_bins_ep(h, C.c_uint(a), b)
            else:
                rc = L.simplecabac_encode_bin_trm(h, a)
            I._lib.check(rc)
        I._lib.check(L.simplecabac_encode_finish(h))
        p, n = C.POINTER(C.c_uint8)(), C.c_uint64()
        I._lib.check(L.simplecabac_get_bytes(h, C.byref(p), C.byref(n)))
        assert bytes(p[: n.value]).hex() == k["bytes"], name
        for c, want in enumerate(k["ctx_final"]):
            st, mps = C.c_uint(), C.c_uint()
            I._lib.check(L.simplecabac_get_ctx_state(h, 0, c, C.byref(st), C.byref(mps)))
            assert (st.value << 1) | mps.value == want
        # decode from the memory sink
        I._lib.check(L.simplecabac_decode_start(h))
        got = []
        for kind, a, b in k["script"]:
            v = C.c_uint()
            if kind == 0:
                rc = L.simplecabac_decode_bin(h, b, C.byref(v))
            elif kind == 1:
                rc = L.simplecabac_decode_bin_ep(h, C.byref(v))
            elif kind == 2:
                rc = L.simplecabac_decode_bins_ep(h, b, C.byref(v))
            else:
                rc = L.simplecabac_decode_bin_trm(h, C.byref(v))
            I._lib.check(rc)
            got.append(v.value)
        assert got == k["decoded"], name
        I._lib.check(L.simplecabac_decode_finish(h))
        L.simplecabac_destroy(h)


def test_handle_errors():
    import isscabac_b200 as I
    L = I.lib()
    h = C.c_void_p()
    I._lib.check(L.simplecabac_create(C.byref(h), b"/nonexistent_dir/x.bin"))
    assert L.simplecabac_encode_start(h) == -7          # ISSCABAC_ERR_IO
    assert L.simplecabac_encode_bin(h, 1, 0) == -6       # ISSCABAC_ERR_STATE
    assert L.simplecabac_decode_start(h) == -7
    L.simplecabac_destroy(h)
    h2 = C.c_void_p()
    I._lib.check(L.simplecabac_create(C.byref(h2), None))
    I._lib.check(L.simplecabac_set_bytes(h2, (C.c_uint8 * 3)(0x12, 0x34, 0x56), 3))
    I._lib.check(L.simplecabac_decode_start(h2))
    assert L.simplecabac_decode_finish(h2) == -8         # ISSCABAC_ERR_CORRUPT
    L.simplecabac_destroy(h2)
