"""CPU: the oracle (C restatement) against every golden vector produced by the
unmodified reference engine (oracle/gen_golden.py -> tests/golden)."""
import json
import os

import numpy as np
import pytest

import oracle as O


@pytest.fixture(scope="module")
def kat(golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as f:
        return json.load(f)


SURVEY_KATS = {  # SURVEY.md section 4.1, measured from the unmodified reference
    "K0": "fe80", "K1": "abb12f09", "K2": "ab4424", "K3": "ddcf11310f80", "K4": "ddcf11310f80",
    "K5": "87002de0", "K6": "feffffffffffffffff80", "K7": "5b000503ea6508fb70", "K8": "2ec949bc9948410658",
}


def test_golden_file_matches_survey(kat):
    for k, v in SURVEY_KATS.items():
        assert kat[k]["bytes"] == v


def test_oracle_encode_kats(kat):
    for name, k in kat.items():
        data, fin = O.encode_script([tuple(e) for e in k["script"]], k["ctx"])
        assert data.hex() == k["bytes"], name
        assert [int(c) for c in fin] == k["ctx_final"], name


def test_oracle_decode_kats(kat):
    for name, k in kat.items():
        if k["decoded"] is None:
            continue
        ds = [(e[0], 0, e[2]) for e in k["script"]]
        got = O.decode_script(ds, k["ctx"], bytes.fromhex(k["bytes"]))
        assert got == k["decoded"], name
        # and the decoded values are the encoded ones
        want = [e[1] for e in k["script"]]
        assert got == want, name


@pytest.mark.parametrize("fname", ["random_ops.npz", "random_ops16.npz"])
def test_oracle_random_ops(golden_dir, fname):
    z = np.load(os.path.join(golden_dir, fname))
    ops, off = z["ops"], z["op_off"]
    for tag, ci in (("shared", z["ctx_shared"]), ("per", z["ctx_per"])):
        slab, lens = O.encode_ops(ops, off, ci, n_threads=4)
        assert (lens == z["lens_" + tag]).all()
        payload, boff = O.compact(slab, lens)
        assert (payload == z["payload_" + tag]).all()
        bins, ok = O.decode_ops(payload, boff, ops, off, ci, n_threads=4)
        assert ok.all()
        assert (bins == (ops & 1)).all()


def test_oracle_prob_to_state(golden_dir):
    with open(os.path.join(golden_dir, "prob_to_state.json")) as f:
        g = json.load(f)
    got = O.ctx_from_p0(g["p0"])
    assert [int(x) for x in got] == g["ctx"]
    # gotcha 12/13 of SURVEY appendix A
    assert int(O.ctx_from_p0([0.5])[0]) == 0            # mps=0,state=0
    assert int(O.matlab_uint8([127.5])[0]) == 128
    assert int(O.matlab_uint8([300.0])[0]) == 255 and int(O.matlab_uint8([-3.0])[0]) == 0


def test_oracle_symbol_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "symbols_refengine.npz"))
    names = sorted({k[:-4] for k in z.files if k.endswith("_sym")})
    assert len(names) >= 7
    for nm in names:
        prof, meth, Nq, Nlbp, types, rows = [int(x) for x in z[nm + "_cfg"]]
        cfg = O.make_cfg(prof, meth, Nq, Nlbp, types, rows)
        sym = z[nm + "_sym"]
        ops = O.symbols_to_ops(cfg, sym)
        assert (ops == z[nm + "_ops"]).all(), nm
        off = np.array([0, len(sym)], dtype=np.uint64)
        want = z[nm + "_bytes"]
        slab, lens = O.encode_symbols(cfg, sym, off, z[nm + "_ctx"], out_stride=len(want) + 16)
        assert lens[0] == len(want) and (slab[0, :lens[0]] == want).all(), nm
        dec, ok = O.decode_symbols(cfg, want, np.array([0, len(want)], dtype=np.uint64), off, z[nm + "_ctx"])
        assert ok.all() and (dec == sym).all(), nm
