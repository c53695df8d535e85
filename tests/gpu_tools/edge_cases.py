import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import isscabac_b200 as I
import oracle as O
rng = np.random.default_rng(77)
def case(n_streams, n_ops, n_ctx, check_oracle):
    lens = np.full(n_streams, n_ops) if n_streams < 1000 else rng.integers(0, n_ops + 1, size=n_streams)
    off = np.zeros(n_streams + 1, dtype=np.uint64); np.cumsum(lens, out=off[1:])
    n = int(off[-1])
    code = rng.integers(0, n_ctx, size=n).astype(np.uint8)
    code[rng.random(n) < 0.25] = O.OP8_EP
    bins = (rng.random(n) < 0.3).astype(np.uint8)
    ops = ((code << 1) | bins).astype(np.uint8)
    ci = rng.integers(0, 126, size=n_ctx).astype(np.uint8)
    stride = ((n_ops // 4 + 80) + 15) & ~15
    res = {}
    for split in ("0", "1"):
        os.environ["ISSCABAC_ENC_SPLIT"] = split
        t0 = time.time()
        enc = I.encode_ops(ops, off.astype(np.int64), ci, slab_stride=stride)
        pay = I.compact(enc)
        torch.cuda.synchronize()
        enc.check_overflow()
        res[split] = (enc.lengths.cpu().numpy().copy(), pay.payload.cpu().numpy().copy())
    os.environ.pop("ISSCABAC_ENC_SPLIT")
    assert (res["0"][0] == res["1"][0]).all() and (res["0"][1] == res["1"][1]).all(), "formulations differ"
    dbins, ok = I.decode_ops(pay, ops, off.astype(np.int64), ci)
    assert bool(ok.all().item()) and (dbins.cpu().numpy() == bins).all(), "round trip"
    if check_oracle:
        s_ref, l_ref = O.encode_ops(ops, off, ci, out_stride=stride, n_threads=8)
        p_ref, _ = O.compact(s_ref, l_ref)
        assert (res["0"][0].astype(np.uint32) == l_ref).all() and (res["0"][1][:len(p_ref)] == p_ref).all(), "oracle"
    print("ok", n_streams, n_ops, n_ctx, n, "payload", len(res["0"][1]))
case(1, 1 << 26, 23, True)          # one 64 M-op stream
case(3, (1 << 24) + 5, 124, True)   # long streams, many contexts
case(1 << 20, 40, 4, True)          # a million tiny ragged streams
case(300000, 700, 23, False)        # many mid-size streams (round trip + formulations agree)
