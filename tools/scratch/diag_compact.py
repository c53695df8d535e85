import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
import isscabac_b200 as I
import ctypes as C
dev = torch.device("cuda")
n = 65520
stride = 256
slab = torch.zeros((n, stride), dtype=torch.uint8, device=dev)
lens = torch.randint(60, 110, (n,), dtype=torch.int32, device=dev)
enc = I.Encoded(slab, lens, torch.zeros(4, dtype=torch.int32, device=dev))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p = I.compact(enc)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("compact total ms", (t1 - t0) * 1e3)
import isscabac_b200.engine as E
L = I.lib()
for rep in range(2):
    t = [time.perf_counter()]
    scratch = torch.empty(int(L.cabac_compact_scratch_bytes(C.c_uint32(n))), dtype=torch.uint8, device=dev); t.append(time.perf_counter())
    byte_off = torch.empty(n + 1, dtype=torch.int64, device=dev); t.append(time.perf_counter())
    E.check(L.cabac_compact(C.c_uint32(n), E.vp(enc.slab), C.c_uint64(stride), E.vp(enc.lengths), None, C.c_uint64(0), E.vp(byte_off), E.vp(scratch), E.vp(enc.overflow), E._stream_ptr())); t.append(time.perf_counter())
    last = byte_off[-1]; t.append(time.perf_counter())
    cap = int(last.item()); t.append(time.perf_counter())
    payload = torch.empty(cap, dtype=torch.uint8, device=dev); t.append(time.perf_counter())
    E.check(L.cabac_compact(C.c_uint32(n), E.vp(enc.slab), C.c_uint64(stride), E.vp(enc.lengths), E.vp(payload), C.c_uint64(cap), E.vp(byte_off), E.vp(scratch), E.vp(enc.overflow), E._stream_ptr())); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    print(["%.3f" % ((b - a) * 1e3) for a, b in zip(t, t[1:])])
