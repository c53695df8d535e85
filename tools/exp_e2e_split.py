import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import isscabac_b200 as I
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda", 0)
S, B = 65536, 65536
ops = bench.gen_ops_device(torch, 2, S, B, dev)
total = S * B
h_ops = torch.empty(total, dtype=torch.uint8, pin_memory=True); h_ops.copy_(ops)
h_off = np.arange(S + 1, dtype=np.uint64) * np.uint64(B)
h_ctx = np.full(23, 1, dtype=np.uint8)
h_pay = torch.empty(total // 7, dtype=torch.uint8, pin_memory=True)
h_boff = np.empty(S + 1, dtype=np.uint64)
h_bins = torch.empty(total, dtype=torch.uint8, pin_memory=True)
np_ops, np_pay, np_bins = h_ops.numpy(), h_pay.numpy(), h_bins.numpy()
# raw link
d = torch.empty(total, dtype=torch.uint8, device=dev)
for name, fn in (("h2d", lambda: d.copy_(h_ops, non_blocking=True)), ("d2h", lambda: h_bins.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(name, "%.1f GB/s" % (total / dt / 1e9))
s1 = torch.cuda.Stream(); s2 = torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h_ops, non_blocking=True)
d2 = torch.empty(total, dtype=torch.uint8, device=dev)
with torch.cuda.stream(s2): h_bins.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("duplex: %.1f GB/s each way" % (total / dt / 1e9))
del d, d2
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p, bo = I.encode_ops_host(np_ops, h_off, h_ctx, payload_out=np_pay, byte_off_out=h_boff)
    t1 = time.perf_counter()
    I.decode_ops_host(p, bo, np_ops, h_off, h_ctx, bins_out=np_bins)
    t2 = time.perf_counter()
    print("encode_host %.1f ms  decode_host %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
assert (np_bins[:1 << 20] == (np_ops[:1 << 20] & 1)).all()
