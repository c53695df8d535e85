import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
import isscabac_b200 as I
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(1)
T = I.CM_COND0 | I.CM_COND1 | I.CM_CONDS0 | I.CM_CONDS1
n_streams, rows = 65520, 400
u = torch.rand(n_streams * rows, generator=g, device=dev)
sym = torch.where(u < 0.7, torch.zeros_like(u), 1 + torch.floor(torch.log(torch.rand(u.shape, generator=g, device=dev)) / np.log(0.6))).clamp_(0, 7).to(torch.uint8)
off = torch.arange(n_streams + 1, dtype=torch.int64, device=dev) * rows
cfg = I.make_cfg(I.PROFILE_ISS, I.BIN_EG0, 8, 3, T, rows=rows)
ctx = torch.full((23,), 1, dtype=torch.uint8, device=dev)
ops, op_off = I.binarize_symbols(cfg, sym, off)
longest = int((op_off[1:] - op_off[:-1]).max().item())
stride = (longest // 4 + 64 + 15) & ~15
print("stride", stride, "longest", longest)
enc = I.Encoded(torch.empty((n_streams, stride), dtype=torch.uint8, device=dev),
                torch.empty(n_streams, dtype=torch.int32, device=dev), torch.zeros(4, dtype=torch.int32, device=dev))
I.encode_ops(ops, op_off, ctx, out=enc)
torch.cuda.synchronize()
print("lens", enc.lengths[:8].tolist(), int(enc.lengths.max()), int(enc.lengths.min()))
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p = I.compact(enc)
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("compact ms launch %.3f total %.3f" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
