import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import isscabac_b200 as I
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(4)
cfg = I.make_cfg(I.PROFILE_FLAT_EPSUF, I.BIN_EG2, 256, 3, 0, rows=0)
ctx = torch.full((4,), 1, dtype=torch.uint8, device=dev)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r
for n_streams, L in ((1, 34000), (32, 34000), (2048, 34000), (65536, 2000), (1 << 20, 422)):
    sym = torch.floor(-6.0 * torch.log(torch.rand(n_streams * L, generator=g, device=dev))).clamp_(0, 255).to(torch.uint8)
    off = torch.arange(n_streams + 1, dtype=torch.int64, device=dev) * L
    enc = I.encode_symbols(cfg, sym, off, ctx, slab_stride=((L * 5) // 4 + 64 + 15) & ~15)
    pay = I.compact(enc)
    ms_e, _ = t(lambda: I.encode_symbols(cfg, sym, off, ctx, slab_stride=enc.slab.shape[1]))
    ms_d, (dec, ok) = t(lambda: I.decode_symbols(cfg, pay, off, ctx, sym_dtype=torch.uint8))
    assert bool(ok.all().item()) and bool((dec == sym).all().item())
    bins = n_streams * L * 4.3
    print(f"streams {n_streams:8d} x {L:6d} symbols: encode {ms_e:8.3f} ms  decode {ms_d:8.3f} ms  -> per stream-bin {ms_d*1e-3*1.965e9/(L*4.3):7.1f} cycles (decode), {bins/ms_d/1e6:8.1f} Gbins/s")
