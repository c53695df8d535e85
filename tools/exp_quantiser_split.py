import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
from isscabac_b200 import quantizer as QZ
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(5)
sizes = np.tile(np.array([400 * 20, 109 * 20], dtype=np.int64), 1638)
off = np.zeros(len(sizes) + 1, dtype=np.int64); np.cumsum(sizes, out=off[1:])
x = torch.log(torch.empty(int(off[-1]), dtype=torch.float64, device=dev).exponential_(1.0, generator=g) ** (1 / 0.6) + 1e-5)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, kw in (("lloyd", dict(GMM=1)), ("lloyd 1 iter", dict(GMM=1, nIter=1)), ("uniform", dict(GMM=0))):
    cfg = QZ.make_quant_cfg(N=8, deadzoneQuant=0.7, **kw)
    print(name, "%.3f ms" % timed(lambda: QZ.quantize_matrices((x, off), cfg)))
