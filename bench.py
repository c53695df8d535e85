#!/usr/bin/env python
"""bench.py -- many-stream CABAC encode+decode throughput (BASELINE.json metric) on N B200s.

Workload (config C3 of SURVEY.md 8(d), BASELINE.json configs[2] -- the configuration the metric
"Gbins/s over 64K independent CABAC streams" is quoted on; it fits one GPU):
  65,536 independent streams x 65,536 bins per GPU, raw-ops profile: each bin is a bypass bin
  with probability 0.25 (uniform value), else context-coded with ctx ~ U{0..22} and
  P(bin=1) = 0.20 + 0.10*(ctx mod 5); all 23 contexts start at (mps=1, state=0).
One STEP = encode every stream (k_encode_ops) + device-wide length scan + compaction into one
contiguous bitstream + decode every stream from that bitstream (k_decode_ops).
value = (bins encoded + bins decoded) per second, whole job (all ranks), inputs resident in HBM.
e2e   = the same through the host-buffer C ABI (cabac_encode_ops_host / cabac_decode_ops_host)
        with pinned host arrays, H2D/D2H inside the timed region.
Multi-GPU: streams shard across ranks with no data-path collective (weak scaling: every rank
codes its own 65,536 streams); the only exchange is the all-gather of per-stream lengths that
lets every rank compute global offsets (torch.distributed / NCCL), inside the timed step.

The same run also measures the other BASELINE.json configurations and the strong-scaling form of C3 and reports them in a
`configs` block of the same JSON line (see run_configs): c3_strong (65,536 streams TOTAL sharded over the N ranks), c2 (ISS
column streams, fused and two-pass encoders), c4 (2^20 segments x 1,024 symbols sharded over N: encode + length exchange +
device scan + compaction + payload assembly over NCCL / NVLink inside the timed step) and c5 (decode only, 2^20 ragged
streams, work-balanced sharding).  Each entry carries ms, Gbins/s, its HBM and int32-issue fractions and -- at N = 1 -- a CPU
baseline on a stated sample of its streams whose GPU bytes are verified identical.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--streams S] [--bins B] [--configs c3_strong,c2,c4,c5 | --no-configs]
  python bench.py --impl reference ...   # the reference CPU engine (oracle/_ref) on host cores, one stream per core
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CTX = 23
P_EP = 0.25


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=65536, help="streams per GPU")
    ap.add_argument("--bins", type=int, default=65536, help="bins per stream")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--configs", default="c3_strong,c2,c4,c5", help="BASELINE configurations measured beside the headline")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--config-steps", type=int, default=8)
    ap.add_argument("--config-scale", type=float, default=1.0, help="shrinks the stream counts of the configs block (tests)")
    return ap.parse_args()


def workload_name(a):
    return f"C3: {a.streams} streams x {a.bins} bins/stream per GPU, 23 ctx, 25% bypass, P(1)=0.2+0.1*(ctx%5)"


# ---------------------------------------------------------------------------------------
# synthetic ops
# ---------------------------------------------------------------------------------------
def gen_ops_numpy(seed, n_streams, n_bins):
    """CPU generator (reference arm / cpu baseline sample): same distribution as the device one."""
    rng = np.random.default_rng(seed)
    n = n_streams * n_bins
    ctx = rng.integers(0, N_CTX, size=n, dtype=np.uint8)
    p1 = (0.20 + 0.10 * (np.arange(N_CTX) % 5)).astype(np.float32)
    u = rng.random(n, dtype=np.float32)
    bins = (u < p1[ctx]).astype(np.uint8)
    ep = rng.random(n, dtype=np.float32) < P_EP
    code = ctx
    code[ep] = 126
    bins[ep] = (rng.integers(0, 2, size=int(ep.sum()), dtype=np.uint8))
    return ((code << 1) | bins).astype(np.uint8)


def gen_ops_device(torch, seed, n_streams, n_bins, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    total = n_streams * n_bins
    ops = torch.empty(total, dtype=torch.uint8, device=device)
    p1 = (0.20 + 0.10 * (torch.arange(N_CTX, device=device) % 5)).float()
    chunk = 1 << 26
    for a in range(0, total, chunk):
        b = min(total, a + chunk)
        m = b - a
        ctx = torch.randint(0, N_CTX, (m,), generator=g, device=device)
        u = torch.rand(m, generator=g, device=device)
        bins = (u < p1[ctx]).to(torch.uint8)
        ep = torch.rand(m, generator=g, device=device) < P_EP
        code = torch.where(ep, torch.full_like(ctx, 126), ctx).to(torch.uint8)
        bins = torch.where(ep, (torch.rand(m, generator=g, device=device) < 0.5).to(torch.uint8), bins)
        ops[a:b] = (code << 1) | bins
        del ctx, u, bins, ep, code
    return ops


# ---------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms in the background; every row is
    stamped on arrival so that only the rows inside the timed region are summarised."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first(self, timeout=5.0):
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    def summary(self, t0, t1):
        sm, mx, pw, reasons, n_all = [], 0, [], set(), 0
        for ts, r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            n_all += 1
            if not (t0 - 0.02 <= ts <= t1 + 0.05):
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons),
                "samples": len(sm), "samples_total": n_all}


# ---------------------------------------------------------------------------------------
# reference / cpu baseline (oracle/_ref = the unmodified reference engine; else the oracle port)
# ---------------------------------------------------------------------------------------
def _cpu_worker(job):
    """One process of the CPU baseline: encode + decode a contiguous chunk of streams on ONE thread with the reference
    engine.  -> (seconds_enc, seconds_dec, kind, slab, lens).  Runs in a spawned process (no torch, no CUDA)."""
    ops, n_streams, n_bins = job
    import oracle as O
    impl, kind = ("ref", "reference") if O.ref() is not None else ("oracle", "port")
    off = (np.arange(n_streams + 1, dtype=np.uint64) * np.uint64(n_bins))
    ci = np.full(N_CTX, 1, dtype=np.uint8)
    stride = n_bins // 4 + 64
    t0 = time.perf_counter()
    slab, lens = O.encode_ops(ops, off, ci, out_stride=stride, n_threads=1, impl=impl)
    t1 = time.perf_counter()
    payload, boff = O.compact(slab, lens)
    t2 = time.perf_counter()
    bins, ok = O.decode_ops(payload, boff, ops, off, ci, n_threads=1, impl=impl)
    t3 = time.perf_counter()
    assert ok.all() and (bins == (ops & 1)).all(), "reference round trip failed"
    w = int(lens.max()) if n_streams else 0
    return t1 - t0, t3 - t2, kind, slab[:, :w].copy(), lens


_POOL = None


def cpu_pool(procs):
    """Spawned once, reused by every CPU-baseline call of the run (spawn, not fork: the parent may hold a CUDA context)."""
    global _POOL
    if _POOL is None:
        import multiprocessing as mp
        _POOL = mp.get_context("spawn").Pool(procs)
        _POOL.map(_cpu_warm, range(procs))
    return _POOL


def _cpu_warm(_):
    import oracle as O
    O.lib()
    O.ref()
    return 0


def cpu_roundtrip(ops, n_streams, n_bins, procs):
    """encode + decode `ops` with the reference engine, one stream per core: `procs` single-threaded PROCESSES, each with
    a contiguous share of the streams (threads of one process contend inside the engine's file streams: decode ran at 15
    instead of 40+ Mbins/s per core in round 1).  Times are the slowest process's own compute times.
    -> (seconds_enc, seconds_dec, kind, slab, lens)"""
    procs = max(1, min(procs, n_streams))
    cuts = [n_streams * k // procs for k in range(procs + 1)]
    jobs = [(ops[cuts[k] * n_bins: cuts[k + 1] * n_bins], cuts[k + 1] - cuts[k], n_bins) for k in range(procs)]
    res = cpu_pool(procs).map(_cpu_worker, jobs)
    te, td = max(r[0] for r in res), max(r[1] for r in res)
    w = max(r[3].shape[1] for r in res)
    slab = np.zeros((n_streams, w), dtype=np.uint8)
    for k, r in enumerate(res):
        slab[cuts[k]:cuts[k + 1], :r[3].shape[1]] = r[3]
    lens = np.concatenate([r[4] for r in res])
    return te, td, res[0][2], slab, lens


HOT_KERNEL_SOURCES = ("kernels.cu", "cabac_wide.cuh", "wide_common.cuh", "cabac_lane.cuh")


def hot_kernel_hash():
    """Stamp of the sources the hot kernels are compiled from: profiles/traffic.json (an ncu capture) is only cited while
    it matches; after a kernel change it reads as stale instead of silently describing other code."""
    import hashlib
    h = hashlib.sha256()
    for name in HOT_KERNEL_SOURCES:
        with open(os.path.join(ROOT, "isscabac_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def int_peak_lanes():
    """Measured int32 issue peaks of one SM (tools/int_peak.cu -> profiles/int_peak.json): lanes per clock on the ALU pipe
    alone and on ALU + FMA pipes together.  -> (alu, alu_plus_fma, source); falls back to 64 / 128 and says so."""
    try:
        with open(os.path.join(ROOT, "profiles", "int_peak.json")) as f:
            j = json.load(f)
        return float(j["alu_pipe_lanes_per_clk_per_sm"]), float(j["alu_plus_fma_lanes_per_clk_per_sm"]), \
            "measured: profiles/int_peak.json (tools/int_peak.cu)"
    except Exception:
        return 64.0, 128.0, "FALLBACK (profiles/int_peak.json missing): 64 ALU lanes, 128 issue lanes per clock and SM"


def bind_to_gpu_numa(local_index):
    """Run this process (and so its first-touch / pinned allocations) on the CPUs of the NUMA node the GPU hangs off
    (sysfs: the PCI device's numa_node -> that node's cpulist).  -> a description for the result line."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        if node < 0 or len(nodes) < 2:
            return f"single NUMA node ({len(nodes)} node(s), GPU reports node {node}): nothing to bind"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return f"GPU on NUMA node {node}, none of its CPUs is available to this process"
        os.sched_setaffinity(0, allowed)
        return f"bound to NUMA node {node} of GPU {bdf} ({len(allowed)} CPUs) before the pinned buffers were allocated"
    except Exception as e:   # best effort: a missing sysfs entry must not fail the bench
        return f"not bound ({type(e).__name__})"


def pcie_bandwidth(torch, dev, nbytes=1 << 30):
    """Host <-> device copy bandwidth of this rank, pinned memory, measured in the run (the e2e numbers are read against it)."""
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    res = {}
    for name, (dst, src) in (("h2d", (d, h)), ("d2h", (h, d))):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        res[name] = 3 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    del h, d
    return res


def pcie_bandwidth_concurrent(torch, dist, dev, world, nbytes=1 << 30):
    """All ranks copy host -> device AT THE SAME TIME (they share the host's PCIe root complexes and DRAM): aggregate and
    per-rank bandwidth, the roof the e2e numbers at N > 1 are read against."""
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    del h, d
    agg = world * 4 * nbytes / dt / 1e9
    return {"h2d_aggregate": agg, "h2d_per_rank": agg / world, "ranks": world}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    # bounded sample of the same workload: a few hundred streams of the full length
    n_streams = int(min(a.streams, max(64, min(threads * 16, 4096))))
    ops = gen_ops_numpy(a.seed, n_streams, a.bins)
    times, tes, tds = [], [], []
    kind = "port"
    for i in range(a.warmup + a.steps):
        te, td, kind, _, _ = cpu_roundtrip(ops, n_streams, a.bins, threads)
        if i >= a.warmup:
            times.append(te + td)
            tes.append(te)
            tds.append(td)
    t = float(np.mean(times))
    bins = 2.0 * n_streams * a.bins
    value = bins / t / 1e9
    sample = (f"{n_streams} of {a.streams} streams x {a.bins} bins, encode+decode, one stream per core: {threads} single-threaded "
              f"processes, files on tmpfs")
    print(json.dumps({
        "impl": "reference", "metric": "CABAC encode+decode throughput over independent streams",
        "value": value, "unit": "Gbins/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Gbins/s", "cores": threads, "kind": kind, "sample": sample,
                         "encode_mbins_per_core": n_streams * a.bins / float(np.mean(tes)) / 1e6 / threads,
                         "decode_mbins_per_core": n_streams * a.bins / float(np.mean(tds)) / 1e6 / threads},
        "e2e": {"value": value, "unit": "Gbins/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    if _POOL is not None:
        _POOL.close()
        _POOL.join()


# ---------------------------------------------------------------------------------------
# the other BASELINE configurations + strong scaling, measured in the same run ("configs" block)
# ---------------------------------------------------------------------------------------
class _Env:
    """What every config needs: torch / dist / the package, this rank's place in the job, the peaks."""

    def __init__(self, a, torch, dist, I, dev, rank, world, hbm_peak, sm_mhz):
        self.a, self.torch, self.dist, self.I, self.dev, self.rank, self.world = a, torch, dist, I, dev, rank, world
        self.hbm_peak = hbm_peak
        lanes_alu, lanes_both, self.int_src = int_peak_lanes()
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        self.int_peak = sms * lanes_both * sm_mhz * 1e6          # per GPU, ops/s
        self.steps, self.warmup = max(1, a.config_steps), 3
        self.threads = host_threads()
        self._mg = None

    def mg(self):
        if self._mg is None:
            from isscabac_b200 import multi_gpu as MG
            self._mg = MG.default_handle()
        return self._mg

    def timed(self, fn, steps=None):
        """W warm-up steps, then `steps` steps between barrier + synchronize on both sides, CUDA events on the current
        stream, MAX over ranks.  -> ms per step"""
        torch, dist = self.torch, self.dist
        steps = steps or self.steps
        for _ in range(self.warmup):
            fn()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def total(self, x):
        """sum of a per-rank count over the ranks"""
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t)
        return float(t.item())

    def fracs(self, ms, alg_bytes_rank, int_ops_rank):
        """fractions of the per-GPU HBM and int32-issue peaks reached by the slowest rank's launch(es)"""
        sec = ms * 1e-3
        return {"hbm_frac": alg_bytes_rank / sec / 1e9 / self.hbm_peak, "int_frac": int_ops_rank / sec / self.int_peak,
                "algorithmic_bytes_per_gpu": int(alg_bytes_rank)}


def _int_ops(n_ctx_bins, n_ep_bins, payload_bytes, enc):
    """algorithmic int32 ops: 16 per context bin, 6 per bypass bin, 12 (enc) | 4 (dec) per payload byte"""
    return n_ctx_bins * 16 + n_ep_bins * 6 + payload_bytes * (12 if enc else 4)


def _sample_ids(n, want, run=8):
    """runs of `run` consecutive streams spread evenly over [0, n) (first and last included)"""
    if n <= want:
        return np.arange(n, dtype=np.int64)
    starts = np.unique(np.round(np.linspace(0, n - run, max(1, want // run))).astype(np.int64))
    return np.unique(np.concatenate([np.arange(st, st + run) for st in starts]))


def _cpu_symbols(env, cfg_args, sym, sym_off_np, ids, ctx_np, enc, label):
    """CPU baseline of a symbol-level config on a sample of its streams (oracle port: binarizer + context selection +
    engine restatement, env.threads threads) with byte parity of the GPU encoder's output asserted."""
    import oracle as O
    torch = env.torch
    ocfg = O.make_cfg(*cfg_args)
    lens = (sym_off_np[ids + 1] - sym_off_np[ids]).astype(np.int64)
    off_s = np.zeros(len(ids) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off_s[1:])
    gather = np.concatenate([np.arange(sym_off_np[i], sym_off_np[i + 1]) for i in ids]) if len(ids) else np.zeros(0, np.int64)
    sym_s = sym[torch.as_tensor(gather, device=sym.device)].cpu().numpy().astype(np.uint32)
    stride = int(enc.slab.shape[1])
    t0 = time.perf_counter()
    slab_ref, lens_ref = O.encode_symbols(ocfg, sym_s, off_s, ctx_np, stride, n_threads=env.threads)
    t1 = time.perf_counter()
    payload, boff = O.compact(slab_ref, lens_ref)
    t2 = time.perf_counter()
    dec, ok = O.decode_symbols(ocfg, payload, boff, off_s, ctx_np, n_threads=env.threads)
    t3 = time.perf_counter()
    assert ok.all() and (dec == sym_s).all(), label + ": oracle round trip failed"
    idt = torch.as_tensor(ids, device=enc.slab.device)
    lens_gpu = enc.lengths[idt].cpu().numpy().astype(np.uint32)
    assert (lens_gpu == lens_ref).all(), label + ": GPU stream lengths differ from the oracle"
    w = int(lens_ref.max()) if len(ids) else 0
    live = np.arange(w)[None, :] < lens_ref[:, None]
    assert (enc.slab[idt][:, :w].cpu().numpy()[live] == slab_ref[:, :w][live]).all(), label + ": GPU bytes differ from the oracle"
    n_bins = int(sum(len(O.symbols_to_ops(ocfg, sym_s[int(off_s[k]):int(off_s[k + 1])])) for k in range(min(len(ids), 64))))
    bins_per_sym = n_bins / max(1, int(off_s[min(len(ids), 64)]))
    bins = bins_per_sym * len(sym_s)
    return {"unit": "Gbins/s", "kind": "port", "cores": env.threads,
            "encode_gbins": bins / (t1 - t0) / 1e9, "decode_gbins": bins / (t3 - t2) / 1e9,
            "value": 2 * bins / ((t1 - t0) + (t3 - t2)) / 1e9,
            "sample": f"{len(ids)} streams ({len(sym_s)} symbols, runs of 8 spread over the stream range), oracle binarizer + engine on "
                      f"{env.threads} threads; GPU encoder output of these streams verified byte-identical"}


def _count_bins(env, cfg, sym, sym_off):
    """(bins, bypass bins) of this rank's symbols, from the binarizer's op array (untimed bookkeeping for the rooflines)"""
    ops, op_off = env.I.binarize_symbols(cfg, sym, sym_off)
    n = int(ops.numel())
    n_ep = int(((ops >> 1) == 126).sum().item())
    del ops
    return n, n_ep, op_off


def cfg_c3_strong(env, headline):
    """BASELINE configs[2] as SURVEY.md 8(d) states it: 65,536 streams x 65,536 bins TOTAL (2^32 bins), the streams
    sharded over the N ranks; step = encode + length exchange + device scan + compaction + decode."""
    a, torch, I = env.a, env.torch, env.I
    S_total, B = int(a.streams * a.config_scale), a.bins
    if env.world == 1:
        d = dict(headline)
        d["note"] = "N = 1: the strong-scaling job IS the headline job (65,536 streams on one GPU); values copied from it"
        return d
    from isscabac_b200 import multi_gpu as MG
    lo, hi = MG.shard_range(S_total, env.rank, env.world)
    S = hi - lo
    first = np.array([MG.shard_range(S_total, r, env.world)[0] for r in range(env.world)] + [S_total], dtype=np.uint32)
    ops = gen_ops_device(torch, a.seed + 31 * env.rank, S, B, env.dev)
    op_off = torch.arange(S + 1, dtype=torch.int64, device=env.dev) * B
    ctx = torch.full((N_CTX,), 1, dtype=torch.uint8, device=env.dev)
    stride = (B // 4 + 64 + 15) & ~15
    enc = I.Encoded(torch.empty((S, stride), dtype=torch.uint8, device=env.dev), torch.empty(S, dtype=torch.int32, device=env.dev),
                    torch.zeros(4, dtype=torch.int32, device=env.dev))
    L = I.lib()
    scratch = torch.empty(int(L.cabac_compact_scratch_bytes(S_total)), dtype=torch.uint8, device=env.dev)
    byte_off = torch.empty(S + 1, dtype=torch.int64, device=env.dev)
    payload = torch.empty(S * (B // 6 + 64), dtype=torch.uint8, device=env.dev)
    bins = torch.empty(S * B, dtype=torch.uint8, device=env.dev)
    ok = torch.empty(S, dtype=torch.uint8, device=env.dev)
    all_len = torch.empty(S_total, dtype=torch.int32, device=env.dev)
    g_off = torch.empty(S_total + 1, dtype=torch.int64, device=env.dev)
    g_scr = torch.empty(int(L.cabac_compact_scratch_bytes(S_total)), dtype=torch.uint8, device=env.dev)
    mg = env.mg()
    parts = {}

    def step():
        I.encode_ops(ops, op_off, ctx, out=enc)
        mg.gather_table(first, enc.lengths, all_lengths=all_len, byte_off=g_off, scratch=g_scr)
        pay = I.compact(enc, payload=payload, byte_off=byte_off, scratch=scratch)
        I.decode_ops(pay, ops, op_off, ctx, bins=bins, finish_ok=ok)

    ms = env.timed(step)
    parts["encode_ms"] = env.timed(lambda: I.encode_ops(ops, op_off, ctx, out=enc))
    pay = I.compact(enc, payload=payload, byte_off=byte_off, scratch=scratch)
    parts["decode_ms"] = env.timed(lambda: I.decode_ops(pay, ops, op_off, ctx, bins=bins, finish_ok=ok))
    parts["length_exchange_scan_ms"] = env.timed(lambda: mg.gather_table(first, enc.lengths, all_lengths=all_len, byte_off=g_off, scratch=g_scr))
    assert int(enc.overflow[0].item()) == 0 and bool(ok.all().item()) and bool(((ops & 1) == bins).all().item())
    pb = int(byte_off[-1].item())
    assert int(g_off[int(first[env.rank + 1])].item()) - int(g_off[int(first[env.rank])].item()) == pb, "global table disagrees with the local one"
    nb = S * B
    out = {"workload": f"{S_total} streams x {B} bins TOTAL, sharded over {env.world} GPUs ({S} streams on this rank)",
           "scaling": "strong", "ms": ms, "gbins": 2.0 * S_total * B / (ms * 1e-3) / 1e9, **parts,
           "step": "encode + NCCL all-gather-v of lengths + device scan (cabac_multi_gpu_gather_table) + compaction + decode",
           "limit": "per-stream serial chain: a launch lasts as long as ONE stream takes whatever the stream count (encode_ms + decode_ms "
                    "barely move with N); see DESIGN.md"}
    out.update(env.fracs(ms, (nb + pb + 4 * S) + (2 * nb + pb + S), _int_ops(nb * (1 - P_EP), nb * P_EP, pb, True) + _int_ops(nb * (1 - P_EP), nb * P_EP, pb, False)))
    return out


def _iss_symbols(torch, g, n, dev):
    u = torch.rand(n, generator=g, device=dev)
    return torch.where(u < 0.7, torch.zeros_like(u), 1 + torch.floor(torch.log(torch.rand(u.shape, generator=g, device=dev)) / np.log(0.6))).clamp_(0, 7).to(torch.uint8)


def cfg_c2(env):
    """BASELINE configs[1]: ISS NTF parameter coding, one stream per column: 65,520 column streams x 400 symbols (1,638 tracks
    x 2 factor matrices x 20 columns), ISS context rule, EG0, Nq = 8 (ISS/+coder/cabacEncode.m:45-70)."""
    a, torch, I = env.a, env.torch, env.I
    from isscabac_b200 import multi_gpu as MG
    n_total, rows = max(64, int(65520 * a.config_scale)), 400
    lo, hi = MG.shard_range(n_total, env.rank, env.world)
    n = hi - lo
    g = torch.Generator(device=env.dev)
    g.manual_seed(1 + 101 * env.rank)
    sym = _iss_symbols(torch, g, n * rows, env.dev)
    off = torch.arange(n + 1, dtype=torch.int64, device=env.dev) * rows
    T = I.CM_COND0 | I.CM_COND1 | I.CM_CONDS0 | I.CM_CONDS1
    cfg_args = (I.PROFILE_ISS, I.BIN_EG0, 8, 3, T, rows)
    cfg = I.make_cfg(*cfg_args)
    ctx = torch.full((23,), 1, dtype=torch.uint8, device=env.dev)
    n_bins, n_ep, op_off = _count_bins(env, cfg, sym, off)
    stride = (int((op_off[1:] - op_off[:-1]).max().item()) // 4 + 64 + 15) & ~15
    enc = I.encode_symbols(cfg, sym, off, ctx, slab_stride=stride)
    pay = I.compact(enc)
    enc.check_overflow()
    pb = int(pay.byte_off[-1].item())
    scratch = torch.empty(int(I.lib().cabac_compact_scratch_bytes(n)), dtype=torch.uint8, device=env.dev)
    ms_enc = env.timed(lambda: I.encode_symbols(cfg, sym, off, ctx, slab_stride=stride))
    ms_cmp = env.timed(lambda: I.compact(enc, payload=pay.payload, byte_off=pay.byte_off, scratch=scratch))
    ms_dec = env.timed(lambda: I.decode_symbols(cfg, pay, off, ctx, sym_dtype=torch.uint8))
    dec, ok = I.decode_symbols(cfg, pay, off, ctx, sym_dtype=torch.uint8)
    assert bool(ok.all().item()) and bool((dec == sym).all().item()), "c2 round trip failed"
    ms_bin = env.timed(lambda: I.binarize_symbols(cfg, sym, off))
    ops, op_off2 = I.binarize_symbols(cfg, sym, off)
    enc2 = I.Encoded(torch.empty((n, stride), dtype=torch.uint8, device=env.dev), torch.empty(n, dtype=torch.int32, device=env.dev),
                     torch.zeros(4, dtype=torch.int32, device=env.dev))
    ms_eops = env.timed(lambda: I.encode_ops(ops, op_off2, ctx, out=enc2))
    assert bool((enc2.lengths == enc.lengths).all().item()), "fused and two-pass encoders disagree"
    tot_bins = env.total(n_bins)
    ms = ms_enc + ms_cmp + ms_dec
    out = {"workload": f"{n_total} ISS column streams x {rows} symbols (P(0) = 0.7, Nq = 8, EG0, 23 contexts), {tot_bins / 1e6:.1f} M bins",
           "scaling": "strong", "ms": ms, "gbins": 2.0 * tot_bins / (ms * 1e-3) / 1e9,
           "encode_fused_ms": ms_enc, "compact_ms": ms_cmp, "decode_ms": ms_dec,
           "encode_two_pass_ms": {"binarize": ms_bin, "encode_ops": ms_eops},
           "encode_gbins": tot_bins / (ms_enc * 1e-3) / 1e9, "decode_gbins": tot_bins / (ms_dec * 1e-3) / 1e9,
           "step": "fused binarize + context selection + encode, scan + compaction, fused decode + debinarize",
           "limit": "launch-bound: a stream is a serial chain of ~740 bins; the job is too small to fill the GPU for long"}
    out.update(env.fracs(ms, (n * rows + pb + 4 * n) + 2 * pb + (pb + n * rows), _int_ops(n_bins - n_ep, n_ep, pb, True) + _int_ops(n_bins - n_ep, n_ep, pb, False)))
    if env.rank == 0 and env.world == 1 and not a.no_cpu:
        out["cpu_baseline"] = _cpu_symbols(env, cfg_args, sym, off.cpu().numpy(), _sample_ids(n, 2048), np.full(23, 1, np.uint8), enc, "c2")
    return out


def cfg_c4(env):
    """BASELINE configs[3]: one 1G-symbol array cut into 2^20 fixed segments of 1,024 symbols with reset contexts; per rank:
    fused encode of its segments, then the length exchange + device-wide scan and the compaction into ONE bitstream that
    every rank holds -- the payload exchange inside the timed step, once fused into the compaction kernel (peer stores
    over NVLink), once as grouped NCCL broadcasts."""
    a, torch, I = env.a, env.torch, env.I
    from isscabac_b200 import multi_gpu as MG
    n_total, per = max(256, int((1 << 20) * a.config_scale)), 1024
    lo, hi = MG.shard_range(n_total, env.rank, env.world)
    n = hi - lo
    first = np.array([MG.shard_range(n_total, r, env.world)[0] for r in range(env.world)] + [n_total], dtype=np.uint32)
    g = torch.Generator(device=env.dev)
    g.manual_seed(3 + 101 * env.rank)
    sym = torch.floor(torch.log(torch.rand(n * per, generator=g, device=env.dev)) / np.log(0.5)).clamp_(0, 15).to(torch.uint8)
    off = torch.arange(n + 1, dtype=torch.int64, device=env.dev) * per
    cfg_args = (I.PROFILE_FLAT, I.BIN_EG0, 16, 3, 0, 0)
    cfg = I.make_cfg(*cfg_args)
    ctx = torch.full((8,), 1, dtype=torch.uint8, device=env.dev)
    n_bins, n_ep, op_off = _count_bins(env, cfg, sym, off)
    stride = (int((op_off[1:] - op_off[:-1]).max().item()) // 4 + 64 + 15) & ~15
    del op_off
    enc = I.encode_symbols(cfg, sym, off, ctx, slab_stride=stride)
    enc.check_overflow()
    L = I.lib()
    mg = env.mg()
    all_len = torch.empty(n_total, dtype=torch.int32, device=env.dev)
    g_off = torch.empty(n_total + 1, dtype=torch.int64, device=env.dev)
    g_scr = torch.empty(int(L.cabac_compact_scratch_bytes(n_total)), dtype=torch.uint8, device=env.dev)
    mg.gather_table(first, enc.lengths, all_lengths=all_len, byte_off=g_off, scratch=g_scr)
    total_bytes = int(g_off[-1].item())
    pb = int(g_off[int(first[env.rank + 1])].item()) - int(g_off[int(first[env.rank])].item())
    sym_buf = mg.symmetric_alloc(total_bytes + 4096)
    scratch = torch.empty(int(L.cabac_compact_scratch_bytes(n)), dtype=torch.uint8, device=env.dev)
    loc_off = torch.empty(n + 1, dtype=torch.int64, device=env.dev)
    loc_pay = torch.empty(pb + 4096, dtype=torch.uint8, device=env.dev)
    full2 = torch.empty(total_bytes + 4096, dtype=torch.uint8, device=env.dev)

    def step_fused():
        e = I.encode_symbols(cfg, sym, off, ctx, slab_stride=stride)
        mg.gather_table(first, e.lengths, all_lengths=all_len, byte_off=g_off, scratch=g_scr)
        mg.compact_p2p(first, e, g_off)
        mg.barrier()

    def step_nccl():
        e = I.encode_symbols(cfg, sym, off, ctx, slab_stride=stride)
        mg.gather_table(first, e.lengths, all_lengths=all_len, byte_off=g_off, scratch=g_scr)
        p = I.compact(e, payload=loc_pay, byte_off=loc_off, scratch=scratch)
        mg.assemble(first, g_off, p.payload, full2)

    ms_fused = env.timed(step_fused)
    ms_nccl = env.timed(step_nccl)
    ms_enc = env.timed(lambda: I.encode_symbols(cfg, sym, off, ctx, slab_stride=stride))
    ms_tab = env.timed(lambda: mg.gather_table(first, enc.lengths, all_lengths=all_len, byte_off=g_off, scratch=g_scr))
    ms_p2p = env.timed(lambda: (mg.compact_p2p(first, enc, g_off), mg.barrier()))
    torch.cuda.synchronize()
    assert int(enc.overflow[0].item()) == 0, "c4: slab / payload overflow"
    assert bool((sym_buf[:total_bytes] == full2[:total_bytes]).all().item()), "c4: fused and NCCL assemblies differ"
    # every rank decodes its own segments out of the assembled bitstream
    a0, a1 = int(first[env.rank]), int(first[env.rank + 1])
    b0 = int(g_off[a0].item())
    my_off = (g_off[a0:a1 + 1] - b0).contiguous()
    my_pay = (sym_buf[b0:b0 + pb], my_off)
    ms_dec = env.timed(lambda: I.decode_symbols(cfg, my_pay, off, ctx, sym_dtype=torch.uint8))
    dec, ok = I.decode_symbols(cfg, my_pay, off, ctx, sym_dtype=torch.uint8)
    assert bool(ok.all().item()) and bool((dec == sym).all().item()), "c4 round trip failed"
    tot_bins = env.total(n_bins)
    out = {"workload": f"{n_total} segments x {per} symbols (geometric, Nq = 16, EG0, 8 contexts reset per segment), {tot_bins / 1e9:.2f} G bins, "
                       f"sharded over {env.world} GPUs",
           "scaling": "strong", "ms": ms_fused, "gbins": tot_bins / (ms_fused * 1e-3) / 1e9,
           "step": "fused encode + length all-gather-v + device scan + compaction fused with the payload exchange (peer stores over NVLink) "
                   "+ closing barrier; the assembled bitstream is on every rank",
           "ms_with_nccl_broadcast_assembly": ms_nccl, "encode_ms": ms_enc, "length_exchange_scan_ms": ms_tab,
           "compaction_plus_exchange_ms": ms_p2p, "decode_ms": ms_dec, "decode_gbins": tot_bins / (ms_dec * 1e-3) / 1e9,
           "assembled_payload_bytes": total_bytes, "bits_per_symbol": 8.0 * total_bytes / (n_total * per)}
    out.update(env.fracs(ms_fused, (n * per + 4 * n) + 2 * pb + pb * max(env.world - 1, 0), _int_ops(n_bins - n_ep, n_ep, pb, True)))
    if env.world == 1:
        # the two-pass route beside the fused one: symbol-parallel binarizer (the two-call API: offsets, then ops) -> op-array
        # encoder; same bytes.  The binarizer is HBM work: algorithmically 2 x symbols in (once to size the op array, once to fill it) + the ops out.
        ms_bin = env.timed(lambda: I.binarize_symbols(cfg, sym, off))
        ops, op_off2 = I.binarize_symbols(cfg, sym, off)
        enc2 = I.Encoded(torch.empty((n, stride), dtype=torch.uint8, device=env.dev), torch.empty(n, dtype=torch.int32, device=env.dev),
                         torch.zeros(4, dtype=torch.int32, device=env.dev))
        ms_eops = env.timed(lambda: I.encode_ops(ops, op_off2, ctx, out=enc2))
        assert bool((enc2.lengths == enc.lengths).all().item()), "c4: fused and two-pass encoders disagree"
        for s_id in _sample_ids(n, 64):       # the bytes of a stream sample (a slab row is only defined up to its length)
            ln = int(enc.lengths[s_id].item())
            assert bool((enc2.slab[s_id, :ln] == enc.slab[s_id, :ln]).all().item()), "c4: fused and two-pass encoders disagree"
        bin_bytes = 2 * n * per + int(ops.numel()) + 16 * (n + 1)
        # ... and as ONE stream-ordered sequence when the caller knows a bound for the op array (here: the total of the
        # sizing call above): binarizer full call into the preallocated buffer + op-array encoder, no host read in between
        bscr = torch.empty(int(L.cabac_binarize_scratch_bytes(n * per, n)), dtype=torch.uint8, device=env.dev)
        opsb, offb = torch.empty_like(ops), torch.empty_like(op_off2)

        def two_pass():
            I.binarize_symbols(cfg, sym, off, ops=opsb, op_off=offb, scratch=bscr)
            I.encode_ops(opsb, offb, ctx, out=enc2)

        ms_bin1 = env.timed(lambda: I.binarize_symbols(cfg, sym, off, ops=opsb, op_off=offb, scratch=bscr))
        ms_two = env.timed(two_pass)
        assert bool((enc2.lengths == enc.lengths).all().item()) and bool((offb == op_off2).all().item()), "c4: two-pass route, one sequence"
        out["encode_two_pass_ms"] = {"binarize_two_call_api": ms_bin, "encode_ops": ms_eops,
                                     "binarize_hbm_frac": bin_bytes / (ms_bin * 1e-3) / 1e9 / env.hbm_peak,
                                     "binarize_algorithmic_bytes": bin_bytes,
                                     "binarize_full_call_known_bound": ms_bin1,
                                     "binarize_full_call_hbm_frac": bin_bytes / (ms_bin1 * 1e-3) / 1e9 / env.hbm_peak,
                                     "binarize_plus_encode_ops_known_bound": ms_two,
                                     "gbins_known_bound": tot_bins / (ms_two * 1e-3) / 1e9}
        del ops, op_off2, enc2, opsb, offb, bscr
    if env.rank == 0 and env.world == 1 and not a.no_cpu:
        out["cpu_baseline"] = _cpu_symbols(env, cfg_args, sym, off.cpu().numpy(), _sample_ids(n, 2048), np.full(8, 1, np.uint8), enc, "c4")
    mg.close_symmetric()
    if not a.no_e2e:
        # end to end through the host-buffer C ABI at symbol level: pinned symbols in -> payload + offset table out ->
        # symbols back (cabac_encode_symbols_host / cabac_decode_symbols_host, stream groups pipelined over 4 CUDA streams)
        h_sym = torch.empty(n * per, dtype=torch.uint8, pin_memory=True)
        h_sym.copy_(sym)
        h_pay = torch.empty(pb + (1 << 20), dtype=torch.uint8, pin_memory=True)
        h_back = torch.empty(n * per, dtype=torch.uint8, pin_memory=True)
        h_off = off.cpu().numpy().astype(np.uint64)
        h_boff = np.empty(n + 1, dtype=np.uint64)
        h_ctx = np.full(8, 1, dtype=np.uint8)
        np_sym, np_pay, np_back = h_sym.numpy(), h_pay.numpy(), h_back.numpy()

        def e2e_step():
            p, bo = I.encode_symbols_host(cfg, np_sym, h_off, h_ctx, payload_out=np_pay, byte_off_out=h_boff)
            I.decode_symbols_host(cfg, p, bo, h_off, h_ctx, dtype=np.uint8, out=np_back)
            return p

        p = e2e_step()
        torch.cuda.synchronize()
        if env.world > 1:
            env.dist.barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            p = e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 2
        if env.world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=env.dev)
            env.dist.all_reduce(t, op=env.dist.ReduceOp.MAX)
            dt = float(t.item())
        assert len(p) == pb and (np_back[:1 << 20] == np_sym[:1 << 20]).all(), "c4 host round trip failed"
        out["e2e"] = {"value": 2.0 * tot_bins / dt / 1e9, "unit": "Gbins/s", "ms_per_step": dt * 1e3,
                      "h2d_bytes_per_step": int(n * per + pb + 3 * (n + 1) * 8 + 16), "d2h_bytes_per_step": int(pb + (n + 1) * 8 + n * per + n),
                      "api": "cabac_encode_symbols_host + cabac_decode_symbols_host (u8 symbols in, payload + offsets out, symbols back)"}
        del h_sym, h_pay, h_back
    return out


def cfg_c5(env):
    """BASELINE configs[4]: 2^20 streams with lognormal lengths (EG2, bypass-coded suffixes), DECODE ONLY, sharded by WORK
    (balanced_ranges over the symbol counts, tile granularity).  Every rank first encodes its shard (untimed)."""
    a, torch, I = env.a, env.torch, env.I
    from isscabac_b200 import multi_gpu as MG
    n_total = max(256, int((1 << 20) * a.config_scale))
    rng = np.random.default_rng(4)
    lens = np.clip(np.round(rng.lognormal(np.log(256), 1.0, size=n_total)), 1, 65536).astype(np.int64)
    offn = np.zeros(n_total + 1, dtype=np.int64)
    np.cumsum(lens, out=offn[1:])
    parts = MG.balanced_ranges(offn, env.world)
    lo, hi = parts[env.rank]
    n = hi - lo
    my_off = offn[lo:hi + 1] - offn[lo]
    g = torch.Generator(device=env.dev)
    g.manual_seed(4 + 101 * env.rank)
    sym = torch.floor(-6.0 * torch.log(torch.rand(int(my_off[-1]), generator=g, device=env.dev))).clamp_(0, 255).to(torch.uint8)
    off = torch.as_tensor(my_off, device=env.dev)
    cfg_args = (I.PROFILE_FLAT_EPSUF, I.BIN_EG2, 256, 3, 0, 0)
    cfg = I.make_cfg(*cfg_args)
    ctx = torch.full((4,), 1, dtype=torch.uint8, device=env.dev)
    n_bins, n_ep, op_off = _count_bins(env, cfg, sym, off)
    longest_bins = int((op_off[1:] - op_off[:-1]).max().item()) if n else 0
    stride = (longest_bins // 4 + 64 + 15) & ~15
    del op_off
    enc = I.encode_symbols(cfg, sym, off, ctx, slab_stride=stride)
    pay = I.compact(enc)
    enc.check_overflow()
    pb = int(pay.byte_off[-1].item())
    ms_dec = env.timed(lambda: I.decode_symbols(cfg, pay, off, ctx, sym_dtype=torch.uint8))
    dec, ok = I.decode_symbols(cfg, pay, off, ctx, sym_dtype=torch.uint8)
    assert bool(ok.all().item()) and bool((dec == sym).all().item()), "c5 round trip failed"
    ms_enc = env.timed(lambda: I.encode_symbols(cfg, sym, off, ctx, slab_stride=stride), steps=max(2, env.steps // 2))
    # beside it: the same call with the decoder's lone SM for the longest streams switched off
    os.environ["ISSCABAC_TREE_SOLO"] = "0"
    try:
        ms_nosolo = env.timed(lambda: I.decode_symbols(cfg, pay, off, ctx, sym_dtype=torch.uint8))
        dec2, ok2 = I.decode_symbols(cfg, pay, off, ctx, sym_dtype=torch.uint8)
        assert bool(ok2.all().item()) and bool((dec2 == sym).all().item()), "c5 round trip failed (one launch)"
        del dec2, ok2
    finally:
        os.environ.pop("ISSCABAC_TREE_SOLO", None)
    tot_bins = env.total(n_bins)
    work = [int(offn[b] - offn[a_]) for a_, b in parts]
    out = {"workload": f"{n_total} streams, lognormal lengths (median 256, up to {int(lens.max())} symbols), EG2 with bypass suffixes, "
                       f"{tot_bins / 1e9:.2f} G bins, work-balanced over {env.world} GPUs",
           "scaling": "strong", "ms": ms_dec, "gbins": tot_bins / (ms_dec * 1e-3) / 1e9,
           "step": "decode only (fused decoder + debinarizer, persistent warps, longest streams first)",
           "encode_ms": ms_enc, "encode_gbins": tot_bins / (ms_enc * 1e-3) / 1e9,
           "ms_one_launch": ms_nosolo, "gbins_one_launch": tot_bins / (ms_nosolo * 1e-3) / 1e9,
           "lone_sm": "the 128 longest streams of the call run on an SM of their own (a second launch of the same kernel beside the main "
                      "one; ISSCABAC_TREE_SOLO=0 gives the one-launch form measured as ms_one_launch; profiles/r2_tree_solo_experiment.txt)",
           "streams_per_rank": [b - a_ for a_, b in parts], "symbols_per_rank_max_over_mean": max(work) / (sum(work) / len(work)),
           "longest_stream_symbols": int(lens.max()), "longest_stream_bins_this_rank": longest_bins,
           "limit": "the longest stream: a stream is a serial chain, the launch cannot end before its longest stream does "
                    "(longest_stream_bins x cycles per step of a lone lane); more GPUs shorten everything but that"}
    out.update(env.fracs(ms_dec, pb + int(my_off[-1]) + n, _int_ops(n_bins - n_ep, n_ep, pb, False)))
    if env.rank == 0 and env.world == 1 and not a.no_cpu:
        out["cpu_baseline"] = _cpu_symbols(env, cfg_args, sym, my_off, _sample_ids(n, 4096), np.full(4, 1, np.uint8), enc, "c5")
    return out


def run_configs(env, headline):
    """-> {name: result}; every rank runs every config (they contain collectives), rank 0 reports."""
    torch = env.torch
    want = [c.strip() for c in env.a.configs.split(",") if c.strip()]
    fns = {"c3_strong": lambda: cfg_c3_strong(env, headline), "c2": lambda: cfg_c2(env), "c4": lambda: cfg_c4(env), "c5": lambda: cfg_c5(env)}
    out = {}
    for name in want:
        if name not in fns:
            continue
        t0 = time.time()
        out[name] = fns[name]()
        out[name]["n_gpus"] = env.world
        out[name]["wall_s"] = round(time.time() - t0, 1)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist

    import isscabac_b200 as I

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S, B = a.streams, a.bins
    total_bins = S * B

    ops = gen_ops_device(torch, a.seed + 1000 * rank, S, B, dev)
    op_off = (torch.arange(S + 1, dtype=torch.int64, device=dev) * B)
    ctx = torch.full((N_CTX,), 1, dtype=torch.uint8, device=dev)
    stride = (B // 4 + 64 + 15) & ~15
    enc = I.Encoded(torch.empty((S, stride), dtype=torch.uint8, device=dev),
                    torch.empty(S, dtype=torch.int32, device=dev), torch.zeros(4, dtype=torch.int32, device=dev))
    L = I.lib()
    scratch = torch.empty(int(L.cabac_compact_scratch_bytes(S)), dtype=torch.uint8, device=dev)
    byte_off = torch.empty(S + 1, dtype=torch.int64, device=dev)
    payload = torch.empty(S * (B // 6 + 64), dtype=torch.uint8, device=dev)   # > 1.3 bit/bin: never reached by adaptive CABAC
    bins = torch.empty(total_bins, dtype=torch.uint8, device=dev)
    ok = torch.empty(S, dtype=torch.uint8, device=dev)
    # N > 1: the length exchange + global scan go through the library's own NCCL path (cabac_multi_gpu_gather_table)
    gathered = g_off = g_scr = mg = first = None
    if world > 1:
        from isscabac_b200 import multi_gpu as MG
        mg = MG.default_handle()
        first = (np.arange(world + 1, dtype=np.uint64) * S).astype(np.uint32)
        gathered = torch.empty(world * S, dtype=torch.int32, device=dev)
        g_off = torch.empty(world * S + 1, dtype=torch.int64, device=dev)
        g_scr = torch.empty(int(L.cabac_compact_scratch_bytes(world * S)), dtype=torch.uint8, device=dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = []

    def step(record=False):
        e = [ev() for _ in range(4)] if record else None
        if record:
            e[0].record()
        I.encode_ops(ops, op_off, ctx, out=enc)
        if record:
            e[1].record()
        pay = I.compact(enc, payload=payload, byte_off=byte_off, scratch=scratch)
        if world > 1:
            mg.gather_table(first, enc.lengths, all_lengths=gathered, byte_off=g_off, scratch=g_scr)
        if record:
            e[2].record()
        I.decode_ops(pay, ops, op_off, ctx, bins=bins, finish_ok=ok)
        if record:
            e[3].record()
            marks.append(e)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize()
    # parity inside the bench: round trip + all finish checks (size-independent properties)
    assert int(enc.overflow[0].item()) == 0, "slab/payload overflow"
    assert bool(ok.all().item()), "decoder finish() check failed"
    assert bool(((ops & 1) == bins).all().item()), "decoded bins differ from the encoded ones"
    payload_bytes = int(byte_off[-1].item())

    if rank == 0:
        sampler.wait_first()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_start, t_end = ev(), ev()
    wall0 = time.time()
    t_start.record()
    for _ in range(a.steps):
        step(record=True)
    t_end.record()
    torch.cuda.synchronize()
    wall1 = time.time()
    if world > 1:
        dist.barrier()
    clocks = None
    if rank == 0:
        time.sleep(0.05)
        sampler.stop()
        clocks = sampler.summary(wall0, wall1)
    ms = t_start.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_enc = float(np.mean([m[0].elapsed_time(m[1]) for m in marks]))
    ms_cmp = float(np.mean([m[1].elapsed_time(m[2]) for m in marks]))
    ms_dec = float(np.mean([m[2].elapsed_time(m[3]) for m in marks]))
    value = 2.0 * total_bins * world * a.steps / (ms * 1e-3) / 1e9

    # ---- end to end through the host-buffer C ABI (pinned host arrays) ------------------
    e2e = e2e_u8 = None
    if not a.no_e2e:
        numa = bind_to_gpu_numa(local)
        pcie = pcie_bandwidth(torch, dev)
        pcie_all = pcie_bandwidth_concurrent(torch, dist, dev, world) if world > 1 else None
        h_ops = torch.empty(total_bins, dtype=torch.uint8, pin_memory=True)
        h_ops.copy_(ops)
        h_off = np.arange(S + 1, dtype=np.uint64) * np.uint64(B)
        h_ctx = np.full(N_CTX, 1, dtype=np.uint8)
        h_pay = torch.empty(payload_bytes + 4096, dtype=torch.uint8, pin_memory=True)
        h_boff = np.empty(S + 1, dtype=np.uint64)
        h_bins = torch.empty(total_bins, dtype=torch.uint8, pin_memory=True)
        h_ok = np.empty(S, dtype=np.uint8)
        np_ops, np_pay, np_bins = h_ops.numpy(), h_pay.numpy(), h_bins.numpy()

        h_bits = torch.empty((total_bins + 7) // 8, dtype=torch.uint8, pin_memory=True)
        np_bits = h_bits.numpy()

        def e2e_step(packed):
            p, bo = I.encode_ops_host(np_ops, h_off, h_ctx, payload_out=np_pay, byte_off_out=h_boff)
            I.decode_ops_host(p, bo, np_ops, h_off, h_ctx, bins_out=np_bits if packed else np_bins, packed=packed)
            return p

        def e2e_run(packed):
            e2e_step(packed)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(a.e2e_steps):
                p = e2e_step(packed)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            h2d = total_bins + (S + 1) * 8 + N_CTX            # encode: ops, offsets, ctx
            h2d += total_bins + len(p) + 2 * (S + 1) * 8 + N_CTX   # decode: op kinds, payload, both offset tables, ctx
            d2h = len(p) + (S + 1) * 8 + (len(np_bits) if packed else total_bins) + S        # payload, offsets, bins, finish flags
            return {"value": 2.0 * total_bins * world * a.e2e_steps / dt / 1e9, "unit": "Gbins/s",
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": a.e2e_steps,
                    "ms_per_step": dt / a.e2e_steps * 1e3,
                    "gbytes_per_s_over_pcie": (h2d + d2h) / (dt / a.e2e_steps) / 1e9,
                    "api": "cabac_encode_ops_host + " + ("cabac_decode_ops_host_packed (decoded bins bit-packed, 8 per byte)" if packed
                                                         else "cabac_decode_ops_host (decoded bins one per byte)")}

        e2e_u8 = e2e_run(False)
        assert (np_bins[:1 << 20] == (np_ops[:1 << 20] & 1)).all()
        e2e = e2e_run(True)
        assert (np.unpackbits(np_bits[:1 << 17], bitorder="little") == (np_ops[:1 << 20] & 1)).all()
        for e in (e2e, e2e_u8):
            # the roof of the call pair: its host-to-device bytes at the measured H2D bandwidth (the larger direction)
            e["pcie_gbs_measured"] = pcie
            e["frac_of_pcie"] = (e["h2d_bytes_per_step"] / (e["ms_per_step"] * 1e-3) / 1e9) / pcie["h2d"]
            if pcie_all:
                # N > 1: every rank moves its bytes through the same host; the roof is what the ranks reach copying together
                e["pcie_gbs_all_ranks_concurrent"] = pcie_all
                e["frac_of_pcie_concurrent"] = (e["h2d_bytes_per_step"] / (e["ms_per_step"] * 1e-3) / 1e9) / pcie_all["h2d_per_rank"]
        e2e["numa"] = numa
        e2e["note"] = ("host buffers in, host buffers out, every copy inside the timed calls; PCIe-bound: frac_of_pcie = achieved "
                       "host-to-device GB/s over the H2D bandwidth measured in this run (at N > 1 the ranks share the host's PCIe / DRAM)")
        del h_bits
        del h_ops, h_bins, h_pay

    # ---- CPU baseline beside it (rank 0 only, N=1 only): bounded sample + byte parity -------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        threads = host_threads()
        n_s = int(min(S, max(64, min(threads * 16, 4096))))
        # stratified sample: runs of 16 streams spread over the whole stream range (first and last streams included), so
        # the byte comparison covers first / middle / last tiles of CTAs all over the grid, not a prefix
        run = 16 if S >= 16 else 1
        n_runs = max(1, n_s // run)
        starts = np.unique(np.round(np.linspace(0, max(S - run, 0), n_runs)).astype(np.int64) // 1 )
        ids = np.unique(np.concatenate([np.arange(st, min(st + run, S)) for st in starts]))
        n_s = int(ids.size)
        idt = torch.as_tensor(ids, device=dev)
        ops_s = ops.view(S, B)[idt].reshape(-1).cpu().numpy()
        te, td, kind, slab_ref, lens_ref = cpu_roundtrip(ops_s, n_s, B, threads)
        lens_gpu = enc.lengths[idt].cpu().numpy().astype(np.uint32)
        assert (lens_gpu == lens_ref).all(), "GPU stream lengths differ from the reference"
        w = int(lens_ref.max())
        live = np.arange(w)[None, :] < lens_ref[:, None]       # bytes past a stream's length are not part of it
        assert (enc.slab[idt][:, :w].cpu().numpy()[live] == slab_ref[:, :w][live]).all(), "GPU bytes differ from the reference"
        cpu = {"value": 2.0 * n_s * B / (te + td) / 1e9, "unit": "Gbins/s", "cores": threads, "kind": kind,
               "encode_gbins": n_s * B / te / 1e9, "decode_gbins": n_s * B / td / 1e9,
               "encode_mbins_per_core": n_s * B / te / 1e6 / threads, "decode_mbins_per_core": n_s * B / td / 1e6 / threads,
               "sample": f"{n_s} of {S} streams x {B} bins ({len(starts)} runs of {run} streams spread evenly over the stream range), "
                         f"encode+decode, one stream per core: {threads} single-threaded processes; GPU output of these streams "
                         f"verified byte-identical"}

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    # ---- the other BASELINE configurations + strong scaling (all ranks: they hold collectives) -------------
    configs = None
    if not a.no_configs:
        f_sm0 = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
        if world > 1:   # every rank needs the same clock figure for its fractions
            t = torch.tensor([f_sm0 if rank == 0 else 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            f_sm0 = float(t.item())
        enc_b = total_bins + payload_bytes + 4 * S
        dec_b = 2 * total_bins + payload_bytes + S
        headline = {"workload": f"{S} streams x {B} bins", "scaling": "strong", "ms": ms / a.steps,
                    "gbins": 2.0 * total_bins / (ms / a.steps * 1e-3) / 1e9, "encode_ms": ms_enc, "decode_ms": ms_dec,
                    "step": "encode + length scan + compaction + decode"}
        ops = bins = enc = payload = byte_off = scratch = ok = gathered = g_off = g_scr = None   # (the step closures see the same cells)
        torch.cuda.empty_cache()
        env = _Env(a, torch, dist, I, dev, rank, world, hbm_peak, f_sm0)
        headline.update(env.fracs(ms / a.steps, enc_b + dec_b, _int_ops(total_bins * (1 - P_EP), total_bins * P_EP, payload_bytes, True) +
                                  _int_ops(total_bins * (1 - P_EP), total_bins * P_EP, payload_bytes, False)))
        if cpu is not None:
            headline["cpu_baseline"] = cpu
        configs = run_configs(env, headline)

    if _POOL is not None:
        _POOL.close()
        _POOL.join()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel --------------------------------------------------
    enc_bytes = total_bins + payload_bytes + 4 * S          # 1 B/bin in + payload out + 4 B/stream
    dec_bytes = total_bins + payload_bytes + total_bins + S  # kinds in + payload in + 1 B/bin out + flag
    # the dominant kernel: the longer one; the two are within a few per cent of each other and which one is ahead changes
    # from run to run, so inside 5 % it is the decoder -- it moves twice the bytes (both are in roofline.kernels either way)
    enc_kernel = (I.lib().cabac_encode_ops_kernel(S, 23) or b"k_encode_ops_wide").decode()     # the formulation this shape runs
    dec_kernel = (I.lib().cabac_decode_ops_kernel(S, 23) or b"k_decode_ops_wide").decode()
    dom = enc_kernel if ms_enc > 1.05 * ms_dec else dec_kernel
    dom_ms, dom_bytes = (ms_enc, enc_bytes) if dom == enc_kernel else (ms_dec, dec_bytes)
    ach = dom_bytes / (dom_ms * 1e-3) / 1e9
    # DRAM traffic of that kernel per launch, from the committed ncu --set full capture
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)[dom]
        if tj.get("csrc_hash") != hot_kernel_hash():
            raise ValueError("stale")
        traffic = float(tj["dram_bytes_per_launch"])
        traffic_src = tj["source"]
        if int(tj["bins_per_launch"]) != total_bins:   # captured at another size: bytes scale with the bins
            traffic *= total_bins / float(tj["bins_per_launch"])
            traffic_src += f" (scaled from {tj['bins_per_launch']} bins per launch)"
    except ValueError:
        traffic_src = "profiles/traffic.json was captured from other kernel sources (csrc hash mismatch): re-run tools/profile_round.sh"
    except Exception:
        pass
    # both hot kernels side by side (they are within a few per cent of each other, so which one is "dominant" can change
    # from run to run; the decoder moves twice the bytes per bin)
    def _traffic(name):
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                t = json.load(f)[name]
            if t.get("csrc_hash") != hot_kernel_hash():
                return None
            return float(t["dram_bytes_per_launch"]) * (total_bins / float(t["bins_per_launch"]))
        except Exception:
            return None
    per_kernel = {n: {"ms": m, "algorithmic_bytes": b, "achieved": b / (m * 1e-3) / 1e9, "frac": b / (m * 1e-3) / 1e9 / hbm_peak, "traffic": _traffic(n)}
                  for n, m, b in ((enc_kernel, ms_enc, enc_bytes), (dec_kernel, ms_dec, dec_bytes))}
    # integer-issue roofline (the binding one): algorithmic int32 ops
    sm, _, _ = (torch.cuda.get_device_properties(dev).multi_processor_count, 0, 0)
    f_sm = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    ctx_bins, ep_bins = total_bins * (1 - P_EP), total_bins * P_EP
    enc_ops = ctx_bins * 16 + ep_bins * 6 + payload_bytes * 12
    dec_ops = ctx_bins * 16 + ep_bins * 6 + payload_bytes * 4
    lanes_alu, lanes_both, int_src = int_peak_lanes()
    int_peak = sm * lanes_both * f_sm * 1e6
    roof_int = {"bound": "int32-issue", "unit": "Tops/s", "peak": int_peak / 1e12, "peak_note":
                f"{sm} SMs x {lanes_both:.1f} int32 lanes/clk/SM (ALU + FMA pipes, dependent-free IADD3:IMAD 1:1) x {f_sm:.0f} MHz "
                f"(median SM clock in the timed region); {int_src}",
                "peak_alu_pipe": sm * lanes_alu * f_sm * 1e6 / 1e12,
                "peak_alu_pipe_note": f"ALU pipe alone: {lanes_alu:.1f} lanes/clk/SM (IADD3 / LOP3 / SHF / PRMT / ISETP+SEL all measure the same)",
                "encode": {"achieved": enc_ops / (ms_enc * 1e-3) / 1e12, "frac": enc_ops / (ms_enc * 1e-3) / int_peak,
                           "frac_alu_pipe": enc_ops / (ms_enc * 1e-3) / (sm * lanes_alu * f_sm * 1e6)},
                "decode": {"achieved": dec_ops / (ms_dec * 1e-3) / 1e12, "frac": dec_ops / (ms_dec * 1e-3) / int_peak,
                           "frac_alu_pipe": dec_ops / (ms_dec * 1e-3) / (sm * lanes_alu * f_sm * 1e6)},
                "ops_per_bin": "16/ctx bin, 6/bypass bin, 12 (enc) | 4 (dec) per payload byte (SURVEY.md 8(d))"}
    out = {
        "metric": "CABAC encode+decode throughput over independent streams",
        "value": value, "unit": "Gbins/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(a), "streams_per_gpu": S, "bins_per_stream": B,
                   "l2": "inputs (4 GiB of ops per GPU) are larger than L2; no flush needed",
                   "step": "encode + length scan + compaction + decode" +
                           (" + all-gather-v of lengths and global scan (cabac_multi_gpu_gather_table, NCCL)" if world > 1 else "")},
        "encode_gbins": total_bins * world / (ms_enc * 1e-3) / 1e9,
        "decode_gbins": total_bins * world / (ms_dec * 1e-3) / 1e9,
        "kernel_ms": {enc_kernel: ms_enc, "k_scan_init+k_scan_u32_u64+k_compact_copy": ms_cmp, dec_kernel: ms_dec},
        "payload_bytes_per_gpu": payload_bytes, "bits_per_bin": 8.0 * payload_bytes / total_bins,
        "gpu_launches": (5 + (2 if world > 1 else 0)) * a.steps,   # encode, scan_init, scan, compact_copy, decode (+ the global scan at N > 1)
        "clocks": clocks,
        "e2e": e2e,
        "e2e_u8_bins": e2e_u8,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                     "frac": ach / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes": dom_bytes, "peak_source": peak_src, "kernels": per_kernel,
                     "note": "the path is integer-issue bound, not HBM bound: see roofline_int"},
        "roofline_int": roof_int,
        "cpu_baseline": cpu,
        "configs": configs,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    # stdout carries exactly ONE line, the JSON result: whatever libraries print there (NCCL's version banner at communicator
    # creation, for one) is sent to stderr; the result line is written to the real stdout at the end
    sys.stdout.flush()
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    _real_stdout.flush()
