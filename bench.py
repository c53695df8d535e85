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

  python bench.py [--gpus N] [--steps K] [--warmup W] [--streams S] [--bins B]
  python bench.py --impl reference ...   # the reference CPU engine (oracle/_ref) on host cores
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
def cpu_roundtrip(ops, n_streams, n_bins, threads):
    """encode + decode `ops` with the reference engine on `threads` host threads.
    -> (seconds_enc, seconds_dec, kind, slab, lens)"""
    import oracle as O
    impl, kind = ("ref", "reference") if O.ref() is not None else ("oracle", "port")
    off = (np.arange(n_streams + 1, dtype=np.uint64) * np.uint64(n_bins))
    ci = np.full(N_CTX, 1, dtype=np.uint8)
    stride = n_bins // 4 + 64
    t0 = time.perf_counter()
    slab, lens = O.encode_ops(ops, off, ci, out_stride=stride, n_threads=threads, impl=impl)
    t1 = time.perf_counter()
    payload, boff = O.compact(slab, lens)
    t2 = time.perf_counter()
    bins, ok = O.decode_ops(payload, boff, ops, off, ci, n_threads=threads, impl=impl)
    t3 = time.perf_counter()
    assert ok.all() and (bins == (ops & 1)).all(), "reference round trip failed"
    return t1 - t0, t3 - t2, kind, slab, lens


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
    times = []
    kind = "port"
    for i in range(a.warmup + a.steps):
        te, td, kind, _, _ = cpu_roundtrip(ops, n_streams, a.bins, threads)
        if i >= a.warmup:
            times.append(te + td)
    t = float(np.mean(times))
    bins = 2.0 * n_streams * a.bins
    value = bins / t / 1e9
    sample = f"{n_streams} of {a.streams} streams x {a.bins} bins, encode+decode, one stream per thread, files on tmpfs"
    print(json.dumps({
        "impl": "reference", "metric": "CABAC encode+decode throughput over independent streams",
        "value": value, "unit": "Gbins/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Gbins/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Gbins/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


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
    gathered = torch.empty(world * S, dtype=torch.int32, device=dev) if world > 1 else None

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
            dist.all_gather_into_tensor(gathered, enc.lengths)
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
    e2e = None
    if not a.no_e2e:
        h_ops = torch.empty(total_bins, dtype=torch.uint8, pin_memory=True)
        h_ops.copy_(ops)
        h_off = np.arange(S + 1, dtype=np.uint64) * np.uint64(B)
        h_ctx = np.full(N_CTX, 1, dtype=np.uint8)
        h_pay = torch.empty(payload_bytes + 4096, dtype=torch.uint8, pin_memory=True)
        h_boff = np.empty(S + 1, dtype=np.uint64)
        h_bins = torch.empty(total_bins, dtype=torch.uint8, pin_memory=True)
        h_ok = np.empty(S, dtype=np.uint8)
        np_ops, np_pay, np_bins = h_ops.numpy(), h_pay.numpy(), h_bins.numpy()

        def e2e_step():
            p, bo = I.encode_ops_host(np_ops, h_off, h_ctx, payload_out=np_pay, byte_off_out=h_boff)
            I.decode_ops_host(p, bo, np_ops, h_off, h_ctx, bins_out=np_bins)
            return p

        e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            p = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert (np_bins[:1 << 20] == (np_ops[:1 << 20] & 1)).all()
        h2d = total_bins + (S + 1) * 8 + N_CTX            # encode: ops, offsets, ctx
        h2d += total_bins + len(p) + 2 * (S + 1) * 8 + N_CTX   # decode: op kinds, payload, both offset tables, ctx
        d2h = len(p) + (S + 1) * 8 + total_bins + S        # payload, offsets, bins, finish flags
        e2e = {"value": 2.0 * total_bins * world * a.e2e_steps / dt / 1e9, "unit": "Gbins/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": a.e2e_steps,
               "ms_per_step": dt / a.e2e_steps * 1e3}
        del h_ops, h_bins, h_pay

    # ---- CPU baseline beside it (rank 0 only, N=1 only): bounded sample + byte parity -------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        threads = host_threads()
        n_s = int(min(S, max(64, min(threads * 16, 4096))))
        ops_s = ops[: n_s * B].cpu().numpy()
        te, td, kind, slab_ref, lens_ref = cpu_roundtrip(ops_s, n_s, B, threads)
        lens_gpu = enc.lengths[:n_s].cpu().numpy().astype(np.uint32)
        assert (lens_gpu == lens_ref).all(), "GPU stream lengths differ from the reference"
        w = int(lens_ref.max())
        assert (enc.slab[:n_s, :w].cpu().numpy() == slab_ref[:, :w]).all(), "GPU bytes differ from the reference"
        cpu = {"value": 2.0 * n_s * B / (te + td) / 1e9, "unit": "Gbins/s", "cores": threads, "kind": kind,
               "encode_gbins": n_s * B / te / 1e9, "decode_gbins": n_s * B / td / 1e9,
               "sample": f"first {n_s} of {S} streams x {B} bins, encode+decode, one stream per thread; "
                         f"GPU output of these streams verified byte-identical"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel --------------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    enc_bytes = total_bins + payload_bytes + 4 * S          # 1 B/bin in + payload out + 4 B/stream
    dec_bytes = total_bins + payload_bytes + total_bins + S  # kinds in + payload in + 1 B/bin out + flag
    dom = "k_encode_ops_wide" if ms_enc >= ms_dec else "k_decode_ops_wide"
    dom_ms, dom_bytes = (ms_enc, enc_bytes) if dom == "k_encode_ops_wide" else (ms_dec, dec_bytes)
    ach = dom_bytes / (dom_ms * 1e-3) / 1e9
    # DRAM traffic of that kernel per launch, from the committed ncu --set full capture
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)[dom]
        traffic = float(tj["dram_bytes_per_launch"])
        traffic_src = tj["source"]
        if int(tj["bins_per_launch"]) != total_bins:   # captured at another size: bytes scale with the bins
            traffic *= total_bins / float(tj["bins_per_launch"])
            traffic_src += f" (scaled from {tj['bins_per_launch']} bins per launch)"
    except Exception:
        pass
    # integer-issue roofline (the binding one): algorithmic int32 ops
    sm, _, _ = (torch.cuda.get_device_properties(dev).multi_processor_count, 0, 0)
    f_sm = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    ctx_bins, ep_bins = total_bins * (1 - P_EP), total_bins * P_EP
    enc_ops = ctx_bins * 16 + ep_bins * 6 + payload_bytes * 12
    dec_ops = ctx_bins * 16 + ep_bins * 6 + payload_bytes * 4
    int_peak = sm * 64 * f_sm * 1e6       # 64 int32 lanes/clk/SM on the ALU pipe (B300_MICROARCH: rt_SMSP = 2)
    roof_int = {"bound": "int32-issue", "unit": "Tops/s", "peak": int_peak / 1e12, "peak_note":
                f"{sm} SMs x 64 int32 lanes/clk (ALU pipe) x {f_sm:.0f} MHz (median SM clock in the timed region)",
                "encode": {"achieved": enc_ops / (ms_enc * 1e-3) / 1e12, "frac": enc_ops / (ms_enc * 1e-3) / int_peak},
                "decode": {"achieved": dec_ops / (ms_dec * 1e-3) / 1e12, "frac": dec_ops / (ms_dec * 1e-3) / int_peak},
                "ops_per_bin": "16/ctx bin, 6/bypass bin, 12 (enc) | 4 (dec) per payload byte (SURVEY.md 8(d))"}
    out = {
        "metric": "CABAC encode+decode throughput over independent streams",
        "value": value, "unit": "Gbins/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(a), "streams_per_gpu": S, "bins_per_stream": B,
                   "l2": "inputs (4 GiB of ops per GPU) are larger than L2; no flush needed",
                   "step": "encode + length scan + compaction + decode" + (" + all-gather of lengths" if world > 1 else "")},
        "encode_gbins": total_bins * world / (ms_enc * 1e-3) / 1e9,
        "decode_gbins": total_bins * world / (ms_dec * 1e-3) / 1e9,
        "kernel_ms": {"k_encode_ops_wide": ms_enc, "k_scan_init+k_scan_u32_u64+k_compact_copy": ms_cmp, "k_decode_ops_wide": ms_dec},
        "payload_bytes_per_gpu": payload_bytes, "bits_per_bin": 8.0 * payload_bytes / total_bins,
        "gpu_launches": 5 * a.steps,
        "clocks": clocks,
        "e2e": e2e,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                     "frac": ach / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes": dom_bytes, "peak_source": peak_src,
                     "note": "the path is integer-issue bound, not HBM bound: see roofline_int"},
        "roofline_int": roof_int,
        "cpu_baseline": cpu,
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
