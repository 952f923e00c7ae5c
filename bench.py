#!/usr/bin/env python
"""bench.py -- latent symbols/s, encode + decode, bit-exact (BASELINE.json metric).

A step = one pass of the whole hot path over one batch of synthetic W+ latents:
  quantise -> encode -> compact -> decode (+ fused dequantise).
Headline workload = BASELINE.json configs[1]: 1024 latents of 16x512 at 8 bits per GPU ("enc_like" synthetic
latents).  With N GPUs every rank codes its own shard (no collective on the hot path; the
only NCCL traffic is the optional size gather after the timed region).  The same run then measures the other
BASELINE.json configs (`sweep`: cfg3 at 4/8/10 bits, cfg4 as stated = 65,536 latents over the N GPUs, cfg5 =
hierarchical latents through quantiser B), each gated on bit-exact parity with the CPU oracle, and the HBM-bound
quantiser kernels standalone at cfg4 size (`hbm_kernels`).

  python bench.py [--gpus N --steps K --warmup W]          this framework
  python bench.py --impl reference [...]                    the CPU coder on the host cores

One JSON line on stdout (rank 0).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "latent symbols/sec enc+dec (bit-exact)"
UNIT = "symbols/s"
R, C = 16, 512
SYMS = R * C
SM_CLOCK_HZ = 1.965e9  # clocks.max.sm of the B200 (profiling guide); the issue-slot peak is 148 SMs x 4 schedulers x this


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="streams per GPU of the headline workload")
    ap.add_argument("--bits", type=int, default=8)
    ap.add_argument("--kind", default="enc_like")
    ap.add_argument("--cpu-sample-streams", type=int, default=0, help="0 = auto (about 10-30 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the cfg3/cfg4/cfg5 workloads and the HBM kernels")
    ap.add_argument("--sweep-steps", type=int, default=10)
    return ap.parse_args()


HIER_SIGMA = [0.115] * 5 + [0.158] * 7 + [0.131] * 4  # measured random-init per-level stds


def synth(kind, B, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    if kind == "enc_like":
        return torch.randn(B, R, C, generator=g) * 0.14
    if kind == "wide":
        return torch.randn(B, R, C, generator=g) * 0.4
    if kind == "uniform":
        return torch.rand(B, R, C, generator=g) * 2 - 1
    if kind == "hier":
        return torch.randn(B, R, C, generator=g) * torch.tensor(HIER_SIGMA).view(1, R, 1)
    raise ValueError(kind)


def workload_name(args):
    return "cfg2: %d synthetic W+ latents 16x512 (%s) per GPU, %d-bit codebook quantise + CABAC round trip" % (
        args.batch, args.kind, args.bits)


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (oracle/latent_oracle.c) on the host cores -- a reported baseline
# ------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def _cpu_worker(task):
    """Runs in a worker process: encode+decode a chunk of streams with the C oracle."""
    codes_chunk, n = task
    from oracle import oracle as orc
    bad = 0
    for i in range(codes_chunk.shape[0]):
        st, _ = orc.roundtrip_stream(codes_chunk[i:i + 1], n, "repaired")
        bad += int(st != 0)
    return bad


_POOL = None


def _pool(procs):
    """One process per host core (threads do not scale inside sandboxed containers: page faults
    of one address space serialise).  'spawn' so the children never inherit a CUDA context."""
    global _POOL
    if _POOL is None:
        import multiprocessing as mp

        import numpy as np
        from oracle import oracle as orc
        orc.build()
        _POOL = mp.get_context("spawn").Pool(procs)
        _POOL.map(_cpu_worker, [(np.zeros((1, 2, 8), np.int32), 16)] * procs)  # load the library everywhere
    return _POOL


def _close_pool():
    global _POOL
    if _POOL is not None:
        _POOL.close()
        _POOL.join()
        _POOL = None


def cpu_roundtrip_rate(codes_np, n, procs, repeats=1):
    """codes_np int32 [S,16,512]: encode+decode every stream with the C oracle on `procs` host
    processes. Returns (symbols/s, seconds, streams)."""
    pool = _pool(procs)
    S = codes_np.shape[0]
    tasks = [(codes_np[i:i + 1], n) for i in range(S)]
    t0 = time.perf_counter()
    bad = 0
    for _ in range(repeats):
        bad += sum(pool.map(_cpu_worker, tasks, chunksize=max(1, S // (procs * 4))))
    dt = time.perf_counter() - t0
    assert bad == 0, "CPU oracle round trip failed"
    return S * repeats * SYMS / dt, dt, S * repeats


def cpu_sample(args, n):
    """A bounded sample of the bench workload for the CPU legs: quantised with the oracle."""
    import torch

    from oracle import oracle as orc
    threads = host_threads()
    S = args.cpu_sample_streams
    if S <= 0:
        S = int(min(args.batch, 1024))
    lat = synth(args.kind, S, 1000 + 2 * 100000).numpy()
    cb = torch.linspace(-1, 1, n).float().numpy()
    return orc.quantize_codebook(lat, cb), threads


def cpu_repeats(S, n, threads, target_s=12.0):
    """How many passes over the S sample streams give about target_s seconds of wall clock."""
    per = {16: 0.003, 256: 0.03, 1024: 0.13}.get(n, 0.03)  # core-seconds per stream (enc+dec), measured
    one_pass = S * per / max(threads, 1)
    return int(max(1, min(64, round(target_s / max(one_pass, 1e-3)))))


def run_reference(args):
    """--impl reference: the reference's CPU path (C port of it -- the reference itself is Python
    and cannot travel to the GPU box) on all host threads; each step = a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 1 << args.bits
    codes, threads = cpu_sample(args, n)
    for _ in range(min(args.warmup, 1)):
        cpu_roundtrip_rate(codes[: max(2, threads)], n, threads)
    t_total, streams = 0.0, 0
    reps = cpu_repeats(codes.shape[0], n, threads, target_s=max(2.0, 20.0 / max(args.steps, 1)))
    for _ in range(args.steps):
        _, dt, s = cpu_roundtrip_rate(codes, n, threads, repeats=reps)
        t_total += dt
        streams += s
    value = streams * SYMS / t_total
    sample = "%d streams x %d passes per step x %d steps of the bench workload (seeded %s latents), C port of the reference coder, one process per core, %s" % (
        codes.shape[0], reps, args.steps, args.kind, cpu_model())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample_streams_per_step": int(codes.shape[0])},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for ln in out.strip().splitlines():
                    self.rows.append([c.strip() for c in ln.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# this framework
# ------------------------------------------------------------------------------------------------
def kernel_source_hash():
    """Identifies the kernel sources a set of ncu-derived facts belongs to (profiles/*_kernel_facts.json)."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "image_compression_2_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def load_kernel_facts():
    """ncu-derived per-kernel facts (instructions per symbol, DRAM bytes): they cannot be measured outside a profiler,
    so they come from the newest committed capture; `current` says whether that capture was taken from exactly the
    kernel sources this run uses."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if f.endswith("_kernel_facts.json"):
            best = os.path.join(pdir, f)
    if best is None:
        return {}, None
    try:
        return json.load(open(best)), os.path.relpath(best, ROOT)
    except Exception:
        return {}, None


def oracle_parity(idx_host, streams, nbits_h, n, sample):
    """Bit-for-bit comparison of `sample` bitstreams with the CPU oracle; returns (count, oracle bits/symbol)."""
    from oracle import oracle as orc
    B = idx_host.shape[0]
    picks = sorted(set(int(round(i * (B - 1) / max(sample - 1, 1))) for i in range(min(sample, B))))
    obits = 0
    for b in picks:
        ref = orc.encode_stream(idx_host[b:b + 1].astype("int32"), n, "repaired")
        assert ref["status"] == 0, "oracle encoder fault on stream %d" % b
        assert int(nbits_h[b]) == ref["nbits"], "stream %d: %d coded bits, oracle %d" % (b, int(nbits_h[b]), ref["nbits"])
        assert streams[b] == ref["packed"], "stream %d: bitstream differs from the oracle" % b
        obits += ref["nbits"]
    return picks, obits / (len(picks) * idx_host.shape[1] * idx_host.shape[2])


def run_workload(pipe, lat_host, dev, steps, warmup, flush, barrier, want_e2e, rank, tag):
    """Parity gate + device-resident timing (+ optional end-to-end timing) of one workload on this rank.
    Returns a dict of LOCAL numbers: ms_total, e2e_s, launches, coded_bits, parity strings."""
    import torch
    from image_compression_2_b200 import _native
    lib = _native.load()
    B = lat_host.shape[0]
    n = pipe.n
    lat = lat_host.to(dev)
    stream = torch.cuda.current_stream()

    def step(events=None):
        idx = pipe.quantize(lat)
        enc = pipe.encode(idx, reuse_output=True)
        if events is not None:
            events[0].record(stream)
        dec_idx, deq, dstatus, _ = pipe.decode(enc.data, enc.offsets, enc.nbits, B, reuse_output=True)
        if events is not None:
            events[1].record(stream)
        return idx, enc, dec_idx, deq, dstatus

    # ---- parity gate on this rank's actual workload before any timing
    idx, enc, dec_idx, deq, dstatus = step()
    torch.cuda.synchronize()
    assert int(enc.status.abs().sum()) == 0 and int(dstatus.abs().sum()) == 0, "%s: coder fault" % tag
    assert torch.equal(dec_idx.view(B, R, C), idx), "%s: round trip is not the identity" % tag
    assert torch.equal(deq.view(B, R, C), pipe.deq_table[idx.long()]), "%s: dequantised latents differ" % tag
    coded_bits = float(enc.nbits.double().mean()) / SYMS
    out = {"streams": B, "coded_bits_per_symbol": coded_bits, "parity": "round-trip identity on all %d streams" % B,
           "enc_bytes": int(enc.offsets[-1])}
    if rank == 0:
        streams, nbits_h, _, _ = enc.to_host()
        from oracle import oracle as orc
        # the coder against the oracle on 8 sampled streams, the quantiser on the same streams' latents
        picks, obits = oracle_parity(idx.cpu().numpy(), streams, nbits_h, n, 8)
        sub = lat_host[picks].numpy()
        want_idx = (orc.quantize_codebook(sub, pipe.codebook.cpu().numpy()) if pipe.quantizer == "codebook"
                    else orc.quantize_affine(sub, pipe.bits)[0].clip(0, n - 1))
        assert (idx[picks].cpu().numpy().astype("int64") == want_idx).all(), "%s: quantiser differs from the oracle" % tag
        out["parity"] += "; quantised indices and bitstreams of %d sampled streams bit-identical to the CPU oracle" % len(picks)
        out["oracle_coded_bits_per_symbol_sample"] = obits
        out["coded_bits_per_symbol_sample"] = float(sum(int(nbits_h[b]) for b in picks)) / (len(picks) * SYMS)
    del idx, enc, dec_idx, deq, dstatus

    # ---- device-resident metric
    for _ in range(max(warmup, 3)):
        step()
        flush.zero_()
    barrier()
    lib.lc_debug_launch_count(1)
    t_steps, t_dec = [], []
    for _ in range(steps):
        flush.zero_()  # L2 flush between timed iterations (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step((d0, d1))
        e1.record(stream)
        e1.synchronize()
        t_steps.append(e0.elapsed_time(e1))
        t_dec.append(d0.elapsed_time(d1))
    out["launches"] = int(lib.lc_debug_launch_count(1))
    barrier()
    out["ms_total"] = sum(t_steps)
    out["dec_ms"] = sum(t_dec) / len(t_dec)
    out["steps"] = steps

    # ---- end-to-end metric: host buffers in, host buffers out, copies inside the timed region
    if want_e2e:
        for _ in range(2):
            res = pipe.roundtrip_host(lat_host)
        barrier()
        e_times = []
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = pipe.roundtrip_host(lat_host)
            e_times.append(time.perf_counter() - t0)
        barrier()
        assert not res["enc_status"].numpy().any() and not res["dec_status"].numpy().any(), "%s: e2e coder fault" % tag
        want = pipe.deq_table.cpu()[pipe.quantize(lat).cpu().long()]
        assert torch.equal(res["deq"], want), "%s: e2e result differs" % tag
        out["e2e_s"] = sum(e_times)
        out["h2d"] = int(res["h2d_bytes"]); out["d2h"] = int(res["d2h_bytes"]); out["chunks"] = int(res["chunks"])
        # ---- the same trip as a STREAM of batches, two in flight (roundtrip_host_stream): every batch's latents are
        # copied up and its results copied down inside the timed region; only the overlap between batches differs
        for res in pipe.roundtrip_host_stream([lat_host] * 4):
            pass
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        bad = 0
        for res in pipe.roundtrip_host_stream([lat_host] * steps):
            bad += int(res["enc_status"].numpy().any()) + int(res["dec_status"].numpy().any())
        torch.cuda.synchronize()
        out["e2e_stream_s"] = time.perf_counter() - t0
        barrier()
        assert bad == 0, "%s: e2e (stream) coder fault" % tag
        assert torch.equal(res["deq"], want), "%s: e2e (stream) result differs" % tag
        out["h2d_stream"] = int(res["h2d_bytes"]); out["d2h_stream"] = int(res["d2h_bytes"])
    del lat
    return out


def hbm_kernels(dev, peak_gbs, flush):
    """K1 / K2 / dequantisers standalone at cfg4 size (65,536 latents = 2 GiB of fp32 in): HBM-bound kernels against
    the measured copy bandwidth.  Algorithmic bytes = fp32 read + index written (+ fp32 written)."""
    import torch
    from image_compression_2_b200 import codec
    Bk = 65536
    n_elem = Bk * SYMS
    out = {}
    lat = torch.empty(Bk, R, C, dtype=torch.float32, device=dev).uniform_(-0.5, 0.5)
    cb = torch.linspace(-1, 1, 256).float().to(dev)
    cases = [
        ("lc_quant_codebook_kernel<u8> (K2, 8 bit)", lambda: codec.quantize_codebook(lat, cb, sorted_ascending=True, idx_dtype=torch.uint8), 5),
        ("lc_quant_codebook_kernel<i32> (K2, reference's int32 indices)", lambda: codec.quantize_codebook(lat, cb, sorted_ascending=True), 8),
        ("lc_quant_affine_kernel<u8> idx only (K1, 8 bit)", lambda: codec.quantize_affine(lat, 8, want_wq=False, idx_dtype=torch.uint8), 5),
        ("lc_quant_affine_kernel fp32 dequantised only (K1, StyleGAN3Compressor.compress)", lambda: codec.quantize_affine(lat, 8, want_idx=False), 8),
        ("lc_quant_affine_kernel<u8> idx + fp32 (K1)", lambda: codec.quantize_affine(lat, 8, idx_dtype=torch.uint8), 9),
    ]
    idx8 = codec.quantize_codebook(lat, cb, sorted_ascending=True, idx_dtype=torch.uint8)[0]
    cases.append(("lc_dequant_codebook_kernel<u8> (dequantiser B)", lambda: codec.dequantize_codebook(idx8, cb), 5))
    cases.append(("lc_dequant_affine_kernel<u8> (dequantiser A)", lambda: codec.dequantize_affine(idx8, 8), 5))
    stream = torch.cuda.current_stream()
    for name, fn, bytes_per_sym in cases:
        for _ in range(3):
            fn()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        gbs = n_elem * bytes_per_sym / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "algorithmic_bytes": n_elem * bytes_per_sym, "bytes_per_symbol": bytes_per_sym,
                     "achieved_gbs": gbs, "frac_of_measured_hbm_peak": gbs / peak_gbs}
    return {"size": "%d latents 16x512 (2 GiB fp32 in, cfg4)" % Bk, "peak_gbs": peak_gbs, "kernels": out}


def run_b200(args):
    import torch
    import torch.distributed as dist

    from image_compression_2_b200 import LatentPipeline, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on STDOUT when the first communicator is created; the contract is one JSON
        # line there, so stdout points at stderr until the communicator exists
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    n = 1 << args.bits
    B = args.batch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    # ---- headline: every rank codes its shard [lo, hi) of the global batch of world*B streams; inputs are generated
    # on the CPU (seeded by the shard's first stream) so the oracle and the kernels see the same bits
    lo, hi = sharding.shard_range(world * B, rank, world)
    lat_host = synth(args.kind, hi - lo, 1000 + 2 * 100000 + lo).pin_memory()
    pipe = LatentPipeline(n_symbols=n, R=R, C=C, quantizer="codebook")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    head = run_workload(pipe, lat_host, dev, args.steps, args.warmup, flush, barrier, True, rank, "cfg2")
    clocks = sampler.stop() if rank == 0 else None
    ms_total, e_total, es_total = reduce_max([head["ms_total"], head["e2e_s"], head["e2e_stream_s"]])
    # optional size gather (not timed; the only collective this framework has)
    sizes, _ = sharding.gather_shard_bytes(head["enc_bytes"], device=dev)
    del lat_host

    # ---- the other BASELINE.json configs, same gate, same timing rules
    sweep = []
    if not args.no_sweep:
        cfg4_total, cfg5_total = 65536, 16384
        c4 = sharding.shard_range(cfg4_total, rank, world) if world > 1 else (0, 8192)
        c5 = sharding.shard_range(cfg5_total, rank, world)
        specs = [
            # (name, description, quantiser, bits, kind, [lo,hi) of this rank, seed base, global streams, e2e)
            ("cfg3_4bit", "cfg3: 4096 latents per GPU, quantization_bits=4 (affine quantiser A), 'wide' latents", "affine", 4,
             "wide", (rank * 4096, rank * 4096 + 4096), 3 * 100000 + 4000, world * 4096, False),
            ("cfg3_8bit", "cfg3: 4096 latents per GPU, quantization_bits=8 (affine quantiser A), 'enc_like' latents", "affine", 8,
             "enc_like", (rank * 4096, rank * 4096 + 4096), 3 * 100000 + 8000, world * 4096, False),
            ("cfg3_10bit", "cfg3: 4096 latents per GPU, quantization_bits=10 (affine quantiser A), 'enc_like' latents", "affine", 10,
             "enc_like", (rank * 4096, rank * 4096 + 4096), 3 * 100000 + 10000, world * 4096, False),
            ("cfg4", ("cfg4 as stated: 65,536 latents batch-sharded over %d GPUs (%d per GPU), 8-bit" % (world, c4[1] - c4[0]))
             if world > 1 else "cfg4 share of one of 8 GPUs: 8192 of the 65,536 latents, 8-bit (run with --gpus 2/4/8 for the config as stated)",
             "codebook", 8, "enc_like", c4, 4 * 100000, cfg4_total if world > 1 else 8192, True),
            ("cfg5", "cfg5: 16,384 hierarchical multi-scale latents (per-level sigma) over %d GPU(s), Gumbel-softmax codebook "
             "quantiser B (argmin, n=256)" % world, "codebook", 8, "hier", c5, 5 * 100000, cfg5_total, False),
        ]
        for name, desc, quant, bits, kind, (s_lo, s_hi), seed, global_streams, want_e2e in specs:
            local = [0.0, 0.0, 0.0, 0.0]  # device ms, e2e s, failed, launches
            info = {}
            # each rank measures locally inside try/except and the ranks then meet in ONE all-reduce, so a failure here
            # (e.g. memory on a shared box) cannot hang the job or cost the headline line above
            try:
                p2 = LatentPipeline(n_symbols=1 << bits, R=R, C=C, quantizer=quant)
                lat2 = synth(kind, s_hi - s_lo, 1000 + seed + s_lo)
                if want_e2e:
                    lat2 = lat2.pin_memory()
                r = run_workload(p2, lat2, dev, args.sweep_steps, args.warmup, flush, barrier, want_e2e, rank, name)
                local = [r["ms_total"], r.get("e2e_s", 0.0), 0.0, float(r["launches"])]
                info = r
                del lat2, p2
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001 -- reported, never fatal
                local[2] = 1.0
                info = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            t_ms, te_s, failed, launches = reduce_max(local)
            entry = {"name": name, "workload": desc, "streams_total": global_streams, "streams_per_gpu": s_hi - s_lo,
                     "n_symbols": 1 << bits, "quantizer": quant, "latents": kind}
            if failed:
                entry["failed"] = info.get("error", "on another rank")
            else:
                steps2 = args.sweep_steps
                v = global_streams * SYMS * steps2 / (t_ms * 1e-3)
                entry.update({"value": v, "unit": UNIT, "streams_per_s": v / SYMS, "steps": steps2,
                              "ms_per_step": t_ms / steps2, "decode_ms": info.get("dec_ms"),
                              "coded_bits_per_symbol": info.get("coded_bits_per_symbol"),
                              "coded_bits_per_symbol_sample": info.get("coded_bits_per_symbol_sample"),
                              "oracle_coded_bits_per_symbol_sample": info.get("oracle_coded_bits_per_symbol_sample"),
                              "parity": info.get("parity"), "gpu_launches": int(launches)})
                if want_e2e:
                    ve = global_streams * SYMS * steps2 / te_s
                    entry["e2e"] = {"value": ve, "unit": UNIT, "streams_per_s": ve / SYMS, "ms_per_step": 1e3 * te_s / steps2,
                                    "h2d_bytes_per_step": info.get("h2d"), "d2h_bytes_per_step": info.get("d2h"),
                                    "host_chunks": info.get("chunks")}
            sweep.append(entry)

    hbm = None
    if not args.no_sweep and rank == 0 and world == 1:
        try:
            hbm = hbm_kernels(dev, peak, flush)
        except Exception as e:  # noqa: BLE001
            hbm = {"failed": "%s: %s" % (type(e).__name__, str(e)[:300])}

    if rank == 0:
        total_syms = world * B * SYMS
        value = total_syms * args.steps / (ms_total * 1e-3)
        e2e_serial = total_syms * args.steps / e_total
        e2e_value = total_syms * args.steps / es_total
        w8 = n == 256 and R == 16 and C == 512
        kname = ("lc_decode_v2_w8_thr_kernel" if B > 8 * 148 else "lc_decode_v2_w8_kernel") if w8 else "lc_decode_v2_kernel"
        facts_all, facts_file = load_kernel_facts()
        facts = facts_all.get(kname, {})
        same_cfg = facts.get("streams") == B and facts.get("n_symbols") == n
        facts_current = facts_all.get("_kernel_source_sha16") == kernel_source_hash()
        dec_ms = head["dec_ms"]
        coded_bits = head["coded_bits_per_symbol"]
        # Algorithmic HBM bytes of the decode launch per symbol: coded bits/8 read + index written + 4 B fp32 written
        idx_bytes = 1 if n <= 256 else 2
        bytes_per_launch = B * SYMS * (coded_bits / 8.0 + idx_bytes + 4.0)
        hbm_achieved = bytes_per_launch / (dec_ms * 1e-3) / 1e9
        # ... but the decoder is a dependent chain per stream: its roofline is the SM issue rate
        slots_peak = 148 * 4 * SM_CLOCK_HZ
        ips = facts.get("warp_inst_per_symbol") if same_cfg else None
        inst_rate = (ips * B * SYMS / (dec_ms * 1e-3)) if ips else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "streams_per_gpu": B, "symbols_per_stream": SYMS,
                       "n_symbols": n, "coder_mode": "repaired", "coded_bits_per_symbol": coded_bits,
                       "oracle_coded_bits_per_symbol_sample": head.get("oracle_coded_bits_per_symbol_sample"),
                       "coded_bits_per_symbol_sample": head.get("coded_bits_per_symbol_sample"),
                       "index_dtype": "uint8" if n <= 256 else "uint16",
                       "l2": "flushed between timed steps (256 MiB memset, untimed)", "parallelism": "streams sharded by image, no collective",
                       "parity": head["parity"], "streams_per_s": value / SYMS},
            # host buffers in, host buffers out, every step's copies inside the timed region.  `value`: the steps as a
            # stream of batches with two in flight (LatentPipeline.roundtrip_host_stream: the host->device copy of one
            # batch runs under the kernels of the one before; K steps, wall clock between two synchronisations).
            # `one_batch_at_a_time`: the same trip with a synchronise after every batch (roundtrip_host), L2 flushed
            # between steps.
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": head["h2d_stream"],
                    "d2h_bytes_per_step": head["d2h_stream"], "ms_per_step": 1e3 * es_total / args.steps,
                    "mode": "stream of batches, 2 in flight (roundtrip_host_stream)",
                    "frac_of_device_value": e2e_value / value,
                    "one_batch_at_a_time": {"value": e2e_serial, "ms_per_step": 1e3 * e_total / args.steps,
                                            "h2d_bytes_per_step": head["h2d"], "d2h_bytes_per_step": head["d2h"],
                                            "frac_of_device_value": e2e_serial / value}},
            # counted by the library (lc_debug_launch_count) over the timed steps of the headline workload
            "gpu_launches": head["launches"],
            "roofline": {"kernel": kname, "bound": "issue_slots",
                         "achieved": inst_rate / 1e9 if inst_rate else None, "peak": slots_peak / 1e9,
                         "unit": "G warp-instructions/s", "frac": inst_rate / slots_peak if inst_rate else None,
                         "peak_source": "148 SMs x 4 schedulers x %.3f GHz (clocks.max.sm)" % (SM_CLOCK_HZ / 1e9),
                         "warp_inst_per_symbol": ips, "kernel_ms": dec_ms,
                         "kernel_share_of_step": dec_ms / (ms_total / args.steps),
                         "cycles_per_symbol_per_stream": dec_ms * 1e-3 * SM_CLOCK_HZ / SYMS if B <= 8 * 148 else None,
                         "traffic": facts.get("dram_bytes_per_launch") if same_cfg else None,
                         "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": peak, "unit": "GB/s",
                                 "frac": hbm_achieved / peak, "algorithmic_bytes_per_launch": bytes_per_launch,
                                 "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback"},
                         "ncu": facts.get("_source", facts_file), "ncu_facts_match_these_kernel_sources": facts_current,
                         "note": "decode = one dependent chain per stream: latency/issue-bound, not HBM-bound (DESIGN.md "
                                 "section 5); instructions per symbol and DRAM traffic come from the committed ncu capture "
                                 "(they cannot be measured outside a profiler), kernel_ms is measured live with CUDA events"},
            "clocks": clocks,
            "rank_bytes": [int(x) for x in sizes.tolist()],
        }
        if sweep:
            line["sweep"] = sweep
        if hbm is not None:
            line["hbm_kernels"] = hbm
        if not args.no_cpu_baseline:
            codes, threads = cpu_sample(args, n)
            cv, cdt, cs = cpu_roundtrip_rate(codes, n, threads, repeats=cpu_repeats(codes.shape[0], n, threads))
            line["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d streams of the same workload in %.1f s, C port of the reference coder "
                                              "(oracle/latent_oracle.c), %s" % (cs, cdt, cpu_model())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)
    finally:
        _close_pool()


if __name__ == "__main__":
    main()
