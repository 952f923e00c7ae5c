#!/usr/bin/env python
"""bench.py -- latent symbols/s, encode + decode, bit-exact (BASELINE.json metric).

A step = one pass of the whole hot path over one batch of synthetic W+ latents:
  quantise (codebook argmin) -> encode -> compact -> decode (+ fused dequantise).
Default workload = BASELINE.json configs[1]: 1024 latents of 16x512 at 8 bits per GPU
("enc_like" synthetic latents).  With N GPUs every rank codes its own 1024
streams (weak scaling, no collective on the hot path; the only NCCL traffic is the optional size
gather after the timed region).

  python bench.py [--gpus N --steps K --warmup W]          this framework
  python bench.py --impl reference [...]                    the CPU coder on the host cores

One JSON line on stdout (rank 0).
"""
import argparse
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="streams per GPU")
    ap.add_argument("--bits", type=int, default=8)
    ap.add_argument("--kind", default="enc_like")
    ap.add_argument("--cpu-sample-streams", type=int, default=0, help="0 = auto (about 10-30 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--large-batch", type=int, default=8192,
                    help="streams per GPU of the supplementary throughput measurement (0 = skip); 8192 is config 4's "
                         "share per GPU (65,536 latents over 8 GPUs)")
    return ap.parse_args()


def synth(kind, B, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    if kind == "enc_like":
        return torch.randn(B, R, C, generator=g) * 0.14
    if kind == "wide":
        return torch.randn(B, R, C, generator=g) * 0.4
    if kind == "uniform":
        return torch.rand(B, R, C, generator=g) * 2 - 1
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

    # every rank codes its own shard of the global batch (world*B streams), inputs generated on the
    # CPU so the oracle and the kernels see the same bits
    lo, hi = sharding.shard_range(world * B, rank, world)
    lat_host = synth(args.kind, B, 1000 + 2 * 100000 + rank).pin_memory()
    lat = lat_host.to(dev)
    pipe = LatentPipeline(n_symbols=n, R=R, C=C, quantizer="codebook")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    stream = torch.cuda.current_stream()

    def step_device(events=None):
        idx = pipe.quantize(lat)
        enc = pipe.encode(idx)
        if events is not None:
            events[0].record(stream)
        dec_idx, deq, dstatus, dfault = pipe.decode(enc.data, enc.offsets, enc.nbits, B)
        if events is not None:
            events[1].record(stream)
        return idx, enc, dec_idx, deq, dstatus

    # ---- parity gate on this rank's actual workload before any timing
    idx, enc, dec_idx, deq, dstatus = step_device()
    torch.cuda.synchronize()
    assert int(enc.status.abs().sum()) == 0 and int(dstatus.abs().sum()) == 0, "coder fault in the bench workload"
    assert torch.equal(dec_idx.view(B, R, C), idx), "round trip is not the identity"
    assert torch.equal(deq.view(B, R, C), pipe.codebook[idx.long()]), "dequantised latents differ"
    coded_bits = float(enc.nbits.double().mean())
    parity_note = "round-trip identity on all streams"
    if rank == 0:
        from oracle import oracle as orc
        streams, nbits_h, _, _ = enc.to_host()
        idx_h = idx.cpu().numpy()
        for b in range(0, B, max(1, B // 8)):
            ref = orc.encode_stream(idx_h[b:b + 1], n, "repaired")
            assert nbits_h[b] == ref["nbits"] and streams[b] == ref["packed"], "bitstream differs from the oracle"
        parity_note += "; %d sampled bitstreams bit-identical to the CPU oracle" % len(range(0, B, max(1, B // 8)))

    # ---- device-resident metric
    for _ in range(max(args.warmup, 3)):
        step_device()
        flush.zero_()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t_steps, t_dec = [], []
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_device((d0, d1))
        e1.record(stream)
        e1.synchronize()
        t_steps.append(e0.elapsed_time(e1))
        t_dec.append(d0.elapsed_time(d1))
    barrier()
    t_total = torch.tensor([sum(t_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
    ms_total = float(t_total)

    # ---- end-to-end metric: host buffers in, host buffers out, copies inside the timed region
    for _ in range(2):
        res = pipe.roundtrip_host(lat_host)
    barrier()
    e_times = []
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = pipe.roundtrip_host(lat_host)
        e_times.append(time.perf_counter() - t0)
    barrier()
    assert not res["enc_status"].numpy().any() and not res["dec_status"].numpy().any()
    assert torch.equal(res["deq"], pipe.codebook.cpu()[idx.cpu().long()]), "e2e result differs"
    e_total = torch.tensor([sum(e_times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_total, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None

    # optional size gather (not timed; the only collective this framework has)
    sizes, _ = sharding.gather_shard_bytes(int(enc.offsets[-1]), device=dev)

    # ---- supplementary: the same round trip at config 4's per-GPU share (several waves of blocks per launch instead
    # of one), device-resident like `value`; reported beside the headline, never instead of it
    large = None
    if args.large_batch > B and args.kind == "enc_like" and args.bits == 8:
        B2, steps2 = args.large_batch, 3
        # Each rank measures locally inside try/except and the ranks then meet in ONE all-reduce, so a failure here
        # (e.g. memory on a shared box) cannot hang the job or cost the headline line above.
        local = [0.0, 0.0, 0.0]  # device ms, e2e s, failed
        info = {}
        try:
            lat2 = synth(args.kind, B2, 1000 + 4 * 100000 + rank).to(dev)

            def step_large():
                idx2 = pipe.quantize(lat2)
                enc2 = pipe.encode(idx2)
                return idx2, enc2, pipe.decode(enc2.data, enc2.offsets, enc2.nbits, B2)

            idx2, enc2, (dec2, _, dst2, _) = step_large()
            torch.cuda.synchronize()
            assert int(enc2.status.abs().sum()) == 0 and int(dst2.abs().sum()) == 0 and torch.equal(dec2.view(B2, R, C), idx2)
            del idx2, enc2, dec2, dst2
            for _ in range(steps2):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                step_large()
                e1.record(stream)
                e1.synchronize()
                local[0] += e0.elapsed_time(e1)
            # ... and end to end with host buffers (chunked over four streams, copies under the kernels)
            lat2_host = lat2.cpu().pin_memory()
            for _ in range(2):
                pipe.roundtrip_host(lat2_host)
            for _ in range(steps2):
                flush.zero_()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                res2 = pipe.roundtrip_host(lat2_host)
                local[1] += time.perf_counter() - t0
            assert not res2["enc_status"].numpy().any() and not res2["dec_status"].numpy().any()
            assert torch.equal(res2["deq"], pipe.codebook.cpu()[pipe.quantize(lat2).cpu().long()]), "e2e result differs"
            info = {"h2d_bytes_per_step": int(res2["h2d_bytes"]), "d2h_bytes_per_step": int(res2["d2h_bytes"]),
                    "host_chunks": int(res2["chunks"])}
        except Exception as e:  # noqa: BLE001 -- reported, never fatal
            local[2] = 1.0
            info = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
        agg = torch.tensor(local, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        t2_ms, te_s, failed = (float(x) for x in agg)
        if failed:
            large = {"streams_per_gpu": B2, "failed": info.get("error", "on another rank")}
        else:
            v2 = world * B2 * SYMS * steps2 / (t2_ms * 1e-3)
            v2e = world * B2 * SYMS * steps2 / te_s
            large = {"workload": "cfg4 share: %d synthetic W+ latents per GPU (65,536 over 8 GPUs), 8-bit round trip" % B2,
                     "streams_per_gpu": B2, "value": v2, "unit": UNIT, "streams_per_s": v2 / SYMS, "steps": steps2,
                     "ms_per_step": t2_ms / steps2, "parity": "round-trip identity on all streams",
                     "e2e": dict({"value": v2e, "unit": UNIT, "ms_per_step": 1e3 * te_s / steps2}, **info)}

    if rank == 0:
        total_syms = world * B * SYMS
        value = total_syms * args.steps / (ms_total * 1e-3)
        e2e_value = total_syms * args.steps / float(e_total)
        # roofline of the dominant kernel (the decoder).  Algorithmic bytes per symbol of that
        # launch: coded bits/8 read + 4 B int32 index + 4 B fp32 dequantised value written
        # (DESIGN.md section 5).  It is NOT an HBM-bound kernel; the fraction is reported as is.
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        dec_ms = sum(t_dec) / len(t_dec)
        facts = {}
        try:
            facts = json.load(open(os.path.join(ROOT, "profiles", "r01_kernel_facts.json"))).get("lc_decode_v2_w8_kernel" if (n == 256 and R == 16 and C == 512) else "lc_decode_v2_kernel", {})
        except Exception:
            pass
        if B > 8 * 148 and n == 256 and R == 16 and C == 512:  # the host picks the throughput build there
            facts = json.load(open(os.path.join(ROOT, "profiles", "r01_kernel_facts.json"))).get("lc_decode_v2_w8_thr_kernel", {})
        same_cfg = facts.get("streams") == B and facts.get("n_symbols") == n
        bytes_per_launch = B * SYMS * (coded_bits / SYMS / 8.0 + 8.0)
        achieved = bytes_per_launch / (dec_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "streams_per_gpu": B, "symbols_per_stream": SYMS,
                       "n_symbols": n, "coder_mode": "repaired", "coded_bits_per_symbol": coded_bits / SYMS,
                       "l2": "flushed between timed steps (256 MiB memset, untimed)", "parallelism": "streams sharded by image, no collective",
                       "parity": parity_note, "streams_per_s": value / SYMS},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(res["h2d_bytes"]),
                    "d2h_bytes_per_step": int(res["d2h_bytes"]), "ms_per_step": 1e3 * float(e_total) / args.steps},
            # per step: quantise, tables, two-visit table, sort, phase A, phase B1, B2, size scan, compaction,
            # tables, decode, redo pass
            "gpu_launches": 12 * args.steps,
            "roofline": {"kernel": ("lc_decode_v2_w8_thr_kernel" if B > 8 * 148 else "lc_decode_v2_w8_kernel") if (n == 256 and R == 16 and C == 512) else "lc_decode_v2_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": facts.get("dram_bytes_per_launch") if same_cfg else None,
                         "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback",
                         "algorithmic_bytes_per_launch": bytes_per_launch,
                         "kernel_ms": dec_ms, "kernel_share_of_step": dec_ms / (ms_total / args.steps),
                         "warp_inst_per_symbol": facts.get("warp_inst_per_symbol") if same_cfg else None,
                         "issue_slot_utilisation": facts.get("issue_slot_utilisation") if same_cfg else None,
                         "ncu": facts.get("_source", "profiles/r01_ncu_all_kernels_v11.md"),
                         "note": "decode = dependent chain per stream: latency/issue-bound, not HBM-bound "
                                 "(DESIGN.md section 5); traffic above the algorithmic bytes is the per-stream "
                                 "context words + 64-byte records"},
            "clocks": clocks,
            "rank_bytes": [int(x) for x in sizes.tolist()],
        }
        if large is not None:
            line["large_batch"] = large
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
