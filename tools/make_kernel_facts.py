#!/usr/bin/env python
"""profiles/<round>_kernel_facts.json from the raw csv of an all-kernels ncu capture (tools/ncu_capture.sh):
the numbers bench.py quotes but cannot measure outside a profiler -- DRAM bytes per launch, executed warp
instructions per symbol, issue-slot utilisation -- for the kernels of config 2 (first occurrence of each kernel in
tools/ncu_all_kernels.py = its config-2 section, 1024 streams) and the throughput decoder at 8192 streams.
Stamped with the hash of the kernel sources the capture was taken from (bench.kernel_source_hash).

  python tools/make_kernel_facts.py gpurun_out/r02_all_raw.csv profiles/r02_kernel_facts.json "<source description>"
"""
import csv, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

raw, out, src = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(raw)))
h = rows[0]
col = {k: i for i, k in enumerate(h)}
WANT = {"lc_quant_codebook_uniform_kernel": 1024, "lc_quant_codebook_kernel": 1024, "lc_v2_tables_kernel": 1024,
        "lc_t2_kernel": 1024, "lc_enc_sort_kernel": 1024, "lc_enc_sort2_kernel": 1024, "lc_enc_phase_a_sparse_kernel": 1024,
        "lc_enc_phase_b1_kernel": 1024, "lc_enc_phase_b2_kernel": 1024, "lc_scan_sizes_kernel": 1024,
        "lc_compact_kernel": 1024, "lc_decode_v2_w8_kernel": 1024, "lc_decode_v2_w8_thr_kernel": 8192}


def num(r, k):
    v = r[col[k]].replace(",", "")
    return float(v) if v not in ("", "n/a") else 0.0


def to_bytes(r, k):
    u = rows[1][col[k]].lower()
    return num(r, k) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


def to_ms(r, k):
    u = rows[1][col[k]].lower()
    return num(r, k) * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3}.get(u, 1)


facts = {"_source": src, "_kernel_source_sha16": bench.kernel_source_hash()}
for r in rows[2:]:
    if len(r) != len(h):
        continue
    m = re.search(r'(lc_[a-z0-9_]+)', r[col["Kernel Name"]])
    if not m or m.group(1) not in WANT or m.group(1) in facts:
        continue
    name = m.group(1)
    streams = WANT[name]
    inst = num(r, "smsp__inst_executed.sum")
    cyc = num(r, "sm__cycles_elapsed.avg")
    facts[name] = {
        "streams": streams, "n_symbols": 256,
        "dram_bytes_per_launch": int(to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")),
        "warp_inst_per_symbol": round(inst / (streams * 8192.0), 2),
        "issue_slot_utilisation": round(inst / (cyc * 592.0), 3) if cyc else None,
        "launch_ms_under_ncu": round(to_ms(r, "gpu__time_duration.sum"), 4),
        "registers_per_thread": int(num(r, "launch__registers_per_thread")),
        "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
    }
json.dump(facts, open(out, "w"), indent=1)
print(json.dumps(facts, indent=1))
