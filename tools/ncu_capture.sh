#!/bin/bash
# On the GPU box: ncu captures of this round, reduced to text on the box (a full report of every kernel is ~200 MB,
# gpurun brings back 64 MB).  Leaves in gpurun_out/: r02_ncu_all_kernels.md (+ raw csv), the decoder-only report with
# source (small), the launch list of `bench.py --steps 2`.
set -u
mkdir -p gpurun_out
python tools/ncu_all_kernels.py > gpurun_out/plain_all.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_all.log; exit 1; }
ncu --set full --clock-control none --profile-from-start off -o /tmp/r02_all python tools/ncu_all_kernels.py > gpurun_out/ncu_all.log 2>&1
echo "ncu all rc=$?"
python tools_ncu_summary.py /tmp/r02_all.ncu-rep gpurun_out/r02_ncu_all_kernels.md > gpurun_out/summary.log 2>&1
ncu -i /tmp/r02_all.ncu-rep --page raw --csv > gpurun_out/r02_all_raw.csv 2>/dev/null
python tools/make_kernel_facts.py gpurun_out/r02_all_raw.csv gpurun_out/r02_kernel_facts.json "profiles/r02_ncu_all_kernels.md (ncu --set full, B200, tools/ncu_all_kernels.py: config 2 = 1024 streams x 8192 symbols at 8 bits; throughput decoder at 8192 streams)" > gpurun_out/facts.log 2>&1
python bench.py --steps 1 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/plain_dec.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lc_decode_v2_w8_kernel -s 3 -c 1 -o gpurun_out/r02_dec python bench.py --steps 1 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/ncu_dec.log 2>&1
echo "ncu dec rc=$?"
python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/plain_launches.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out | head -30
du -sh gpurun_out
