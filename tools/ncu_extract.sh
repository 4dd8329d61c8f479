#!/bin/bash
# usage: ncu_extract.sh <target> <kernel-regex> <count>   (runs on the GPU box; writes small CSV/txt files only)
t=$1; k=$2; c=$3
rep=/tmp/ncu_$t
timeout 500 ncu --set full --clock-control none --import-source on -k regex:$k -c $c -o $rep python tools/ncu_targets.py $t > /tmp/ncu_$t.log 2>&1
tail -1 /tmp/ncu_$t.log
ncu -i $rep.ncu-rep --page raw --csv > /tmp/raw_$t.csv 2>/dev/null
python - "$t" <<'PY'
import csv, sys
t = sys.argv[1]
rows = list(csv.reader(open(f"/tmp/raw_{t}.csv")))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_read.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__memory_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex.sum", "smsp__cycles_elapsed.avg.per_second", "lts__cycles_elapsed.avg.per_second"]
with open(f"gpurun_out/ncu_{t}_summary.txt", "w") as f:
    names = [h for h in hdr]
    for d in data:
        f.write("=" * 100 + "\n")
        for i, h in enumerate(names):
            if h in keep or "pipe" in h and "pct_of_peak_sustained_active" in h or h.startswith("smsp__average_warp") or h.startswith("smsp__average_warps_issue_stalled") or "warp_issue_stalled" in h and "ratio" in h:
                f.write(f"{h} [{units[i]}] = {d[i]}\n")
PY
ncu -i $rep.ncu-rep --page source --csv > /tmp/src_$t.csv 2>/dev/null
python - "$t" <<'PY'
import csv, sys
t = sys.argv[1]
try:
    rows = list(csv.reader(open(f"/tmp/src_{t}.csv")))
except Exception as e:
    print("no source page", e); sys.exit(0)
# keep the 60 hottest SASS lines by "Warp Stall Sampling (All Samples)" if present
hdr = None
out = []
for r in rows:
    if hdr is None and ("Source" in r or "# " in r[0:1] or any("Sampling" in c for c in r)):
        hdr = r; continue
    if hdr: out.append(r)
if hdr:
    col = next((i for i, h in enumerate(hdr) if "Sampling" in h and "All" in h), None)
    if col is not None:
        def val(r):
            try: return float(r[col])
            except: return 0.0
        out.sort(key=val, reverse=True)
    with open(f"gpurun_out/ncu_{t}_hot_sass.csv", "w") as f:
        w = csv.writer(f); w.writerow(hdr)
        for r in out[:70]: w.writerow(r)
PY
ls -la gpurun_out/ncu_${t}_*
