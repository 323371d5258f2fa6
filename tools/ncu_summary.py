"""Summarise an .ncu-rep (raw + source pages) for one kernel: key metrics, opcode mix, stall mix.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [--segments] [--index K]   (K = which captured launch, default 0)"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
K = int(sys.argv[sys.argv.index("--index") + 1]) if "--index" in sys.argv else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + K]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h:72s} {v:>22s} {u}")
print("-- stall reasons (per issue active) --")
st = [(float(v), h) for h, v in zip(hdr, vals) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for v, h in sorted(st, reverse=True)[:8]:
    print(f"   {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:6.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
per = max(1, (len(starts) - 1) // max(1, len(list(csv.reader(io.StringIO(raw)))) - 2))   # sections per launch (SASS, source)
rows = rows[starts[per * K]:starts[per * K + 1]]          # the SASS section of launch K
hdr, data = rows[1], [r for r in rows[2:] if len(r) > 5]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iex]) for r in data)
print(f"-- opcode mix: {tot} warp-instr, {len(data)} SASS lines --")
ops, smp = collections.Counter(), collections.Counter()
for r in data:
    t = r[isrc].strip().split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += int(r[iex]); smp[op] += int(r[ismp])
for op, c in ops.most_common(14):
    print(f"   {op:10s} {c:>12d} {100 * c / tot:5.1f}%   samples {smp[op]}")
if "--segments" in sys.argv:
    acc = 0; k = 0; sacc = 0
    for r in data:
        acc += int(r[iex]); sacc += int(r[ismp])
        if "BAR." in r[isrc] or "SYNCS" in r[isrc]:
            print(f"   seg {k:3d}: {acc:>10d} instr {sacc:>6d} samples  ends at {r[isrc].strip()[:60]}")
            acc = 0; sacc = 0; k += 1
    print(f"   tail   : {acc:>10d} instr {sacc:>6d} samples")
