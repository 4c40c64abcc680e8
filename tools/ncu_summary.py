"""Summarise an .ncu-rep: headline metrics, per-opcode executed instructions, stall mix.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [units_per_launch]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit, val = rows[0], rows[1], rows[-1]
m = dict(zip(hdr, val))
u = dict(zip(hdr, unit))
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in want:
    if k in m:
        print(f"{k:70s} {m[k]} {u.get(k,'')}")
print("-- stalls (warps per issue-active cycle)")
st = {k: float(v) for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v not in ("", "no data")}
for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
    print(f"   {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
iS, iE, iP = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
tot, samp = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= iE:
        continue
    t = r[iS].strip().split()
    if not t:
        continue
    o = t[1] if t[0].startswith("@") else t[0]
    o = o.rstrip(";").split(".")[0]
    tot[o] += int(r[iE]); samp[o] += int(r[iP])
T, S = sum(tot.values()), max(sum(samp.values()), 1)
print(f"-- executed warp instructions: {T}" + (f" = {T*32/units:.1f} thread-instr per unit" if units else ""))
for k, v in tot.most_common(22):
    per = f"{v*32/units:7.2f}/unit" if units else f"{v:12d}"
    print(f"   {k:10s} {per}  samples {100*samp[k]/S:5.1f}%")
