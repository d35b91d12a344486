"""Summarise the SASS page of an ncu report: samples / instructions per execution-count bucket (loops show up as buckets),
the top-sampled instructions, and the headline metrics.  usage: ncu_sass_buckets.py report.ncu-rep"""
import csv, collections, math, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[-1]
for name in ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
             "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
             "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum"]:
    if name in h: print(name, v[h.index(name)])
for i, n in enumerate(h):
    if "pcsamp_warps_issue_stalled" in n and "not_issued" not in n and float(v[i] or 0) > 0: print("  ", n.split("stalled_")[1], v[i])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp] or 0) for r in data); totex = sum(int(r[iex] or 0) for r in data)
bk, bs, be = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    e = int(r[iex] or 0)
    b = 0 if e == 0 else round(math.log10(e), 1)
    bk[b] += 1; bs[b] += int(r[isamp] or 0); be[b] += e
print("log10(exec)  #sass  samples%  instr%")
for b in sorted(bk):
    if bs[b] / tot > 0.002: print(f"{b:6.1f} {bk[b]:6d} {100*bs[b]/tot:8.2f} {100*be[b]/totex:8.2f}")
print("top-sampled instructions:")
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:25]):
    r = data[i]; print(f"{i:5d} {r[iex]:>11s} {100*int(r[isamp])/tot:6.2f}%  {r[isrc][:80]}")
