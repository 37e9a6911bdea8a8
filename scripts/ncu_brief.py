#!/usr/bin/env python
"""Brief counters + stall-sample shares + the hottest SASS lines of an ncu capture."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
want = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum smsp__inst_executed.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed smsp__issue_active.avg.pct_of_peak_sustained_active
sm__warps_active.avg.pct_of_peak_sustained_active launch__registers_per_thread launch__waves_per_multiprocessor
dram__throughput.avg.pct_of_peak_sustained_elapsed lts__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum lts__t_sector_hit_rate.pct""".split()
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"{w} = {vals[i]} {units[i]}")
tot = 0; st = {}
for i, h in enumerate(hdr):
    if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
        st[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(vals[i]); tot += float(vals[i])
print("stall samples:", ", ".join(f"{k} {v / tot * 100:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v / tot > 0.02))
if "--sass" in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = None
    for r in rows:
        if "Source" in r and "# Samples" in r:
            h = r; break
    if h:
        si, ci, ei = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        body = [r for r in rows[rows.index(h) + 1:] if len(r) == len(h)]
        tot_s = sum(int(r[ci] or 0) for r in body)
        print("instructions in kernel:", len(body), " samples:", tot_s)
        for r in sorted(body, key=lambda r: -int(r[ci] or 0))[:25]:
            print(f"{int(r[ci] or 0) / max(tot_s, 1) * 100:5.1f}%  exec {r[ei]:>10}  {r[si][:100]}")
