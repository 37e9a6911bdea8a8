#!/usr/bin/env python
"""Text summary of an `ncu --set full --import-source on` capture for profiles/:
the raw metrics that the DESIGN.md claims rest on + the SASS lines with the most stall samples.

    python scripts/summarize_ncu.py gpurun_out/r01_fwd.ncu-rep > profiles/r01_warp_persist_ncu_raw.txt
"""
import csv
import io
import subprocess
import sys

KEYS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
dram__throughput.avg.pct_of_peak_sustained_elapsed lts__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__throughput.avg.pct_of_peak_sustained_elapsed sm__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum
lts__t_sectors_srcunit_tex_op_red.sum lts__t_sectors_srcunit_tex_op_red.avg.pct_of_peak_sustained_elapsed
lts__t_sector_hit_rate.pct l1tex__t_sector_hit_rate.pct
launch__registers_per_thread launch__grid_size launch__block_size launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem launch__waves_per_multiprocessor sm__warps_active.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum sm__cycles_elapsed.max""".split()


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, val = rows[0], rows[1], rows[-1]
    print(f"# {rep}: {val[hdr.index('Kernel Name')][:150]}")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k} {units[i]} {val[i]}")
    for i, k in enumerate(hdr):
        if "issue_stalled" in k and k.endswith("per_warp_active.pct"):
            try:
                if float(val[i]) >= 2.0:
                    print(f"{k} % {val[i]}")
            except ValueError:
                pass
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    if len(src) > 2:
        h, data = src[1], src[2:]
        iS, iI, iT = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
        iW = h.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in h else None
        iWi = h.index("L1 Wavefronts Shared Ideal") if "L1 Wavefronts Shared Ideal" in h else None
        tot = sum(int(r[iS] or 0) for r in data)
        print(f"total samples {tot}  instructions executed {sum(int(r[iI] or 0) for r in data)}")
        top = sorted(range(len(data)), key=lambda i: -int(data[i][iS] or 0))[:25]
        for i in sorted(top):
            r = data[i]
            w = f" smem_wavefronts {r[iW]} ideal {r[iWi]}" if iW is not None and (r[iW] or "0") != "0" else ""
            print(f"{i:5d} samples {r[iS]:>6s} exec {r[iI]:>9s}{w}  {r[iT].strip()[:80]}")


if __name__ == "__main__":
    main()
