"""Extract the judged metrics from an ncu report into a small text file for profiles/ (run where ncu is installed; no GPU).

    python tools/ncu_summary.py gpurun_out/prof_fwd16.ncu-rep profiles/r2_lstm16_fwd_ncu.txt
"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__thread_inst_executed_per_inst_executed.ratio",
    # the L1 / shared-memory data pipe: tensor-core operand reads and LSU traffic share it
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines, blob = [], {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines.append(f"kernel: {name}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"  {w:75s} {r[i]:>16s} {units[i]}")
                blob[w] = r[i]
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        sh = srows[1]
        stalls = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
        tot = {h: sum(int(x[sh.index(h)] or 0) for x in srows[2:] if len(x) == len(sh)) for h in stalls}
        t = sum(tot.values()) or 1
        lines.append("  warp-stall sampling (all samples): " +
                     ", ".join(f"{k[6:]} {100 * v / t:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
        break
    lines.insert(0, f"# extracted from {rep} with tools/ncu_summary.py (ncu --set full --clock-control none --import-source on, one launch)")
    open(out, "w").write("\n".join(lines) + "\n")
    if out.endswith(".txt"):
        json.dump(blob, open(out[:-4] + ".json", "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
