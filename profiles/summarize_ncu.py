"""Summarises one `ncu --set full` capture (read here, on the CPU box) into a small text file under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_ncu_tile_kernel.txt [traffic.json]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = [f"source: {rep}", f"kernel launches captured: {len(data)}", ""]
    name_i = hdr.index("Kernel Name")
    lines.append("kernel: " + data[0][name_i])
    vals = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            v = [r[i] for r in data]
            vals[k] = v
            lines.append(f"{k:75s} [{units[i]}] " + "  ".join(v))
    lines.append("")
    lines.append("warp stall reasons (average warps stalled per issue-active cycle, first launch):")
    for i, h in enumerate(hdr):
        if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
            v = float(data[0][i])
            if v > 0.03:
                lines.append(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):22s} {v:.3f}")
    open(out, "w").write("\n".join(lines) + "\n")
    if len(sys.argv) > 3:
        def mb(x):
            return float(x.replace(",", "")) * 1e6
        rd = sum(mb(v) for v in vals["dram__bytes_read.sum"]) / len(data)
        wr = sum(mb(v) for v in vals["dram__bytes_write.sum"]) / len(data)
        json.dump({"dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr, "source": rep,
                   "kernel": data[0][name_i].split("<")[0].replace("void ", "")}, open(sys.argv[3], "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
