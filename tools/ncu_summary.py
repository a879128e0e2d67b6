"""Key counters of every launch in an `ncu --set full` report, as kept under profiles/:
    python tools/ncu_summary.py report.ncu-rep "header line" > profiles/rNN_ncu_....txt
With --regions the warp-stall samples of the first launch are also aggregated per 64-instruction
region of the SASS (what the inner loops of a kernel cost and why they stall)."""
import collections, csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]


def page(rep, which, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    print(sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else rep)
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    for n, vals in enumerate(rows[2:]):
        print(f"--- launch {n}" + (f": {vals[name_col]}" if name_col is not None else "") + " ---")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:<90s} {vals[i]:>16s} {units[i]}")
    if "--regions" in sys.argv:
        rows = page(rep, "source", ("--print-source", "sass"))
        start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
        hdr = rows[start]
        ix = {h: i for i, h in enumerate(hdr)}
        data = []
        for r in rows[start + 1:]:
            if r and r[0] == "Kernel Name":
                break           # first launch only
            if len(r) == len(hdr):
                data.append(r)
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        print(f"--- warp-stall samples of launch 0 per 64-instruction region ({tot} samples; regions under 0.5 % omitted) ---")
        for b in range(0, len(data), 64):
            chunk = data[b:b + 64]
            s = sum(int(r[ix["# Samples"]] or 0) for r in chunk)
            if s / tot < 0.005:
                continue
            ops = collections.Counter()
            for r in chunk:
                t = r[ix["Source"]].split()
                if t:
                    ops[t[1] if t[0].startswith("@") and len(t) > 1 else t[0]] += 1
            st = collections.Counter()
            for r in chunk:
                for k in stalls:
                    st[k] += int(r[ix[k]] or 0)
            print(f"  +{b:5d} {100 * s / tot:5.2f} %  opcodes {dict(ops.most_common(3))}  stalls {dict(st.most_common(3))}")


if __name__ == "__main__":
    main()
