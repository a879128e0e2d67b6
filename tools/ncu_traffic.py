"""DRAM traffic of the x-update launches of ONE ADMM iteration from an `ncu --set full` report:
python tools/ncu_traffic.py gpurun_out/prof_xupdate_r2a.ncu-rep profiles/r02_ncu_xupdate_traffic.json
(the report must hold the consecutive phase launches of one iteration: -k regex:admm_xupdate -c 6)."""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
launches = []
for r in rows[2:]:
    rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
    wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
    launches.append({"kernel": r[col["Kernel Name"]], "us": float(r[col["gpu__time_duration.sum"]]),
                     "dram_read_bytes": rd, "dram_write_bytes": wr,
                     "dmma_pipe_pct": float(r[col["sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"]]),
                     "registers": int(r[col["launch__registers_per_thread"]])})
tot = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches)
json.dump({"source": rep.split("/")[-1] + " (ncu --set full --clock-control none, cfg-2, packed-count x-update)",
           "dram_bytes_per_xupdate": tot, "launches_per_xupdate": len(launches), "launches": launches},
          open(out, "w"), indent=1)
print(tot / 1e6, "MB per ADMM iteration's x-update over", len(launches), "launches")
