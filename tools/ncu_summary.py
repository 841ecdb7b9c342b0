"""Summarise an .ncu-rep (ncu --set full) into a small text + json: duration, DRAM traffic, tensor/DRAM utilisation."""
import csv, json, subprocess, sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def main(rep, out_prefix):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        d = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                v = r[i]
                try:
                    v = float(v.replace(",", "")) * UNIT.get(units[i], 1.0) if k != "Kernel Name" else v
                except ValueError:
                    pass
                d[k] = v
        d["dram_bytes"] = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        recs.append(d)
    with open(out_prefix + ".json", "w") as f:
        json.dump(recs, f, indent=1)
    with open(out_prefix + ".txt", "w") as f:
        for d in recs:
            f.write("%s\n  grid %s block %s regs %s\n  duration %.3f ms  dram read %.3f GB write %.3f GB  (dram %.1f%% of peak)\n"
                    "  tensor pipe active %.1f%%  sm throughput %.1f%%  warps active %.1f%%  issue active %.1f%%\n" % (
                        d["Kernel Name"], int(d["launch__grid_size"]), int(d["launch__block_size"]), int(d["launch__registers_per_thread"]),
                        d["gpu__time_duration.sum"] * 1e3, d["dram__bytes_read.sum"] / 1e9, d["dram__bytes_write.sum"] / 1e9,
                        d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"],
                        d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", float("nan")),
                        d["sm__throughput.avg.pct_of_peak_sustained_elapsed"],
                        d["sm__warps_active.avg.pct_of_peak_sustained_active"],
                        d["smsp__issue_active.avg.pct_of_peak_sustained_active"]))
    print(open(out_prefix + ".txt").read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
