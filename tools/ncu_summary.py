"""Summarise an `ncu --set full` report: per launch duration, DRAM traffic, tensor-pipe and memory utilisation."""
import csv, subprocess, sys

WANT = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%act"),
        ("sm__inst_executed_pipe_uniform.sum", "uniform_inst"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "hmma%"),
        ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_rt%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"), ("lts__t_bytes.sum", "l2_bytes"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("smsp__cycles_active.avg", "cyc"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__issue_active.avg.pct", "issue%")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    name_i = col["Kernel Name"]
    print("%-40s" % "kernel" + "".join("%14s" % s for _, s in WANT if _ in col))
    for r in rows[2:]:
        line = "%-40s" % r[name_i].split("(")[0][-40:]
        for k, s in WANT:
            if k in col:
                line += "%14s" % (r[col[k]][:10] + " " + units[col[k]][:3])
        print(line)


if __name__ == "__main__":
    main(sys.argv[1])
