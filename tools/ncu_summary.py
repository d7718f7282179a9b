"""Print the handful of ncu metrics this repo's roofline arguments rest on, from an .ncu-rep (needs `ncu` on PATH)."""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed_op_shared_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]
def main(path, extra):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        d = dict(zip(h, zip(u, v)))
        print("==", d.get("Kernel Name", ("", "?"))[1][:60], "grid", d.get("launch__grid_size", ("", "?"))[1])
        for k in h:
            if k in WANT or any(e in k for e in extra):
                if k in d and d[k][1] != "":
                    print("   %-90s %-10s %s" % (k, d[k][0], d[k][1]))
if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
