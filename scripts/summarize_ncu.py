"""Text summary of an `ncu --set full` report: per launch, the metrics the DESIGN.md roofline discussion uses.

    python scripts/summarize_ncu.py gpurun_out/x.ncu-rep > profiles/x_summary.txt
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors (from SMs)"),
    ("lts__t_sectors_srcunit_tex_op_write.sum", "L2 write sectors (from SMs)"),
    ("lts__t_sectors_srcunit_tex_op_red.sum", "L2 reduction sectors"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "L1->XBAR request cycles active %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "XBAR->SM read bytes"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "XBAR->SM read rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    print("# %s  (ncu --set full --clock-control none; per-launch values, cold caches, serialised)" % rep.split("/")[-1])
    for n, r in enumerate(rows[2:]):
        print("\n[%d] %s" % (n, r[ik][:160]))
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("    %-42s %s %s" % (label, r[i], units[i]))


if __name__ == "__main__":
    main(sys.argv[1])
