"""Condenses an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / profiles/ quote.
    python tools/ncu_summary.py gpurun_out/prof_pass1_kernel.ncu-rep [...]"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("launch__occupancy_limit_registers", "CTAs/SM (reg limit)"),
    ("launch__waves_per_multiprocessor", "waves"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "tensor hmma cycles active (avg/SM)"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "TMEM/tensor-mem active %"),
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    for path in sys.argv[1:]:
        hdr, units, launches = raw(path)
        for vals in launches:
            d = {h.split(".", 2)[-1] if h.split(".")[0].isupper() else h: (v, u) for h, u, v in zip(hdr, units, vals)}
            d.update({h: (v, u) for h, u, v in zip(hdr, units, vals)})
            name = d.get("Kernel Name", ("?", ""))[0]
            print("## %s  (%s)" % (name.split("(")[0], path.split("/")[-1]))
            for key, label in KEYS:
                hit = [k for k in d if k.endswith(key)]
                if hit:
                    v, u = d[hit[0]]
                    print("- %s: %s %s" % (label, v, u))
            stalls = []
            for k, (v, u) in d.items():
                if "smsp__average_warps_issue_stalled_" in k and k.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(v), k.split("stalled_")[1].replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            if stalls:
                print("- top warp stall reasons (warps per issue): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:6]))
            print()


if __name__ == "__main__":
    main()
