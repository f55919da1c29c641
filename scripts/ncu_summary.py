"""Summarise an `ncu --set full` report of the two render kernels into profiles/<name>.json (the file bench.py reads
for roofline.traffic): python scripts/ncu_summary.py gpurun_out/prof_f.ncu-rep profiles/r2_ncu_traffic.json "<source note>" [workload]."""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "duration_ms_under_ncu": "gpu__time_duration.sum",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warp_inst": "smsp__inst_executed.sum",
    "avg_active_threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "achieved_occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "registers": "launch__registers_per_thread",
    "alu_pipe_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "stall_barrier_per_issue": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "stall_long_scoreboard_per_issue": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_wait_per_issue": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "warps_active_per_scheduler": "smsp__warps_active.avg.per_cycle_active",
}
UNIT_SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3,
              "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    workload = sys.argv[4] if len(sys.argv) > 4 else "C3ii_4k"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "")
        short = "primary_kernel" if "primary_kernel" in name else "shade_kernel" if "shade_kernel" in name else None
        if short is None or short in kernels:
            continue
        k = {"kernel": name.split("(")[0].replace("vxrt::", "").replace("(bool)", "")}
        for key, metric in KEYS.items():
            v = d.get(metric, "")
            if v == "":
                continue
            x = float(v.replace(",", ""))
            x *= UNIT_SCALE.get(u.get(metric, ""), 1.0) if key.startswith(("dram", "duration")) else 1.0
            k[key] = round(x, 6)
        kernels[short] = k
    json.dump({"workload": workload, "source": note, "kernels": kernels}, open(out, "w"), indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == "__main__":
    main()
