"""Condenses an `ncu --set full` capture of the extend launches of one wave (primary + bounces) into the files
profiles/ keeps: <tag>_extend_ncu_summary.txt (scripts/ncu_summary.py on the raw page) and extend_traffic.json
(the DRAM bytes per ray bench.py's roofline.traffic uses).
usage: make_extend_profile.py capture.ncu-rep tag rays_in_capture "kernel description" """
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag, rays, desc = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
tmp = "/tmp/_extend_raw.csv"
open(tmp, "w").write(raw)
summary = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), tmp], capture_output=True,
                         text=True, check=True).stdout
out = os.path.join(ROOT, "profiles", f"{tag}_extend_ncu_summary.txt")
open(out, "w").write(summary)
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
def col(name, scale_by_unit=True):
    i = hdr.index(name)
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(units[i], 1.0) if scale_by_unit else 1.0
    return [float(r[i].replace(",", "")) * mult for r in rows[2:]]
per_launch = [a + b for a, b in zip(col("dram__bytes_read.sum"), col("dram__bytes_write.sum"))]
dur_unit = units[hdr.index("gpu__time_duration.sum")]
dur = [v * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[dur_unit] for v in col("gpu__time_duration.sum", False)]
doc = {
    "kernel": desc,
    "source": f"ncu --set full --clock-control none, profiles/{tag}_extend_ncu_summary.txt (the {len(per_launch)} extend "
              f"launches of one 1080p x 4 spp wave: primary + bounces, {rays} rays)",
    "dram_bytes_per_launch": sum(per_launch) / len(per_launch),
    "per_launch": per_launch,
    "rays_in_capture": rays,
    "dram_bytes_per_ray": sum(per_launch) / rays,
    "duration_ms_under_ncu": dur,
    "issue_slots_busy_pct_per_launch": [round(v, 1) for v in col("sm__throughput.avg.pct_of_peak_sustained_elapsed", False)],
    "active_lanes_per_instruction_per_launch": [round(v, 2) for v in col("smsp__thread_inst_executed_per_inst_executed.ratio", False)],
    "stall_cycles_per_instruction_long_scoreboard_per_launch":
        [round(v, 2) for v in col("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", False)],
    "warps_per_scheduler": round(max(col("smsp__warps_active.avg.per_cycle_active", False)), 1),
}
json.dump(doc, open(os.path.join(ROOT, "profiles", "extend_traffic.json"), "w"), indent=1)
print(out)
print(json.dumps(doc, indent=1))
