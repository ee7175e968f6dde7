#!/usr/bin/env python
"""Turn `ncu --set full` captures (.ncu-rep, brought back from the GPU box in gpurun_out/) into the committed evidence:
  profiles/ncu_summary.json           per config tag: kernel, DRAM bytes per launch, duration, hit rates, ... (read by bench.py
                                      for roofline.traffic — regenerate it whenever the dominant kernel changes)
  profiles/r02_ncu_<tag>_<kernel>.txt the metrics behind each entry, as text

  python tools/ncu_summary.py TAG=REP:UNITS[:KERNEL_REGEX] ...
      TAG     bench config tag (c1..c5) or any other label
      REP     path of the .ncu-rep
      UNITS   units (k-mers, windows) one launch processes
Runs here (no GPU needed): it only reads the reports with `ncu -i`.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_summary.json")

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum",
    "lts__t_sectors_srcunit_tex_lookup_hit.sum", "lts__t_requests_srcunit_tex.sum", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio",
]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * mult.get(unit, 1)


def to_ms(v, unit):
    mult = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3, "nsecond": 1e-6}
    return float(v) * mult.get(unit, 1)


def main():
    summary = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for arg in sys.argv[1:]:
        tag, rest = arg.split("=", 1)
        parts = rest.split(":")
        rep, units = parts[0], int(parts[1])
        pat = re.compile(parts[2]) if len(parts) > 2 else None
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, check=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, unit = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        launches = [r for r in rows[2:] if (pat is None or pat.search(r[col["Kernel Name"]]))]
        if not launches:
            raise SystemExit(f"{rep}: no launch matches")
        r = launches[-1]  # the last captured launch (warm)
        g = lambda m: (r[col[m]].replace(",", ""), unit[col[m]])  # noqa: E731
        rd, wr = to_bytes(*g("dram__bytes_read.sum")), to_bytes(*g("dram__bytes_write.sum"))
        kname = r[col["Kernel Name"]].strip()
        short = re.sub(r"^void\s+", "", kname).split("(")[0]
        entry = {
            "kernel": short, "source": os.path.join("profiles", f"r02_ncu_{tag}_{re.sub(r'[^A-Za-z0-9_]+', '_', short).strip('_')}.txt"),
            "units_per_launch": units, "launches_captured": len(launches),
            "dram_bytes_per_launch": rd + wr, "dram_read_bytes_per_launch": rd, "dram_write_bytes_per_launch": wr,
            "dram_bytes_per_unit": (rd + wr) / units,
            "duration_ms_under_ncu": to_ms(*g("gpu__time_duration.sum")),
            "lts_hit_rate_pct": float(g("lts__t_sector_hit_rate.pct")[0]),
            "dram_pct_of_peak": float(g("dram__throughput.avg.pct_of_peak_sustained_elapsed")[0]) if "dram__throughput.avg.pct_of_peak_sustained_elapsed" in col else None,
            "sm_pct_of_peak": float(g("sm__throughput.avg.pct_of_peak_sustained_elapsed")[0]),
            "registers_per_thread": int(float(g("launch__registers_per_thread")[0])),
        }
        summary[tag] = entry
        with open(os.path.join(ROOT, entry["source"]), "w") as f:
            f.write(f"# ncu --set full --clock-control none, {os.path.basename(rep)}; {len(launches)} launch(es) captured, last one shown\n")
            f.write(f"# kernel: {kname}\n# units per launch: {units}\n")
            for m in KEEP:
                if m in col:
                    f.write(f"{m:90s} {r[col[m]]:>20s} {unit[col[m]]}\n")
            f.write(f"{'derived: dram bytes per unit':90s} {(rd + wr) / units:20.3f} byte\n")
        print(tag, json.dumps(entry))
    with open(OUT, "w") as f:
        json.dump(summary, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
