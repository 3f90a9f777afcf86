"""profiles/current_launch_metrics.json from an ncu launch list of the bench command.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct \
      --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline
  python tools/launch_metrics.py gpurun_out/launches.csv profiles/r2_launches_bench.csv

Takes the LAST reference-precision render_wave launch of the device-timed leg (the timed step; warm-up precedes it, the
end-to-end leg's launches follow -- all the same frame size) and writes its counters together with a hash of the kernel sources,
so that bench.py can tell whether the numbers still describe the code it is running (roofline.traffic, roofline_issue)."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

src = sys.argv[1]
keep = sys.argv[2] if len(sys.argv) > 2 else None
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ix = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Grid Size", "Block Size")}
launches = {}
for r in rows:
    d = launches.setdefault(int(r[ix["ID"]]), {"kernel": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
waves = [(i, d) for i, d in sorted(launches.items()) if "render_wave<double" in d["kernel"]]
if not waves:
    raise SystemExit("no render_wave<double, ...> launch in " + src)
# bench.py --steps 1 --warmup 1 --no-extras: warm-up frame, timed frame, then the e2e leg
i, d = waves[1] if len(waves) > 1 else waves[0]
total_ns = sum(x.get("gpu__time_duration.sum", 0) for x in launches.values())
step_ids = [k for k in launches if waves[1][0] <= k < (waves[2][0] if len(waves) > 2 else 1 << 30)] if len(waves) > 1 else [i]
step_ns = sum(launches[k].get("gpu__time_duration.sum", 0) for k in step_ids)
try:
    commit = subprocess.check_output(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], text=True).strip()
except Exception:
    commit = None
out = {
    "source": os.path.relpath(keep, ROOT) if keep else os.path.basename(src),
    "source_hash": bench.source_hash(), "commit_at_capture": commit,
    "command": "python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline (under ncu --clock-control none)",
    "kernel": d["kernel"], "grid": d["grid"], "block": d["block"], "wave_launches_per_step": 1,
    "duration_ms_under_ncu": d.get("gpu__time_duration.sum", 0) / 1e6,
    "share_of_step": d.get("gpu__time_duration.sum", 0) / step_ns if step_ns else None,
    "dram_bytes_read": d.get("dram__bytes_read.sum"), "dram_bytes_write": d.get("dram__bytes_write.sum"),
    "inst_executed": d.get("smsp__inst_executed.sum"), "l2_hit_rate_pct": d.get("lts__t_sector_hit_rate.pct"),
}
json.dump(out, open(os.path.join(ROOT, "profiles", "current_launch_metrics.json"), "w"), indent=1)
if keep:
    shutil.copyfile(src, keep)
print(json.dumps(out, indent=1))
