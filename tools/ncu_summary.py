"""Summarise an .ncu-rep (raw page + per-source-line hot spots) into text for profiles/."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], (rows[2] if len(rows) > 2 else rows[1])
d = dict(zip(hdr, vals))
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__warps_active.avg.per_cycle_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum']
print("== kernel:", d.get('Kernel Name', '?'))
for k in keys:
    if k in d: print(f"{k} = {d[k]}")
st = []
for h in hdr:
    if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued'):
        try: st.append((float(d[h]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
        except ValueError: pass
tot = sum(v for v, _ in st) or 1
print("== warp stall sampling (all samples):")
for v, h in sorted(st, reverse=True)[:10]: print(f"  {v / tot * 100:5.1f}%  {h}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hd, data = None, None, []
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hd = r; continue
    if hd is None or len(r) < 10 or r[2] != '-': continue
    try: data.append((float(r[6]), float(r[7]), float(r[8]), cur, r[0], r[1][:96]))
    except ValueError: pass
ts, ti = sum(x[0] for x in data) or 1, sum(x[1] for x in data) or 1
print(f"== source hot spots (stall samples {ts:.0f}, warp instructions {ti:.0f}, avg active threads/inst {sum(x[2] for x in data) / ti:.2f}):")
for x in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"  {x[0] / ts * 100:5.2f}% samp {x[1] / ti * 100:5.2f}% inst thr/inst={x[2] / max(x[1], 1):5.1f} | {x[3]}:{x[4]} {x[5]}")
