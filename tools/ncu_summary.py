"""Summarises an .ncu-rep: headline metrics, stall ratios, hottest SASS lines.  usage: ncu_summary.py rep [kernel_index]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
vals = rows[2 + kidx]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
print("kernel:", d.get("Kernel Name", ("?",))[0][:80], "grid", d.get("Grid Size", ("?",))[0], "block", d.get("Block Size", ("?",))[0])
keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed']
for k in keys:
    if k in d: print(f"  {k:82s} {d[k][0]:>16s} {d[k][1]}")
st = []
for h in hdr:
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        try: st.append((float(d[h][0]), h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')))
        except ValueError: pass
print("  stalls per issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-id", f":::{kidx+1}"] if kidx else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = None
for i, r in enumerate(rows):
    if 'Address' in r and 'Source' in r: h2 = i; break
if h2 is not None:
    hd = rows[h2]
    ia, isrc, iss, iex = hd.index('Address'), hd.index('Source'), hd.index('Warp Stall Sampling (All Samples)'), hd.index('Instructions Executed')
    data = []
    for r in rows[h2 + 1:]:
        try: data.append((int(r[iss]), r[ia][-5:], r[isrc], int(r[iex])))
        except (ValueError, IndexError): pass
    tot = sum(x[0] for x in data) or 1
    print(f"  hottest SASS (of {tot} samples):")
    for s, a, sc, ex in sorted(data, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 14]:
        print(f"   {100*s/tot:5.1f}%  {a}  ex={ex:9d}  {sc[:90]}")
