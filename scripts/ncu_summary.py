#!/usr/bin/env python
"""Summarise an .ncu-rep: headline raw metrics, stall mix, opcode mix per sample, top stalled SASS lines."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
nsamp = float(sys.argv[2]) if len(sys.argv) > 2 else 268435456.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
print("kernel:", r[hdr.index("Kernel Name")])
for w in ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'launch__registers_per_thread', 'smsp__inst_executed.sum', 'lts__t_bytes.sum', 'sm__cycles_elapsed.avg']:
    if w in hdr:
        print("  %-72s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot, ops, T, lines = collections.Counter(), collections.Counter(), 0, []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[0].startswith('0x'):
        if r and r[0] == 'Kernel Name':
            break
        continue
    for h in st:
        try:
            tot[h] += int(r[idx[h]])
        except ValueError:
            pass
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[idx['Source']])
    full = m.group(2) if m else '?'
    op = full.split('.')[0]
    e = int(r[idx['Instructions Executed']])
    T += e
    ops[full if op in ('LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL', 'MUFU', 'SYNCS', 'UBLKCP') else op] += e
    lines.append((int(r[idx['# Samples']]), r[idx['Source']].strip(), {h[6:]: int(r[idx[h]]) for h in st if r[idx[h]] not in ('', '0')}))
s = sum(tot.values())
print("warp instructions %d  = %.2f thread-instr per input sample" % (T, T * 32 / nsamp))
print("stall mix: " + ", ".join("%s %.1f%%" % (k[6:], 100 * v / s) for k, v in tot.most_common(9)))
print("opcode mix (thread instr / sample): " + ", ".join("%s %.2f" % (k, v * 32 / nsamp) for k, v in ops.most_common(26)))
lines.sort(key=lambda t: -t[0])
for n, srcl, d in lines[:12]:
    print("  %6d  %-58s %s" % (n, srcl[:58], d))
