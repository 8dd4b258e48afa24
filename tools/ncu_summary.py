"""Summarise an ncu report exported with `--page raw --csv` and `--page source --csv`:
headline metrics, stall reasons per issue, and stall samples by code region.

    python tools/ncu_summary.py raw.csv src.csv [frames]
"""
import csv
import sys

raw, src = sys.argv[1], sys.argv[2]
frames = float(sys.argv[3]) if len(sys.argv) > 3 else None
rows = list(csv.reader(open(raw)))
hdr, vals = rows[0], rows[2]
metrics = dict(zip(hdr, vals))
keys = [
    'gpu__time_duration.sum', 'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'sm__warps_active.avg.per_cycle_active', 'launch__registers_per_thread',
    'dram__bytes_read.sum', 'dram__bytes_write.sum']
for key in keys:
    value = metrics.get(key)
    extra = ''
    if frames and key in ('smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'):
        extra = f'  ({float(value) / frames:.1f} per frame)'
    print(f'{key}: {value}{extra}')
stalls = {
    k.split('issue_stalled_')[1].split('_per_issue')[0]: float(v)
    for k, v in metrics.items() if 'issue_stalled' in k and 'per_issue_active' in k}
print('stalls per issue:', ', '.join(
    f'{k} {v:.2f}' for k, v in sorted(stalls.items(), key=lambda x: -x[1]) if v > 0.05))

rows = list(csv.reader(open(src)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
total = sum(int(r[idx['# Samples']] or 0) for r in data)
executed = sum(int(r[idx['Instructions Executed']] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print('samples', total, 'instructions', executed)
chunk = 60
for i in range(0, len(data), chunk):
    seg = data[i:i + chunk]
    s = sum(int(r[idx['# Samples']] or 0) for r in seg)
    ex = sum(int(r[idx['Instructions Executed']] or 0) for r in seg)
    if s > total * 0.012:
        st = {c: sum(int(r[idx[c]] or 0) for r in seg) for c in stall_cols}
        top = sorted(st.items(), key=lambda x: -x[1])[:4]
        ops = {}
        for r in seg:
            parts = r[idx['Source']].split()
            op = parts[1] if parts and parts[0].startswith('@') else (parts[0] if parts else '')
            ops[op] = ops.get(op, 0) + int(r[idx['Instructions Executed']] or 0)
        topo = sorted(ops.items(), key=lambda x: -x[1])[:3]
        print(f'{i:5d} samples {100 * s / total:5.1f}%  instr {100 * ex / executed:5.1f}%', top, topo)
