"""Turn the outputs of tools/round_sweep.sh (gpurun_out/<tag>_*) into the committed evidence under profiles/:
bench lines, kernel table, launch-list summary and a table of the `ncu --set full` captures.
python tools/make_profile_notes.py r01d"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def last_json(path):
    with open(path) as f:
        lines = [l for l in f.read().strip().splitlines() if l.startswith('{')]
    return json.loads(lines[-1])


for name in ('bench', 'bench_reference', 'bench_fpn', 'bench_backbone', 'bench_infer', 'bench_n2'):
    src = os.path.join(G, f'{tag}_{name}.json')
    if os.path.exists(src):
        with open(os.path.join(P, f'{tag}_{name}.json'), 'w') as f:
            f.write(json.dumps(last_json(src)) + '\n')
src = os.path.join(G, f'{tag}_kernel_table_train_upernext.json')
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, f'{tag}_kernel_table_train_upernext.json'))
launches = os.path.join(G, f'launches_{tag}.csv')
if os.path.exists(launches):
    subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'summarize_launches.py'), launches,
                    os.path.join(P, f'{tag}_launches_summary.md')], check=True, stdout=subprocess.DEVNULL)

KEYS = [('gpu__time_duration.sum', 'duration'), ('sm__cycles_elapsed.avg.per_second', 'SM clock'), ('dram__bytes_read.sum', 'DRAM read'),
        ('dram__bytes_write.sum', 'DRAM write'), ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe'),
        ('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'FMA-heavy pipe'),
        ('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed', 'ALU pipe'),
        ('sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'issue active'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smem wavefronts'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit'), ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'),
        ('launch__block_size', 'block'), ('launch__cluster_dim_x', 'cluster')]
rows_out = []
traffic = {}
for rep in sorted(f for f in os.listdir(G) if f.startswith('prof_') and f.endswith(f'_{tag}.ncu-rep')):
    out = subprocess.run(['ncu', '-i', os.path.join(G, rep), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')].replace('<unnamed>::', '').replace('void ', '').split('(')[0]
        cells = {}
        for k, lab in KEYS:
            if k in hdr:
                v, u = r[hdr.index(k)], units[hdr.index(k)]
                try:
                    fv = float(v)
                    v = f'{fv:.3g}' if abs(fv) < 1000 else f'{fv:.0f}'
                except ValueError:
                    pass
                cells[lab] = f'{v} {u}'.strip()
        st = {k.replace('smsp__pcsamp_warps_issue_stalled_', ''): float(r[i] or 0) for i, k in enumerate(hdr)
              if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('not_issued')}
        tot = sum(st.values()) or 1.0
        cells['top stalls'] = ', '.join(f'{k} {100 * v / tot:.0f}%' for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
        rows_out.append((rep, name, cells))
with open(os.path.join(P, f'{tag}_ncu_kernels.md'), 'w') as f:
    f.write(f'# {tag} — `ncu --set full --clock-control none --import-source on` captures of the dominant kernels\n\n'
            'Produced by `tools/round_sweep.sh` (each profiled command exited 0 without ncu first; the `.ncu-rep` files stay in\n'
            '`gpurun_out/`, 6-15 MB each) and read with `ncu -i … --page raw --csv` by `tools/make_profile_notes.py`.\n'
            'Shapes: head conv = the precise head group at B=32, 640×640 (M = 3 276 800 pixels, K = 9·384, N = 832; NT in\n'
            'CTA-pair mode, TN = its weight gradient); depthwise 7×7 and head-tail backward at their stage-0 / full-size shapes.\n'
            'ncu serialises launches and runs them cold: durations differ from the in-step numbers of the bench.\n\n')
    labs = [lab for _, lab in KEYS] + ['top stalls']
    f.write('| capture | kernel | ' + ' | '.join(labs) + ' |\n|' + '---|' * (len(labs) + 2) + '\n')
    for rep, name, cells in rows_out:
        f.write(f'| {rep} | `{name}` | ' + ' | '.join(cells.get(l, '') for l in labs) + ' |\n')
print(f'wrote profiles/{tag}_*')
