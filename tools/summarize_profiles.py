"""Turn gpurun_out/*.ncu-rep / launch CSVs into small tracked summaries under profiles/.

  python tools/summarize_profiles.py <tag> [launches.csv] [rep1.ncu-rep rep2.ncu-rep ...]
"""

import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
  'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
  'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
  'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'dram__bytes_read.sum',
  'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
  'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
  'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
  'smsp__inst_executed.sum', 'smsp__cycles_active.avg', 'smsp__average_warp_latency_per_inst_issued.ratio',
  'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max',
  'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
]


def launches(tag, path):
  lines = [ln for ln in open(path) if not ln.startswith('==')]
  agg = collections.defaultdict(lambda: [0, 0.0])
  for row in csv.DictReader(lines):
    try:
      v = float(row['Metric Value'].replace(',', ''))
    except (ValueError, KeyError):
      continue
    name = row['Kernel Name'].split('(')[0].replace('void ', '')
    agg[name][0] += 1
    agg[name][1] += v / 1e6  # ns -> ms
  tot = sum(v[1] for v in agg.values())
  out = os.path.join(ROOT, 'profiles', f'{tag}_launches_summary.md')
  with open(out, 'w') as f:
    f.write(f'# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none), one optimizer step of bench.py\n\n')
    f.write('Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n')
    f.write('| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
      f.write(f'| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f} % |\n')
    f.write(f'| **total** | {sum(v[0] for v in agg.values())} | {tot:.3f} | |\n')
  print('wrote', out)


def report(tag, path):
  raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(raw.splitlines()))
  hdr, units, data = rows[0], rows[1], rows[2:]
  name_i = hdr.index('Kernel Name')
  base = os.path.basename(path).replace('.ncu-rep', '')
  out = os.path.join(ROOT, 'profiles', f'{base}_summary.csv')
  with open(out, 'w') as f:
    w = csv.writer(f)
    w.writerow(['metric', 'unit'] + [r[name_i].split('(')[0][:60] for r in data])
    for i, h in enumerate(hdr):
      if h in KEYS:
        w.writerow([h, units[i]] + [r[i] for r in data])
  print('wrote', out)


if __name__ == '__main__':
  tag = sys.argv[1]
  os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
  for p in sys.argv[2:]:
    if p.endswith('.csv'):
      launches(tag, p)
    else:
      report(tag, p)
