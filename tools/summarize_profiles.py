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


def _family(name):
  """bench.py's kernel families (ops.KERNELS_PER_CALL names) from a demangled kernel name."""
  for key, fam in (('gemm_kernel', 'gemm'), ('attn_bwd', 'attn_bwd'), ('attn_delta', 'attn_bwd'), ('dq_finalize', 'attn_bwd'),
                   ('attn_fwd', 'attn_fwd'), ('rmsnorm_bwd', 'rmsnorm_bwd'), ('rmsnorm_fwd', 'rmsnorm_fwd'),
                   ('swiglu_bwd', 'swiglu_bwd'), ('swiglu_fwd', 'swiglu_fwd'), ('ce_', 'ce_fwd_bwd'), ('adamw', 'adamw_step'),
                   ('signsgd', 'signsgd_step'), ('sumsq', 'sumsq'), ('colsum', 'colsum_accum_batched'),
                   ('embed_bwd', 'embed_bwd'), ('embed_fwd', 'embed_fwd')):
    if key in name:
      return fam
  return None


def launches(tag, path):
  """Launch list -> profiles/<tag>_launches_summary.md (+ profiles/traffic.json when the DRAM byte counters are there)."""
  import json

  lines = [ln for ln in open(path) if not ln.startswith('==')]
  agg = collections.defaultdict(lambda: {'ids': set(), 'ms': 0.0, 'rd': 0.0, 'wr': 0.0, 'has_dram': False})
  for row in csv.DictReader(lines):
    try:
      v = float(row['Metric Value'].replace(',', ''))
    except (ValueError, KeyError):
      continue
    name = row['Kernel Name'].split('(')[0].replace('void ', '')
    rec = agg[name]
    rec['ids'].add(row.get('ID'))
    metric, unit = row.get('Metric Name', 'gpu__time_duration.sum'), row.get('Metric Unit', 'ns')
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0)
    if metric.startswith('gpu__time_duration'):
      rec['ms'] += v * scale
    elif metric.startswith('dram__bytes_read'):
      rec['rd'] += v * scale
      rec['has_dram'] = True
    elif metric.startswith('dram__bytes_write'):
      rec['wr'] += v * scale
      rec['has_dram'] = True
  tot = sum(v['ms'] for v in agg.values())
  dram = any(v['has_dram'] for v in agg.values())
  out = os.path.join(ROOT, 'profiles', f'{tag}_launches_summary.md')
  with open(out, 'w') as f:
    f.write(f'# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none) of bench.py --steps 1 --warmup 1\n\n')
    f.write('Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n')
    if dram:
      f.write('DRAM MB/launch = (dram__bytes_read.sum + dram__bytes_write.sum) / launches, same pass.\n')
    f.write('\n| kernel | launches | total ms | share |' + (' DRAM MB/launch |' if dram else '') + '\n')
    f.write('|---|---:|---:|---:|' + ('---:|' if dram else '') + '\n')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]['ms']):
      n = len(v['ids'])
      extra = f' {(v["rd"] + v["wr"]) / n / 1e6:.1f} |' if dram else ''
      f.write(f'| `{k}` | {n} | {v["ms"]:.3f} | {100 * v["ms"] / tot:.1f} % |{extra}\n')
    f.write(f'| **total** | {sum(len(v["ids"]) for v in agg.values())} | {tot:.3f} | |' + (' |' if dram else '') + '\n')
  print('wrote', out)
  if dram:  # per bench.py kernel family: average DRAM bytes per op call (bench.py reads this for roofline.traffic)
    fam = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for k, v in agg.items():
      fm = _family(k)
      if fm:
        main = not any(s in k for s in ('attn_delta', 'dq_finalize', 'ce_reduce', 'ce_stats', 'sumsq_final'))
        fam[fm][0] += len(v['ids']) if main else 0
        fam[fm][1] += v['rd'] + v['wr']
        fam[fm][2] += v['ms']
    tj = {k: {'calls': c, 'dram_bytes_per_call': (b / c if c else None), 'ncu_ms_per_call': (m / c if c else None)}
          for k, (c, b, m) in fam.items()}
    tj['_source'] = f'{tag}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none'
    with open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w') as f:
      json.dump(tj, f, indent=1, sort_keys=True)
    print('wrote profiles/traffic.json')


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
