#!/usr/bin/env python3
"""Second part of the round-1 profile summary: the final default kernels (PML variant 5, count / ZML) captured under bench.py.
Appends to profiles/r01_summary.md, refreshes the traffic files bench.py reads and the launch list; touches nothing else."""
import csv, json, os, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread']


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for r in data:
        d = {'kernel': r[hdr.index('Kernel Name')]}
        for k in KEYS:
            v = r[hdr.index(k)].replace(',', '')
            d[k] = float(v); d[k + '__unit'] = units[hdr.index(k)]
        out.append(d)
    return max(out, key=lambda x: x['gpu__time_duration.sum'])


def to_bytes(d, k):
    return d[k] * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}.get(d[k + '__unit'], 1)


runs = [('PML, c2: `movi_pml_kernel4<u32,false,6,R8>` (variant 5, default)', 'gpurun_out/prof_pml_f_raw.csv', 'r01_pml_v5_bench_raw.csv', 'pml_c2_traffic.json', 'c2', 32, 1.5e8),
        ('ZML, c3: `movi_bs_kernel2<u32,ZML,R8>`', 'gpurun_out/prof_f_c3-zml_raw.csv', 'r01_zml_bench_raw.csv', 'zml_c3_traffic.json', 'c3-zml', 64, 1.5e8),
        ('count, c3: `movi_bs_kernel2<u32,COUNT,R8>`', 'gpurun_out/prof_f_c3-count_raw.csv', 'r01_count_bench_raw.csv', 'count_c3_traffic.json', 'c3-count', 64, None)]
md = ["", "## Final default kernels of the round (captured under `bench.py --steps 4 --warmup 3 --no-cpu-baseline`, `tools/final_run.sh`)", "",
      "| kernel | time ms (under ncu) | DRAM read GB | DRAM write MB | L2 hit % | L1 hit % | warp inst (1e9) | lanes / inst | issue active % | warps active % | regs |",
      "|" + "---|" * 11]
for title, src, dst, tfile, wl, algo, nb in runs:
    d = load(src)
    rd, wr = to_bytes(d, 'dram__bytes_read.sum'), to_bytes(d, 'dram__bytes_write.sum')
    md.append("| %s | %.3f | %.2f | %.0f | %.1f | %.1f | %.2f | %.1f | %.1f | %.1f | %d |" % (
        title, d['gpu__time_duration.sum'], rd / 1e9, wr / 1e6, d['lts__t_sector_hit_rate.pct'], d['l1tex__t_sector_hit_rate.pct'],
        d['smsp__inst_executed.sum'] / 1e9, d['smsp__thread_inst_executed_per_inst_executed.ratio'],
        d['smsp__issue_active.avg.pct_of_peak_sustained_active'], d['sm__warps_active.avg.pct_of_peak_sustained_active'], int(d['launch__registers_per_thread'])))
    shutil.copy(src, os.path.join('profiles', dst))
    json.dump({"kernel": d['kernel'], "workload": wl, "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
               "algorithmic_bytes_per_base": algo, "source": "profiles/%s (ncu --set full, one full-batch launch)" % dst},
              open(os.path.join('profiles', tfile), 'w'), indent=1)
md += ["",
       "* PML variant 5 against variant 4 (table above): warp instructions 1.70 -> 1.5-1.6e9, DRAM traffic 8.79 -> 8.3 GB (evict_first on the streamed bytes), time in the bench 2.12 -> 1.94 ms.  Still the same regime: ~68 % issue active (the ALU pipe at ~70 % of its peak, the FMA pipe at 15 %), 34 % L2 hits on the 124 MB table.",
       "* ZML moves the same DRAM bytes per base as PML (the two interval ends sit in the same or neighbouring rows once a few bases have matched) but issues 1.7x the instructions (two walkers per lane): it is issue bound, not memory bound; 64 B/base is the SURVEY's accounting, 54 B/base is what DRAM sees.",
       "* count stops at the first failed extension (mean 104 of 150 bases): 1.41e9 instructions, 5.3 GB.",
       "* `r01_launches_bench.csv`: launch list of the same bench command -- the timed region holds `movi_pml_kernel4` launches only (full-batch launches on two alternating streams; the 82..328-block launches are the slices of the host-buffer legs `e2e`, `e2e_classify` (+ `classify_kernel`) and `e2e_ml8` (+ `narrow_ml_kernel`)).",
       "* `r01_pipeline_sweep.jsonl`, `r01_host_probe.jsonl`: host-buffer pipeline shapes and what the box's PCIe / host memory can do (H2D 55 GB/s, D2H 57 GB/s alone, 150 MB up + 300 MB down together 5.9 ms)."]
old = open('profiles/r01_summary.md').read()
marker = "\n## Final default kernels of the round"
if marker in old:
    old = old[:old.index(marker)] + "\n"          # regenerate this section
open('profiles/r01_summary.md', 'w').write(old.rstrip("\n") + "\n" + "\n".join(md) + "\n")
shutil.copy('gpurun_out/launches_f.csv', 'profiles/r01_launches_bench.csv')
print("\n".join(md))
