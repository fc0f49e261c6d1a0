#!/usr/bin/env python3
"""Condense the ncu captures under gpurun_out/ into the tracked profiles/ directory (round 1)."""
import csv, json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
os.makedirs('profiles', exist_ok=True)
KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread']
def load(path, pick=None):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out=[]
    for r in data:
        name = r[hdr.index('Kernel Name')]
        if pick and pick not in name: continue
        d={'kernel':name}
        for k in KEYS:
            if k in hdr:
                v=r[hdr.index(k)].replace(',','')
                try: d[k]=float(v)
                except ValueError: d[k]=v
                d[k+'__unit']=units[hdr.index(k)]
        out.append(d)
    return out
def to_bytes(d,k):
    v=d[k]; u=d[k+'__unit']
    return v*{'Gbyte':1e9,'Mbyte':1e6,'Kbyte':1e3,'byte':1}.get(u,1)
runs = [('v1 branchy kernel, 16-byte register-buffered stores (`movi_pml_kernel<u32,1>`)','gpurun_out/prof_pml_raw.csv','movi_pml_kernel<'),
        ('random-gather probe on the same 124 MB table (`gather_probe_kernel<8,dep>`)','gpurun_out/prof_v2_raw.csv','gather_probe_kernel<8, 1>'),
        ('v2 convergent kernel + L2 hints, 8-byte stores (`movi_pml_kernel2<u32,true>`)','gpurun_out/prof_v2_raw.csv','movi_pml_kernel2'),
        ('v4 first cut: + sector staging in shared memory (`movi_pml_kernel3`)','gpurun_out/prof_v4_raw.csv','movi_pml_kernel3'),
        ('v4 after cleanup = bench default (`movi_pml_kernel3<u32,false>`), captured under `bench.py`','gpurun_out/prof_pml_b_raw.csv','movi_pml_kernel3')]
md=["# Round 1 ncu evidence -- PML kernel on the c2 workload (1 M x 150 bp reads, 124 MB row table), 1 x B200\n",
    "All captures: `ncu --set full --clock-control none --import-source on` under `gpurun`; one launch = one pass over 1.5e8 bases.",
    "Times under ncu are serialised / cold-cache; the bench numbers are in `r01_bench.json`.\n"]
md += ["| capture | time ms | DRAM read GB | DRAM write MB | DRAM read B / base | L2 hit % | L1 hit % | warp inst (1e9) | lanes / inst | issue active % | warps active % | regs |", "|"+"---|"*12]
for title,path,pick in runs:
    ds = load(path,pick)
    if not ds: continue
    d = max(ds, key=lambda x: x['gpu__time_duration.sum'])   # the full-batch launch (bench.py's e2e leg launches 16 M-base slices too)
    rd, wr = to_bytes(d,'dram__bytes_read.sum'), to_bytes(d,'dram__bytes_write.sum')
    nb = 1.5e8 if 'probe' not in title else d['l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']
    md.append("| %s | %.3f | %.2f | %.0f | %.1f | %.1f | %.1f | %.2f | %.1f | %.1f | %.1f | %d |" % (
        title, d['gpu__time_duration.sum'], rd/1e9, wr/1e6, rd/nb, d['lts__t_sector_hit_rate.pct'], d['l1tex__t_sector_hit_rate.pct'],
        d['smsp__inst_executed.sum']/1e9, d['smsp__thread_inst_executed_per_inst_executed.ratio'], d['smsp__issue_active.avg.pct_of_peak_sustained_active'],
        d['sm__warps_active.avg.pct_of_peak_sustained_active'], int(d['launch__registers_per_thread'])))
last = max(load('gpurun_out/prof_pml_b_raw.csv','movi_pml_kernel3'), key=lambda x: x['gpu__time_duration.sum'])
rd, wr = to_bytes(last,'dram__bytes_read.sum'), to_bytes(last,'dram__bytes_write.sum')
json.dump({"kernel":"movi_pml_kernel3<uint32_t,false>","workload":"c2","dram_bytes_per_launch": rd+wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "algorithmic_bytes_per_launch": 32*1.5e8, "source":"profiles/r01_pml_v4_bench_raw.csv (ncu --set full, the full-batch launch)"}, open('profiles/pml_c2_traffic.json','w'), indent=1)
md += ["",
"(the probe row is per load, not per base: 1.55e8 independent random 8-byte loads)",
"",
"Reading the table:",
"* **v1 -> v2**: the branchy state machine kept 12 of 32 lanes busy per issued instruction; predicate-driven transitions raise that to 26 and drop the register count to 32, but time barely moves because v1/v2 were DRAM-traffic bound: 74 B of DRAM read per base and 2.1x write amplification -- every lane touched its read bytes and its output 8 bytes at a time, tens of microseconds apart, while the row fetches turn the 126 MB L2 over every ~35 us, so every 64-byte atom of the two streams was fetched / written back several times.",
"* **v4**: read bytes and results are staged per lane in shared memory and cross the memory system one aligned 32-byte sector at a time: DRAM read 11.1 -> 8.3 GB, write amplification 2.15x -> 1.1x, L2 hit 25 -> 34 %.  After that the kernel is ~2/3 instruction-issue bound (issue active 62 %, ~185 instructions per warp iteration, 1.59 iterations per base) and ~1/3 exposed row-fetch latency (long scoreboard on the first use of the row word: `r01_pml_v5_hot_sass.txt`).",
"* **probe**: the practical ceiling of the memory system for independent random 8-byte loads over the same table: 1.24e11 loads/s, DRAM at 5.4 TB/s (83 % of the measured streaming peak) with a 29 % L2 hit rate -- the 124 MB table does not stay L2 resident (effective capacity for this pattern ~ 126 MB / 2 dies / half-filled 128-byte lines; an LRU simulation of the access trace gives the same ~30 % at 32 MB).",
"* Physical floor for PML on this table: one 64-byte DRAM atom per base x (1 - L2 hit) ~ 45 B/base; the algorithmic accounting of SURVEY 8(d) is 32 B/base (`roofline.achieved`), the measured traffic is `roofline.traffic` (`pml_c2_traffic.json`).",
"* Random-gather probe vs table size (`r01_probe.jsonl`, first GPU session): 64 MB 2.8e11 loads/s (L2), 1 GB 4.3e10, 8-32 GB 3.7e10, 96 GB 1.4e10 loads/s -- beyond ~0.25 GB the TLB reach, not DRAM, sets the ceiling for fully random row fetches; this is the number the human-scale (c4) configuration will be measured against.",
"",
"Files: `r01_launches_bench.csv` (launch list of `bench.py --steps 4 --warmup 3 --no-cpu-baseline`: every launch of the timed region is `movi_pml_kernel3`; the torch fill kernels are buffer set-up), `r01_pml_*_raw.csv` (ncu raw pages), `r01_pml_v5_hot_sass.txt` (per-instruction execution counts and stall samples of the hot loop), `r01_bench.json` / `r01_bench_reference_arm.json` (bench lines of the same session), `r01_explore.jsonl` (variant / launch-shape / L2-option sweeps)."]
open('profiles/r01_summary.md','w').write("\n".join(md)+"\n")
for src,dst in (('gpurun_out/launches_b.csv','r01_launches_bench.csv'),('gpurun_out/prof_pml_raw.csv','r01_pml_v1_raw.csv'),('gpurun_out/prof_v2_raw.csv','r01_pml_v2_and_probe_raw.csv'),
                ('gpurun_out/prof_v4_raw.csv','r01_pml_v4_first_raw.csv'),('gpurun_out/prof_pml_b_raw.csv','r01_pml_v4_bench_raw.csv'),('gpurun_out/bench_b.json','r01_bench.json'),
                ('gpurun_out/bench_ref.json','r01_bench_reference_arm.json'),('gpurun_out/explore_pml.jsonl','r01_explore.jsonl'),('gpurun_out/explore.jsonl','r01_probe.jsonl')):
    if os.path.exists(src): shutil.copy(src, os.path.join('profiles',dst))
hot = '/root/.claude/projects/-root-repo/015383a3-1bd8-42e1-b993-330841f1bdeb/tool-results/b0k74ti2o.txt'
if os.path.exists(hot): shutil.copy(hot,'profiles/r01_pml_v5_hot_sass.txt')
print(open('profiles/r01_summary.md').read()[:2500])
