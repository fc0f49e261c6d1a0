set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; cat gpurun_out/bench_f.json
for w in c3-zml c3-count; do python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/bench_f_$w.json 2>/dev/null; cat gpurun_out/bench_f_$w.json; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_f.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_lf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:movi_pml_kernel4 -s 4 -c 1 -f -o gpurun_out/prof_pml_f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pf.log 2>&1
ncu -i gpurun_out/prof_pml_f.ncu-rep --page raw --csv > gpurun_out/prof_pml_f_raw.csv
ncu -i gpurun_out/prof_pml_f.ncu-rep --page source --csv > gpurun_out/prof_pml_f_src.csv 2>/dev/null
for w in c3-zml c3-count; do
ncu --set full --clock-control none --import-source on -k regex:movi_bs_kernel2 -s 4 -c 1 -f -o gpurun_out/prof_f_$w python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f_$w.log 2>&1
ncu -i gpurun_out/prof_f_$w.ncu-rep --page raw --csv > gpurun_out/prof_f_${w}_raw.csv
done
ls -la gpurun_out/*.ncu-rep
