python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final_c2.json 2> gpurun_out/bench_final_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_c2_final.csv python bench.py --steps 3 --warmup 3 --no-graph --no-eval --no-cpu-baseline > /dev/null 2>&1
cut -c1-300 gpurun_out/bench_final_c2.json gpurun_out/bench_final_ref.json
