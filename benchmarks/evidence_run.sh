set -x
python bench.py > gpurun_out/r02f_bench_default.json 2> gpurun_out/r02f_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_ref.err
rm -f gpurun_out/r02f_configs.jsonl
for c in k16 k32 k32_4k dense dense_lab normalised smoothed portrait; do python bench.py --config $c --steps 3 --warmup 3 2>/dev/null | tail -1 >> gpurun_out/r02f_configs.jsonl; done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02f_launches_50.csv python benchmarks/profile_run.py --images 50 --group 50 > gpurun_out/r02f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gabor_tc -c 1 -o gpurun_out/r02f_gabor -f python benchmarks/profile_run.py --images 50 --group 50 > gpurun_out/r02f_ncu2.log 2>&1
python benchmarks/gabor_scale_times.py > gpurun_out/r02f_scale_times.log 2>&1
