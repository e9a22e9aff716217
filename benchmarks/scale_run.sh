N=$1
if [ $N = 1 ]; then
  python bench.py --scaling strong --total-images 10000 > gpurun_out/r02_strong_10k_1gpu.json 2> gpurun_out/strong_1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/weak_$N.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --scaling strong --total-images 10000 > gpurun_out/r02_strong_10k_${N}gpu.json 2> gpurun_out/strong_$N.err
fi
for f in gpurun_out/r02_bench_${N}gpu.json gpurun_out/r02_strong_10k_${N}gpu.json; do [ -s $f ] && python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d['n_gpus'], d['scaling'], round(d['value']), round(d['e2e']['value']), d.get('records_table_sha256','')[:16])" $f; done
