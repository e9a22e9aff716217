# library-variant sweep: GCIS_LIB=build/libgcis_<name>.so python bench.py ... for every name given
for n in "$@"; do echo -n "$n: "; GCIS_LIB=$PWD/build/libgcis_$n.so timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-ref-metrics 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, d['dataset_scores']['recall'], d['dataset_scores']['underseg'])"; done
