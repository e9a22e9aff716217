# launch-group size sweep (images per kernel launch): GCIS_GROUP=n python bench.py ...
for g in 25 40 50 67 100 200; do echo -n "group=$g: "; GCIS_GROUP=$g timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu --no-ref-metrics 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items()}, d['config'].get('images_per_launch'))"; done
