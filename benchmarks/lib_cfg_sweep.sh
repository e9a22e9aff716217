# secondary configurations per library variant: bash benchmarks/lib_cfg_sweep.sh "<configs>" <lib names...>
cfgs=$1; shift
for n in "$@"; do for c in $cfgs; do echo -n "$n $c: "; GCIS_LIB=$PWD/build/libgcis_$n.so timeout 300 python bench.py --config $c --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(round(d['value'],1), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, r['bound'], round(r['frac'],3), round(r['hbm_gbs']), round(r['fp32_tflops'],1))"; done; done
