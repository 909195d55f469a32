for lib in v0 v1 v2 v0 v1 v2; do QB_LIB=$PWD/_ab/lib_$lib.so python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
for x in sys.stdin:
    if x.startswith('{'):
        d=json.loads(x); print('$lib', round(d['value']), d['kernel_ms_per_step'], d['logical_errors'])
"; done
