for m in 2 0 2 0 2; do
  REKF_STAGGER=$m python bench.py --no-cpu-baseline --groups 2 > gpurun_out/gs.json 2>gpurun_out/gs.err || tail -3 gpurun_out/gs.err
  python -c "
import json; d=json.load(open('gpurun_out/gs.json')); print('stagger=$m', round(d['value']), round(d['e2e']['value']), round(d['e2e']['blocking']['value']), d['clocks'])"
done
