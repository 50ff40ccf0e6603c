REKF_TIMELINE=1 REKF_STAGGER=2 python scripts/timeline.py 2 > gpurun_out/timeline_g2.txt 2>&1; tail -34 gpurun_out/timeline_g2.txt
for m in 0 1 2; do
  REKF_STAGGER=$m python bench.py --no-cpu-baseline --groups 2 --steps 100 > gpurun_out/gs.json 2>gpurun_out/gs.err || tail -3 gpurun_out/gs.err
  python -c "
import json; d=json.load(open('gpurun_out/gs.json')); print('stagger=$m', round(d['value']), round(d['e2e']['value']), round(d['e2e']['blocking']['value']))"
done
