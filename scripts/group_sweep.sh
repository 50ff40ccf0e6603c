for hg in 0 1 0 1; do
  REKF_HOST_GRAPHS=$hg python bench.py --no-cpu-baseline --groups 2 > gpurun_out/gs.json 2>gpurun_out/gs.err || tail -3 gpurun_out/gs.err
  python -c "
import json; d=json.load(open('gpurun_out/gs.json')); print('host_graphs=$hg', round(d['value']), round(d['e2e']['value']), round(d['e2e']['blocking']['value']))"
done
