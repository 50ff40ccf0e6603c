python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "i8 or symmetric or batched or groups" 2>&1 | tail -1
for s in 8 1 32; do
  python bench.py --no-cpu-baseline --groups 1 --steps 60 --sessions $s > gpurun_out/sw.json 2>gpurun_out/sw.err || tail -3 gpurun_out/sw.err
  python -c "
import json; d=json.load(open('gpurun_out/sw.json')); print('S=$s', round(d['value']), round(d['roofline']['frac'],3), d['roofline']['kernels_us'])"
done
