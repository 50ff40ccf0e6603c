timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/rl8_gputest.txt 2>&1; echo "rc=$?" >> gpurun_out/rl8_gputest.txt
tail -4 gpurun_out/rl8_gputest.txt
B="timeout 300 python bench.py --no-cpu-baseline --no-adapter --steps 200"
$B > gpurun_out/rl8_g1.json 2> gpurun_out/rl8_g1.err
$B --steps 20 --warmup 5 > gpurun_out/rl8_g1_drv.json 2>/dev/null
$B --sessions 4 > gpurun_out/rl8_s4g1.json 2>/dev/null
$B --sessions 32 > gpurun_out/rl8_s32g1.json 2>/dev/null
$B --config C2 > gpurun_out/rl8_c2g1.json 2>/dev/null
REKF_TIMELINE=1 timeout 200 python scripts/timeline.py 1 8 > gpurun_out/rl8_timeline_g1_s8.txt 2>&1
REKF_TIMELINE=1 timeout 200 python scripts/timeline.py 1 1 > gpurun_out/rl8_timeline_g1_s1.txt 2>&1
for f in gpurun_out/rl8_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'blocking', round(d['e2e']['blocking']['value']), 'single', round(d['single_session']['value']), 'launches', d['gpu_launches'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -12 gpurun_out/rl8_timeline_g1_s8.txt; tail -12 gpurun_out/rl8_timeline_g1_s1.txt
