#!/bin/bash
# Round-end evidence on one B200 (gpurun -- bash scripts/final_evidence.sh): GPU tests, smoke, bench lines, ncu launch list and
# --set full capture, kernel timelines.  Everything lands in gpurun_out/; the files to keep are copied into profiles/ by hand.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/final_gputest.txt 2>&1; echo "rc=$?" >> $O/final_gputest.txt
tail -3 $O/final_gputest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/final_smoke.txt 2>&1; tail -2 $O/final_smoke.txt
timeout 900 python bench.py > $O/final_bench_c3.json 2> $O/final_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/final_bench_c3_drv.json 2>> $O/final_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/final_bench_ref.json 2>> $O/final_bench.err
Q="--no-cpu-baseline --no-adapter"
timeout 300 python bench.py $Q --config C2 > $O/final_bench_c2.json 2>> $O/final_bench.err
timeout 300 python bench.py $Q --config C4 --steps 40 > $O/final_bench_c4.json 2>> $O/final_bench.err
timeout 300 python bench.py $Q --sessions 32 > $O/final_bench_s32.json 2>> $O/final_bench.err
timeout 300 python bench.py $Q --sessions 1 > $O/final_bench_s1.json 2>> $O/final_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 400 --csv --log-file $O/r02b_launches.csv python bench.py --steps 20 --warmup 3 $Q > $O/r02b_ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -s 300 -c 16 -o $O/r02b_full -f python bench.py --steps 20 --warmup 3 $Q > $O/r02b_ncu_b.log 2>&1
REKF_TIMELINE=1 timeout 200 python scripts/timeline.py 1 8 > $O/r02b_timeline_s8.txt 2>&1
REKF_TIMELINE=1 timeout 200 python scripts/timeline.py 1 1 > $O/r02b_timeline_s1.txt 2>&1
for f in $O/final_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], {k: d.get(k) for k in ('impl','value','ms_per_step')}, 'e2e', d.get('e2e',{}).get('value'), 'single', (d.get('single_session') or {}).get('value'), 'frac', (d.get('roofline') or {}).get('frac'), 'parity', d.get('parity'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
cat $O/final_bench.err | tail -5
