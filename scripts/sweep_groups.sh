for st in 0 1 2; do for rs in 6 12 24; do
  echo "stagger=$st reserve=$rs"; REKF_STAGGER=$st python bench.py --no-cpu-baseline --no-adapter --steps 100 --reserve-sms $rs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done; done
