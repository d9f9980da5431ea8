for t in 128 256 512; do
  echo "== LBA_CAM_THREADS=$t"
  LBA_CAM_THREADS=$t python bench.py --no-cpu-baseline --steps 20 --warmup 3 | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))
print(d['roofline']['time_shares'])
print({k:v['avg_us'] for k,v in d['roofline']['per_kernel'].items()})
"
done
