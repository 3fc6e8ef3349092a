for w in cfg3 cfg4; do
python bench.py --workload $w --steps 5 --warmup 3 2>gpurun_out/err_$w.txt | tail -n 1 > gpurun_out/bench_$w.json
python - $w <<'PY'
import json,sys
w=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/bench_{w}.json'))
    st=d['roofline']['stage_ms_per_step']
    print(w, 'ms/step %.3f'%d['ms_per_step'], 'planes/s %.3e'%d['value'], 'bones/s %.1f'%d['bones_per_sec'], 'segs', d['segments_per_step_per_gpu'])
    print('  stages', {k: round(v,3) for k,v in st.items()})
    print('  roofline', d['roofline']['kernel'], round(d['roofline']['frac'],4), 'pipeline', round(d['roofline']['pipeline']['frac'],4))
    print('  e2e', d['e2e']); print('  cpu', d['cpu_baseline'])
except Exception as e:
    print(w, 'FAILED', e); print(open(f'gpurun_out/err_{w}.txt').read()[-1500:])
PY
done
