for mb in 1 11 12; do
SHB_NVCC_EXTRA="-DSHB_RESAMPLE_MIN_BLOCKS=$mb" python -m shoulder_b200.build --force > /dev/null
echo "min blocks $mb"; python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=d['roofline']['stage_ms_per_step']; print('  ms/step %.3f stitch %.3f resample %.3f'%(d['ms_per_step'],st['stitch'],st['resample']))"
done
