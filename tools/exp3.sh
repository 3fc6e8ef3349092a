for mb in "1 1" "10 1" "1 10" "10 10"; do
set -- $mb
SHB_NVCC_EXTRA="-DSHB_RESAMPLE_MIN_BLOCKS=$1 -DSHB_STITCH_MIN_BLOCKS=$2" python -m shoulder_b200.build --force > /dev/null
echo "min blocks resample=$1 stitch=$2"; python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=d['roofline']['stage_ms_per_step']; print('  ms/step %.3f stitch %.3f resample %.3f'%(d['ms_per_step'],st['stitch'],st['resample']))"
done
