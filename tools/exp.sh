timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "passed|failed|FAILED|Error|assert |^E " | head -20
python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=d['roofline']['stage_ms_per_step']; print('  ms/step %.3f sum_stages %.3f'%(d['ms_per_step'],sum(st.values())), st, d['value'])"
