for ov in 0 1; do MICO_ATTN_TAIL_OVERLAP=$ov timeout 300 python bench.py --config vitg --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); f=d['roofline']['families']; print('vitg overlap=$ov', round(d['ms_per_step'],2), 'attn fwd', round(f['attention_fwd']['ms_per_step'],2), 'bwd', round(f['attention_bwd']['ms_per_step'],2))"; done
