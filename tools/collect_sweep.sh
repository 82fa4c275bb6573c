# tuning only: whole-step effect of the collect kernel's grid size (CTAs per SM)
for r in 1 2 3 1 2 3; do DRG_COLLECT_PER_SM=$r timeout 100 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('per_sm=$r', round(d['value'],1), d['kernel_ms_per_step']['topk_collect'])"; done
