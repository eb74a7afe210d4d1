mkdir -p gpurun_out/r3o
for kr in 20 24 32 20 24; do echo -n "K20 kreg $kr: "; HVAE_TOPK_KREG=$kr python tools/bench_kernels.py --K 20 --reps 10 --only tc_score_topk 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print(round(d['ms'],4))
"; done
for kr in 8 12 16; do echo -n "K5 kreg $kr: "; HVAE_TOPK_KREG=$kr python tools/bench_kernels.py --K 5 --reps 10 --only tc_score_topk 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print(round(d['ms'],4))
"; done
