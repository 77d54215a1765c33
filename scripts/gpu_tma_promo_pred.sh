cd /root/repo
for rep in 1 2; do for promo in 2 1; do export E3B_TMA_PROMO=$promo; echo promo=$promo; timeout 300 python bench.py --no-cpu-baseline --no-ref-gpu --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read()); p=b['predictor']; print('train %.4f pred %.4f e2e %.4f domfrac %.3f preddom %.3f' % (b['ms_per_step'], p['seconds_per_volume'], p['e2e']['seconds_per_volume'], b['roofline']['frac'], p['roofline']['frac']))"; done; done
