#!/bin/bash
# times the fused loss kernel for each DVF_LOSS_VARIANT (tuning aid; results land in gpurun_out/)
for v in ${VARIANTS:-0 1 2 3 4 5}; do
  DVF_LOSS_VARIANT=$v python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --roofline-ms 300 --prewarm-ms 100 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant $v kernel_us %.2f frac %.3f value %.3e' % (d['roofline']['kernel_us'], d['roofline']['frac'], d['value']))"
done
