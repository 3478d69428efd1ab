#!/bin/bash
# A/B of kernel variants on the bench workload (run under gpurun): prints step / response / assembly times.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in ${VARIANTS:-0 1 2}; do
echo "== passb_variant=$w"
python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e --passb-wsmem $w 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('step ms %.4f  resp ms %.4f (frac %.3f)  asm ms %.4f (frac %.3f)'%(d['ms_per_step'],d['roofline']['ms'],d['roofline']['frac'],d['roofline']['assembly']['ms'],d['roofline']['assembly']['frac']))"
done
