#!/bin/bash
# Same-box A/B of builds of libahx.so on config4 groups through the farm (one GPU, six groups side by side):
#   gpurun -- 'bash profiles/ab_config4.sh 6 old new'
cd /root/repo
NG=$1; shift
cp allhic_b200/lib/libahx.so /tmp/libahx_keep.so
for round in 1 2; do
  for v in "$@"; do
    cp ab/libahx_$v.so allhic_b200/lib/libahx.so
    python profiles/run_config4.py $NG --cpu 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$round $v wall %.2f s  per group %s  ok=%s' % (d['wall_s'], d['seconds_per_group'], d['all_ok']))"
  done
done
cp /tmp/libahx_keep.so allhic_b200/lib/libahx.so
