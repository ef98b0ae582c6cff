#!/bin/bash
# Same-box A/B of two builds of libahx.so (boxes differ by ~1-2 %): ab/libahx_<name>.so are copied over
# allhic_b200/lib/libahx.so in turn and timed with prof_target.py / prof_ga.py, three rounds.
#   gpurun -- 'bash profiles/ab_compare.sh old new'
cd /root/repo
CFG=${CFG:-config3}; GENS=${GENS:-300}
cp allhic_b200/lib/libahx.so /tmp/libahx_keep.so
for round in 1 2 3; do
  for v in "$@"; do
    cp ab/libahx_$v.so allhic_b200/lib/libahx.so
    e=$(timeout 100 python profiles/prof_target.py $CFG 2>&1 | grep eval_m)
    g=$(timeout 60 python profiles/prof_ga.py $CFG $GENS 2>&1 | grep "GA ")
    echo "$round $v | $e | $g"
  done
done
cp /tmp/libahx_keep.so allhic_b200/lib/libahx.so
