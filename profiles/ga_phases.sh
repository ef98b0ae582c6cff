#!/bin/bash
# Per-phase clock64 totals of ga_persist_kernel (a build with -DAHX_GA_TIMING kept as ab/libahx_T.so).
cd /root/repo
cp allhic_b200/lib/libahx.so /tmp/libahx_keep.so
cp ab/libahx_T.so allhic_b200/lib/libahx.so
for args in "config2 2000 --npop 100" "config2 2000" "config3 300"; do echo "== $args"; timeout 100 python profiles/prof_ga.py $args 2>&1 | tail -3; done
cp /tmp/libahx_keep.so allhic_b200/lib/libahx.so
