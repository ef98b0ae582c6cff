#!/bin/bash
# optimize-many on two copies of the reference fixture (one process, one device context), against a
# plain `optimize` run; then 12 config4 groups through the farm with and without persistent workers.
cd /root/repo
OUT=gpurun_out; mkdir -p $OUT
rm -rf /tmp/fx && mkdir -p /tmp/fx/a /tmp/fx/b /tmp/fx/c
for d in a b c; do cp tests/golden/test.ids tests/golden/test.clm /tmp/fx/$d/; done
printf "/tmp/fx/a/test.ids /tmp/fx/a/test.clm\n/tmp/fx/b/test.ids /tmp/fx/b/test.clm\n" > /tmp/fx/list
( time ./allhic_b200/bin/allhic-b200 optimize-many /tmp/fx/list > $OUT/many.out 2> $OUT/many.err ) 2> $OUT/many.time
./allhic_b200/bin/allhic-b200 optimize /tmp/fx/c/test.ids /tmp/fx/c/test.clm > $OUT/single.out 2>/dev/null
{ cmp /tmp/fx/a/test.tour /tmp/fx/b/test.tour && echo "a == b"; cmp /tmp/fx/a/test.tour /tmp/fx/c/test.tour && echo "a == single"; } > $OUT/many.cmp 2>&1
N=${1:-12}
timeout 400 python profiles/run_config4.py $N --cpu 0 > $OUT/c4_persist.json 2> $OUT/c4_persist.err
timeout 400 python profiles/run_config4.py $N --cpu 0 --per-group-process > $OUT/c4_pergroup.json 2> $OUT/c4_pergroup.err
echo done
