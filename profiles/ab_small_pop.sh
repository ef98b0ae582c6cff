cd /root/repo
cp allhic_b200/lib/libahx.so /tmp/libahx_keep.so
for round in 1 2 3; do for v in head small; do cp ab/libahx_$v.so allhic_b200/lib/libahx.so; g=$(timeout 60 python profiles/prof_ga.py config2 4000 --npop 100 2>&1 | grep "GA "); echo "$round $v | $g"; done; done
cp /tmp/libahx_keep.so allhic_b200/lib/libahx.so
