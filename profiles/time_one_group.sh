# Times one config4 group through the host binary: full optimize (twice) and --skipGA.
cd /root/repo
python - <<'P'
import sys, os
sys.path.insert(0, ".")
from allhic_b200 import synth
c = synth.config4_groups()[0]
ids, clm, _ = synth.materialize(c["n"], seed=c["seed"], outdir="/tmp/c4")
print(ids, clm, os.path.getsize(clm))
P
F=/tmp/c4/n789_s4204_seed4204
for i in 1 2; do echo "full:"; ( time ./allhic_b200/bin/allhic-b200 optimize $F.ids $F.clm > /tmp/o.out 2> /tmp/o.err ) 2>&1 | grep -E "real|user|sys" | tr "\n" " "; echo; grep -c "Current iteration" /tmp/o.out; done
echo "skipGA:"; ( time ./allhic_b200/bin/allhic-b200 optimize $F.ids $F.clm --skipGA > /tmp/o.out 2> /tmp/o.err ) 2>&1 | grep -E "real|user|sys" | tr "\n" " "; echo
grep -E "GA initialized|FLIPONE:|Success" /tmp/o.err | head
