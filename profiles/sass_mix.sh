#!/bin/bash
# Instruction mix of the consumer loop of eval_m_res_kernel<4,int> (between the first and last MUFU.RCP64H).
cd /root/repo
cuobjdump -sass allhic_b200/csrc/build/ahx_core.o > /tmp/core.sass
L0=$(grep -n "Function : _ZN3ahx17eval_m_res_kernelILi4ELb0ELb0ENS_9ResDirectILi4EiE" /tmp/core.sass | cut -d: -f1)
L1=$(awk -v s=$L0 'NR>s && /Function :/ {print NR; exit}' /tmp/core.sass)
sed -n "${L0},${L1}p" /tmp/core.sass | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" > /tmp/res4.sass
F=$(grep -n "MUFU.RCP64H" /tmp/res4.sass | head -1 | cut -d: -f1); La=$(grep -n "MUFU.RCP64H" /tmp/res4.sass | tail -1 | cut -d: -f1)
echo "MUFU count $(grep -c MUFU.RCP64H /tmp/res4.sass), lines $F..$La"
sed -n "$((F-${1:-60})),$((La+${2:-40}))p" /tmp/res4.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9] +//' | awk '{print $1}' | cut -d. -f1 | sort | uniq -c | sort -rn | head -24
