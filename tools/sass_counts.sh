#!/bin/bash
# Count the Blackwell-specific SASS mnemonics of the product library (tcgen05.mma / ld / st / commit, TMA load / store).
so=${1:-head-pose-estimation-model_b200/libhpose.so}
cuobjdump -sass "$so" > /tmp/hpose.sass
for m in UTCHMMA UTCQMMA LDTM STTM UTCBAR UTMALDG UTMASTG UBLKCP SYNCS; do printf "%-8s %s\n" $m $(grep -c "$m" /tmp/hpose.sass); done
