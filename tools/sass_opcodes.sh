#!/bin/bash
# SASS evidence of the Blackwell-native paths: opcode counts per kernel object of the built library.
# Usage: bash tools/sass_opcodes.sh > profiles/rNN_sass_opcodes.txt
L=estorch_b200/lib
echo "# cuobjdump -sass of the objects in $L (built by estorch_b200/csrc/build.sh, -gencode arch=compute_100a,code=sm_100a)"
for o in estk_eval_mlp_f16 estk_eval_mlp_tc estk_rank_grad estk_eval_mlp estk_eval_conv estk_noise estk_misc; do
  f=$L/$o.o; [ -f $f ] || continue
  echo "== $o.o"
  cuobjdump -sass $f > /tmp/_sass.txt
  for op in UTCHMMA UTCBAR LDTM STTM UTMALDG UBLKCP "SYNCS" "USETMAXREG" "LDG.E.*128" "LDG.E.*256" "LDS.128" "STS.128" "FENCE.VIEW.ASYNC" UCGABAR " HMMA\\." "FFMA" "FFMA2" "HADD2.F32" "F2FP" "MEMBAR.*SYS" "ATOMG\|ATOM\." "LDS.64"; do
    n=$(grep -cE "$op" /tmp/_sass.txt); [ "$n" != "0" ] && printf "  %-18s %6d\n" "$op" "$n"
  done
  grep -E "Function :" /tmp/_sass.txt | sed 's/.*Function : /  kernel /' | cut -c1-150
done
