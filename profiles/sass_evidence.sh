#!/bin/bash
# SASS evidence of the built library (no GPU needed): which tensor / TMA mnemonics every kernel contains, and the inner
# k-loop of the benchmarked instantiation.   bash profiles/sass_evidence.sh > profiles/r02/sass_evidence.txt
SO=closedloopreachability.jl_b200/libclr_b200.so
echo "# cuobjdump -sass $SO  ($(date -u +%F), nvcc $(nvcc --version | grep -o 'release [0-9.]*'))"
echo "# embedded cubins:"; cuobjdump -lelf $SO | sed 's/^/#   /'
echo "# per kernel: instruction count, DMMA (mma.sync.m8n8k4.f64), HMMA, UTC*MMA (tcgen05), UTMALDG/UBLKCP (TMA), LDGSTS (cp.async), BAR, LDCU from the user constant bank c[0x3] (weights into uniform registers)"
cuobjdump -sass $SO | awk '
  /Function :/ { if (name != "") printf "%-110s inst %6d DMMA %5d HMMA %d UTCMMA %d TMA %d LDGSTS %d BAR %d LDCUc3 %d\n", name, n, d, h, u, t, l, b, c; name=$3; n=d=h=u=t=l=b=c=0; next }
  /^[ \t]+\/\*[0-9a-f]+\*\// { n++; if ($0 ~ /DMMA/) d++; if ($0 ~ /HMMA/) h++; if ($0 ~ /UTC[A-Z]*MMA/) u++; if ($0 ~ /UTMALDG|UTMASTG|UBLKCP/) t++; if ($0 ~ /LDGSTS/) l++; if ($0 ~ /BAR\./) b++; if ($0 ~ /LDCU.*c\[0x3\]/) c++ }
  END { printf "%-110s inst %6d DMMA %5d HMMA %d UTCMMA %d TMA %d LDGSTS %d BAR %d LDCUc3 %d\n", name, n, d, h, u, t, l, b, c }' | c++filt | sort
echo
echo "# inner k-loop of deepz_kernel<25, false, false, false, false> (the C4 bench instantiation): W fragments in registers (R.. second operand),"
echo "# B fragments LDS.64 from the column store, 4 independent accumulator chains per warp, one uniform exit test per k-step"
cuobjdump -sass -fun '_Z12deepz_kernelILi25ELb0ELb0ELb0ELb0EEvPK6DevNet10DeepzInput11DeepzOutputP12DevCallState10DeepzSched' $SO | grep -n "DMMA" | sed -n '26p' | cut -d: -f1 > /tmp/_l
L=$(cat /tmp/_l)
cuobjdump -sass -fun '_Z12deepz_kernelILi25ELb0ELb0ELb0ELb0EEvPK6DevNet10DeepzInput11DeepzOutputP12DevCallState10DeepzSched' $SO | sed -n "$((L-12)),$((L+40))p" | sed 's/ *\/\* 0x[0-9a-f]* \*\/$//'
echo
echo "# inner loop of controller_cw_kernel<4, 2, 4, 4, true> (Unicycle controller): LDCU.64 from the constant bank into uniform registers, DFMA with uniform operands"
cuobjdump -sass $SO | awk '/Function : /{f=0} /Function : .*controller_cw_kernelILi4ELi2ELi4ELi4ELb1E/{f=1} f{print}' | grep -E '^\s+/\*[0-9a-f]+\*/' | sed 's/ *\/\* 0x[0-9a-f]* \*\/$//' | awk '/LDCU.64 UR.*c\[0x3\]/{f=1} f{print; n++} n>=48{exit}'
