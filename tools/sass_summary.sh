#!/bin/bash
# per-kernel SASS mnemonic counts of the in-tree library -> profiles/<round>/sass_summary.txt (no GPU needed)
cd "$(dirname "$0")/.."
OUT=${1:-profiles/r02/sass_summary.txt}
{
echo "SASS mnemonics per kernel of pfann_b200/csrc/libpfann_b200.so (cuobjdump -sass)"
echo "UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UBLKCP = TMA (tensor / bulk), FP32x2 = FFMA2+FADD2 (packed fp32),"
echo "HFMA2RELU = fused bf16 affine+ReLU, HMMA = legacy mma.sync (none expected), RED/ATOMG = global reductions / atomics"
echo
printf "%-8s %-6s %-6s %-8s %-7s %-7s %-9s %-5s %-9s %s\n" "UTC*MMA" "LDTM" "STTM" "UTMALDG" "UBLKCP" "FP32x2" "HFMA2RELU" "HMMA" "RED/ATOMG" "kernel"
cuobjdump -sass pfann_b200/csrc/libpfann_b200.so | awk '
/Function : /{fn=substr($3,1,200); seen[fn]=1}
/UTC[A-Z]*MMA/{a[fn]++} /LDTM/{b[fn]++} /STTM/{s[fn]++} /UTMALDG/{c[fn]++} /UBLKCP/{d[fn]++} /FFMA2|FADD2/{e[fn]++} /HFMA2.*RELU/{f[fn]++} /[ \t]HMMA/{h[fn]++} /REDG|ATOMG/{r[fn]++}
END{for (k in seen) if (a[k]+b[k]+c[k]+d[k]+e[k]+r[k] > 0) printf "%-8d %-6d %-6d %-8d %-7d %-7d %-9d %-5d %-9d %s\n", a[k],b[k],s[k],c[k],d[k],e[k],f[k],h[k],r[k],k}' | c++filt | sed 's/(anonymous namespace):://g; s/CUtensorMap_st/TMap/g' | cut -c1-190 | sort -k10
} > "$OUT"
