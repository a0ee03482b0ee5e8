#!/bin/bash
# tools/sass_opcodes.sh -- per-kernel count of the SASS mnemonics that prove the Blackwell-native path (tcgen05 = UTC*MMA,
# TMEM = LDTM / STTM, TMA = UTMALDG / UBLKCP, mbarrier = SYNCS) in the shipped library.  Runs in the build container
# (cuobjdump needs no GPU).  Output: profiles/sass_opcodes.txt
SO=lowbit_quant_fa2_paddle_b200/liblowbit_fa_b200.so
{
echo "# cuobjdump -sass $SO  ($(date -u +%Y-%m-%dT%H:%MZ), $(sha256sum $SO | cut -c1-16))"
echo "# kernel | UTCIMMA (kind::i8) | UTCHMMA (kind::f16) | UTCQMMA (kind::f8f6f4) | LDTM | STTM | UTMALDG | UTCBAR (tcgen05.commit) | SYNCS (mbarrier) | MUFU.EX2 | HMMA (legacy mma.sync)"
cuobjdump -sass $SO | awk '
  /Function : /{ if (name != "") print_row(); name=$3; for (k in c) delete c[k]; next }
  { for (i=1;i<=NF;i++) { op=$i; if (op ~ /^UTCIMMA/) c["i"]++; else if (op ~ /^UTCHMMA/) c["h"]++; else if (op ~ /^UTCQMMA/) c["q"]++;
      else if (op ~ /^LDTM/) c["l"]++; else if (op ~ /^STTM/) c["s"]++; else if (op ~ /^UTMALDG/) c["t"]++; else if (op ~ /^UTCBAR/) c["b"]++;
      else if (op ~ /^SYNCS/) c["y"]++; else if (op ~ /^MUFU.EX2/) c["m"]++; else if (op ~ /^HMMA/) c["x"]++; } }
  function print_row() { printf "%s | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d\n", name, c["i"], c["h"], c["q"], c["l"], c["s"], c["t"], c["b"], c["y"], c["m"], c["x"] }
  END { if (name != "") print_row() }' | c++filt | sed 's/CUtensorMap_st/TM/g; s/lowbit:://g' | sort
} > profiles/sass_opcodes.txt
grep -c "|" profiles/sass_opcodes.txt
