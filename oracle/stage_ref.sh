#!/bin/bash
# TEST / MEASUREMENT INFRASTRUCTURE.  Stages the reference's own Triton kernel files (the ones oracle/ref_triton.py
# loads) from /root/reference into oracle/_ref/reference/, a git-ignored directory that travels to the GPU box with
# gpurun (like baseline/_ref would for a pip-installable reference).  Nothing here is committed; nothing in tests/,
# bench.py or smoke() depends on it.  Used only by tools/ref_on_b200.py to time the UNMODIFIED reference kernels on the
# same B200 and to compare outputs at full size.
set -e
SRC=${1:-/root/reference}
DST=$(dirname "$0")/_ref/reference
[ -d "$SRC/src/triton" ] || { echo "no reference at $SRC"; exit 0; }
mkdir -p "$DST/src/triton"
for f in quant_per_block.py attn_qk_int8_per_block.py attn_qk_int8_per_block_causal.py; do
  cp "$SRC/src/triton/$f" "$DST/src/triton/$f"
done
echo "staged $(ls $DST/src/triton | wc -l) reference kernel files under $DST"
