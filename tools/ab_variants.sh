#!/bin/bash
# A/B of library variants (tools/build_variant.py): ab_variants.sh "B mode T iters" base poll1 ...
args="$1"; shift
for v in "$@"; do
  if [ "$v" = base ]; then lib=""; else lib="variants/libclipcap_$v.so"; fi
  echo "== $v ($args)"
  CCB_LIB=$lib timeout 300 python tools/quick_xl.py $args 2>&1 | grep -E "iter [2-9]|checksum|Error|error" 
done
