#!/bin/bash
# Dev tool: A/B of library variants (scripts/build_variant.sh) with scripts/order_sweep.py, one process per variant.
# usage: k1_variants.sh "<tag> <tag> ..." <case> [<case> ...]      (case = N:elements_per_side)
TAGS="$1"; shift
for tag in $TAGS; do
  echo "=== $tag"
  SEMB_LIB="sempy_b200/libsemb200_$tag.so" timeout 300 python scripts/order_sweep.py "$@" 2>&1 | grep -v Warning
done
