#!/bin/bash
# Tuning aid (GPU box): rebuild the library with one constant changed and time the benchmark graph.
# usage: tools/gn_variant_sweep.sh <file> <sed-expr> [<sed-expr> ...]   (each expr is one variant)
f=$1; shift
cp "$f" /tmp/variant_backup
for expr in "$@"; do
  cp /tmp/variant_backup "$f"
  sed -i "$expr" "$f"
  echo "== variant: $expr"
  make -s -C cg_mrslam_b200/csrc 2>&1 | grep -i "error" | head -5
  for b in 1 128; do
    python tools/pgo_profile_run.py 4 $b 2>&1 | grep -o "ms per instance-iteration [0-9.]*" | sed "s/^/batch $b: /"
  done
done
cp /tmp/variant_backup "$f"
make -s -C cg_mrslam_b200/csrc
