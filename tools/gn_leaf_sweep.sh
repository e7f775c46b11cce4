#!/bin/bash
# Tuning aid (GPU): GN throughput of the benchmark graph for several dissection leaf sizes.
for leaf in "$@"; do
  echo "leaf $leaf"
  CGM_PGO_LEAF=$leaf python tools/pgo_profile_run.py 6 32 2>&1 | tail -2 | cut -c1-400 | sed 's/np.int32//g; s/np.float64//g' | grep -o "ms per instance-iteration.*\|stage_ms.*"
done
