#!/bin/bash
# developer helper (GPU box): resident cycle time and sweep time per robot for library variants
#   tools/variant_dev.sh NAME...   (variants built by tools/build_variant.sh; "main" = the in-tree library)
for v in "$@"; do
  if [ "$v" = main ]; then unset KOMPASS_B200_LIB; else export KOMPASS_B200_LIB=$PWD/kompass-core_b200/lib/variants/libkompass_b200_$v.so; fi
  echo "=== variant $v"
  for d in dense_cluster_on_path friendly_ring survey_c2 clutter_in_reach; do
    python tools/family_dev.py --replay $d 400 2>&1 | tail -1
  done
  python tools/sweep_prof_dev.py dense_cluster_on_path 64 2>&1 | tail -1
  python tools/sweep_prof_dev.py friendly_ring 64 2>&1 | tail -1
done
