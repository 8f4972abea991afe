#!/bin/bash
# run the aggregation probe on every experiment build present (libgta_b200_<tag>.so)
for lib in gta_graph_tensor_acclelrator_for_general_gnn_b200/libgta_b200*.so; do
  tag=$(basename $lib .so); tag=${tag#libgta_b200}; tag=${tag#_}
  echo "=== variant '${tag:-default}'"
  GTA_LIB_TAG=$tag python tools/agg_probe.py "$@" 2>&1 | tail -4
done
