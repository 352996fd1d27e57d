#!/bin/bash
# usage: profiles/sweep_shapes.sh lib... — forward/rank times per (shape, batch) for the default build and each library
for cfg in "base 8" "base 1" "native 8" "stress 2" "stress 1"; do set -- $cfg; echo "== $1 B=$2"; SHAPE=$1 BATCH=$2 profiles/sweep_libs.sh "${@:3}" ${LIBS} 2>&1 | sed "s#/root/repo/fusionocc_b200/lib/##; s# bwdplan.*##"; done
