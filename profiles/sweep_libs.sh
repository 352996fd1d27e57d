#!/bin/bash
# usage: profiles/sweep_libs.sh [lib ...]  — time the four phases for each library build (FUSIONOCC_B200_LIB)
cd "$(dirname "$0")/.."
python profiles/fwd_variant_probe.py
for l in "$@"; do FUSIONOCC_B200_LIB=$PWD/$l python profiles/fwd_variant_probe.py; done
