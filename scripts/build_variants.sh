#!/bin/bash
# A/B builds of libgo1mpc.so with different -D settings: scripts/build_variants.sh name1 "-DX=1" name2 "-DY=2" ...
# -> quadrupedal_loco_b200/build/variants/libgo1mpc_<name>.so (select with GO1MPC_LIB)
set -e
cd "$(dirname "$0")/.."
mkdir -p quadrupedal_loco_b200/build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  GO1MPC_NVCC_EXTRA="$flags" python -c "from quadrupedal_loco_b200 import _build; _build.build(force=True)"
  cp quadrupedal_loco_b200/libgo1mpc.so quadrupedal_loco_b200/build/variants/libgo1mpc_$name.so
done
python -c "from quadrupedal_loco_b200 import _build; _build.build(force=True)"
