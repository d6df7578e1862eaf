for v in "" b32 b64; do
  if [ -n "$v" ]; then export GO1MPC_LIB=$PWD/quadrupedal_loco_b200/build/variants/libgo1mpc_$v.so; else unset GO1MPC_LIB; fi
  echo "variant ${v:-default}"
  python scripts/overlap_probe.py 2>&1 | grep "planner only" | sed -n '1p;3p;6p'
done
