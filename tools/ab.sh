#!/bin/bash
# A/B timing of variant builds (libb2r_<name>.so, see __graft_entry__.build_variant): config 4, sort-last default pipeline
L=$PWD/cpp-raytracer-rasterizer_b200/lib
for v in "$@"; do
  if [ "$v" = base ]; then lib=$L/libb2r.so; else lib=$L/libb2r_$v.so; fi
  B2R_LIB=$lib RAS_VARIANTS=${RAS_VARIANTS:-0} python tools/ras_time.py 2>&1 | grep "^ras variant" | sed "s/^/== $v: /"
done
