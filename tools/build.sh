#!/bin/bash
# rebuild libb2r.so (sm_100a) and print errors / register use of the hot kernels
cd "$(dirname "$0")/.." && python -c "
import __graft_entry__ as g
g.build_cuda(force=True)" 2>&1 | tail -3
grep -E "error" cpp-raytracer-rasterizer_b200/lib/build.log | head -5
grep -A2 -E "rt_trace_shade_kernelILb1ELb1ELb0|ras_small_kernel|ras_shade_kernel" cpp-raytracer-rasterizer_b200/lib/build.log | grep -E "registers|spill"
