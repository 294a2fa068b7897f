#!/bin/bash
# Headline grid (bench.py, 3 timed passes) over library builds in scratch/ab/ next to the shipped one
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --traffic-capture none"
for lib in optionslab_b200/libb200mc.so scratch/ab/lib_*.so; do
  n=$(basename $lib .so)
  B200MC_LIB=$PWD/$lib $B > gpurun_out/ab_$n.json 2>gpurun_out/ab_$n.err || tail -3 gpurun_out/ab_$n.err
  python - "$n" gpurun_out/ab_$n.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.4e" % d["value"], "ms/step %.2f" % d["ms_per_step"], "asian", d.get("asian_grid", {}).get("value"), "clk", d["clocks"]["sm_mhz"])
PY
done
