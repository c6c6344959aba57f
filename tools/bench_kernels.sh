#!/bin/bash
# Dev tool: short bench run, prints step time + per-kernel ms (usage: tools/bench_kernels.sh TAG [ENV=VAL ...])
tag=$1; shift
env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/$tag.log 2>&1
python - "$tag" <<EOF
import json, sys
l = [x for x in open("gpurun_out/%s.log" % sys.argv[1]) if x.startswith("{")]
if not l:
    print(open("gpurun_out/%s.log" % sys.argv[1]).read()[-2000:])
else:
    j = json.loads(l[-1])
    print(sys.argv[1], "step %.3f ms  e2e %.3f ms" % (j["ms_per_step"], j["e2e"]["ms_per_step"]))
    print("  ", {k: round(v["ms_per_step"], 3) for k, v in j["roofline"]["kernels"].items() if v["ms_per_step"] > 0.015})
EOF
