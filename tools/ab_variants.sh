#!/bin/bash
# A/B of differently compiled builds (tools/build_variant.sh) on the bench workload; run on the GPU box.
#   tools/ab_variants.sh OUT.jsonl NAME[:ENV=VAL,...] ...
out=$1; shift
: > $out
for spec in "$@"; do
  name=${spec%%:*}; envs=""
  [[ "$spec" == *:* ]] && envs=$(echo "${spec#*:}" | tr ',' ' ')
  lib=build/variants/lib_$name.so
  [[ "$name" == main ]] && lib=gym_lorenz_b200/libchaos_b200.so
  for rep in 1 2; do
    line=$(env CHAOS_B200_LIB=$lib $envs timeout -k 5 300 python bench.py --steps 300 --warmup 5 --no-e2e --no-cpu-baseline ${BENCH_ARGS} 2>/dev/null | tail -1)
    echo "{\"variant\": \"$spec\", \"rep\": $rep, \"line\": $line}" >> $out
  done
done
python - "$out" <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    d = json.loads(ln); l = d["line"]
    print(f'{d["variant"]:40s} rep{d["rep"]} {l["ms_per_step"]:.4f} ms  {l["value"]:.4e} steps/s  frac {l["roofline"]["frac"]:.4f}')
PY
