#!/bin/bash
# developer sweep: K2 kernel time of one workload under engine option variants
# usage: tools/k2_sweep.sh <workload> <outfile> "opt1=v,opt2=v" "..." ...
wl=$1; out=$2; shift 2
: > "$out"
for v in "$@"; do
  args=""
  if [ "$v" != "base" ]; then for kv in ${v//,/ }; do args="$args --opt $kv"; done; fi
  line=$(BENCH_NO_SAMPLER=1 timeout 120 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $args 2>/tmp/k2_sweep.err | tail -1)
  echo "$v :: $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print("step %.3f ms  K2 %.3f ms  launches/step %.1f  frac %.3f  nulls %d" % (d["ms_per_step"], r["kernel_ms_per_step"], r["launches_per_step"], r["frac"] or 0, d["config"]["null_regions_rank0"]))' 2>&1)" >> "$out"
  grep "null2 phases" /tmp/k2_sweep.err | tail -4 >> "$out"
done
cat "$out"
