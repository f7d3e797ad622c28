import sys,json
tag=sys.argv[1]
d=json.loads(sys.stdin.read())
print(tag, round(d["ms_per_step"],3), d["config"]["regions"], d["config"]["null_regions"], round(d["roofline"]["kernel_ms_per_step"],3), d["roofline"].get("other_kernels_ms_per_step"), round(d["e2e"]["ms_per_step"],3))
