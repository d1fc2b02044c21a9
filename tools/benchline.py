import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d["roofline"]
print("%.1f M/s  step %.2f ms  fine %.2f ms coarse %.2f ms frac %.3f e2e %.1f" % (d["value"]/1e6, d["ms_per_step"], r["kernel_ms"], r["coarse_kernel_ms"], r["frac"], d["e2e"]["value"]/1e6))
