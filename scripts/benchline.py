import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],1), d["alm_iters"], round(d["mask_fraction"],5), {k.split("_kernel")[0]:(round(v["ms"],3)) for k,v in d["kernels"].items()}, d.get("breakdown_ms"), d.get("iters_enqueued"))
