import json,sys
d=json.load(open(sys.argv[1]))
for k,v in d.items():
    print(k, {kk:(round(vv["ms"],4), round(vv["frac_of_measured_hbm_peak"],3)) if isinstance(vv,dict) else round(vv,3) for kk,vv in v.items() if kk not in ("points","stencil_planes")})
