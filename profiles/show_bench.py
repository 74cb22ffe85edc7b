import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"].get("inflight")); print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"launches/step",d["gpu_launches"]/d["steps"], "ms/step", round(d["ms_per_step"],3), "latency ms", round(d["latency"]["ms_per_keyframe"],3), "eager", round(d["latency"].get("ms_per_keyframe_eager",0),3))
print({k:round(v,3) for k,v in d["latency"]["stages_ms"].items()})
for k in d["kernels"]: print("  %-34s %8.1f us/step  n=%5.1f  %7.1f us/launch  %7.1f GB/s  %5.1f%%"%(k["kernel"],1000*k["ms_per_step"],k["launches_per_step"],k["us_per_launch"],k.get("achieved_gbs",0),100*k.get("frac_of_hbm_peak",0)))
if d.get("roofline"): print(d["roofline"])
if d.get("cpu_baseline"): print(d["cpu_baseline"])
