"""In-sequence kernel timeline of one graphed SDXL UNet forward (BASELINE config 2) from CUPTI activity records
(torch.profiler): per kernel family the summed duration, the summed exposed time (the part of a kernel that does not
overlap its predecessor -- under programmatic dependent launch a kernel starts before the previous one ends) and the
summed idle gaps in front of it.  Complements tools/quick_bench.py (same-family replays) and the ncu launch list
(serialised, cold): this is what the step really spends, launch by launch.

    python tools/timeline_probe.py [out.json]
"""
import collections
import json
import os
import re
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stabletriton_b200 as st  # noqa: E402
from stabletriton_b200 import UNetConfig, synth  # noqa: E402


def main():
    cfg = UNetConfig.sdxl()
    model = synth.build_unet(cfg, seed=7)
    compiled = st.compile(model, cuda_graph=False)
    inp = synth.synth_inputs(2, 128, cfg, device="cuda", dtype=torch.bfloat16)
    s = torch.cuda.Stream()
    with torch.no_grad(), torch.cuda.stream(s):
        compiled(**inp)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            compiled(**inp)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    reps = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            g.replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    evs.sort(key=lambda e: e.time_range.start)
    if not evs:
        print("no CUDA activity records")
        return
    fam = collections.OrderedDict()
    prev_end = evs[0].time_range.start
    t_first, t_last = evs[0].time_range.start, max(e.time_range.end for e in evs)
    for e in evs:
        name = re.sub(r"[<(].*", "", e.name).replace("void ", "").replace("st::", "")
        d = fam.setdefault(name, {"launches": 0, "duration_us": 0.0, "exposed_us": 0.0, "gap_us": 0.0})
        st_, en = e.time_range.start, e.time_range.end
        d["launches"] += 1
        d["duration_us"] += en - st_
        d["exposed_us"] += max(0.0, en - max(st_, prev_end))
        d["gap_us"] += max(0.0, st_ - prev_end)
        prev_end = max(prev_end, en)
    total = (t_last - t_first) / reps
    print(f"{reps} replays, {len(evs) // reps} kernels each, {total / 1000.0:.3f} ms per replay (first kernel start -> last kernel end)")
    print(f"{'kernel':34s} {'launches':>8s} {'duration ms':>12s} {'exposed ms':>11s} {'gaps ms':>9s}")
    out = {"ms_per_replay": total / 1000.0, "families": {}}
    for name, d in sorted(fam.items(), key=lambda kv: -kv[1]["exposed_us"]):
        row = {k: (v / reps if k != "launches" else v // reps) for k, v in d.items()}
        out["families"][name] = row
        print(f"{name[:34]:34s} {row['launches']:8d} {row['duration_us'] / 1000.0:12.3f} {row['exposed_us'] / 1000.0:11.3f} "
              f"{row['gap_us'] / 1000.0:9.3f}")
    # the same per kernel instantiation (template arguments kept) and duration class, one replay
    per = len(evs) // reps
    one = evs[:per]
    inst = collections.OrderedDict()
    prev_end = one[0].time_range.start
    for e in one:
        name = e.name.replace("void ", "").replace("st::", "")
        name = re.sub(r"\(.*", "", name)
        st_, en = e.time_range.start, e.time_range.end
        exposed = max(0.0, en - max(st_, prev_end))
        prev_end = max(prev_end, en)
        bucket = 1 << max(0, int(exposed).bit_length() - 1)  # power-of-two class of the exposed time (us)
        d = inst.setdefault((name, bucket), {"launches": 0, "exposed_us": 0.0})
        d["launches"] += 1
        d["exposed_us"] += exposed
    print("\nper instantiation and exposed-time class (one replay):")
    rows = sorted(inst.items(), key=lambda kv: -kv[1]["exposed_us"])[:40]
    for (name, bucket), d in rows:
        print(f"  {name[:70]:70s} ~{bucket:4d} us x {d['launches']:4d}  {d['exposed_us'] / 1000.0:8.3f} ms")
    out["instantiations"] = [{"kernel": n, "exposed_class_us": b, **d} for (n, b), d in rows]
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
