#!/usr/bin/env python3
"""DRAM traffic of ONE full sampler launch of the bench workload, measured (not extrapolated):

    python bench.py --steps 1 --warmup 3 --no-cpu --e2e-reps 1 > gpurun_out/plain.log 2>&1 &&
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \\
        -k regex:nuts_kernel --csv --log-file gpurun_out/r02_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu --e2e-reps 1
    python tools/ncu_traffic.py gpurun_out/r02_traffic.csv > profiles/r02_traffic.json

The launch list has the 3 short warm-up launches, the timed launch and the e2e launch; the timed one (4th) is reported.
bench.py puts `dram_bytes_per_launch` into roofline.traffic when its workload is the captured one."""
import csv
import json
import sys


def main(path, combos=1332, it=1150, wu=1000, chains=4):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    iid, iname, imetric, iunit, ival = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    per = {}
    for r in rows[1:]:
        v = float(r[ival].replace(",", ""))
        u = r[iunit]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
                 "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}.get(u, 1.0)
        per.setdefault(int(r[iid]), {"kernel": r[iname][:60]})[r[imetric]] = v * scale
    ids = sorted(per)
    launches = [per[i] for i in ids]
    full = [l for l in launches if l.get("gpu__time_duration.sum", 0) > 0.5 * max(x.get("gpu__time_duration.sum", 0) for x in launches)]
    t = full[0]
    out = {"combos": combos, "iter": it, "warmup_iter": wu, "chains": chains,
           "dram_bytes_read_per_launch": t["dram__bytes_read.sum"], "dram_bytes_write_per_launch": t["dram__bytes_write.sum"],
           "dram_bytes_per_launch": t["dram__bytes_read.sum"] + t["dram__bytes_write.sum"],
           "launch_seconds_under_ncu": t.get("gpu__time_duration.sum"), "n_launches_captured": len(launches),
           "kernel": t["kernel"],
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on `python bench.py --steps 1 "
                     "--warmup 3 --no-cpu --e2e-reps 1` (profiles/r02_traffic.csv), the timed full launch"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
