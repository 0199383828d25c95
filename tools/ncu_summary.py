"""Summarises ncu outputs into the text files kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv "<command that was profiled>" > profiles/launches_rN_summary.txt
  python tools/ncu_summary.py full rep1.ncu-rep[:frames_per_launch] [rep2.ncu-rep[:frames] ...] --traffic profiles/ncu_traffic_rN.json > profiles/ncu_full_rN_summary.txt

`launches`: the CSV written by `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...`.
`full`: reports of `ncu --set full --clock-control none --import-source on`; read through `ncu -i ... --page raw --csv`.
The traffic file holds dram__bytes_read.sum + dram__bytes_write.sum per launch PER FRAME for each kernel (bench.py reads it
for roofline.traffic).
"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict

PEAKS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}.get(unit, 1.0)


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").strip()


def launches(path, command):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        k = short(r[ki]); tot[k] += to_us(r[vi], r[ui]); cnt[k] += 1
    total = sum(tot.values())
    print("# ncu launch list summary (%s)" % os.path.basename(path))
    print("# command: %s" % command)
    print("# %d launches; per-launch times are cold-cache and serialised (ncu runs one kernel at a time): compare SHARES, not absolutes" % sum(cnt.values()))
    print("%-28s %8s %14s %8s" % ("kernel", "launches", "total_us", "share"))
    for k in sorted(tot, key=lambda k: -tot[k]):
        print("%-28s %8d %14.1f %7.2f%%" % (k, cnt[k], tot[k], 100 * tot[k] / total))


WANT = OrderedDict([
    ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct_of_peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"), ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("smsp__inst_executed.sum", "inst_executed"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "cycles_per_issued_inst"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
])


def full(specs, traffic_path):
    peak = None
    try:
        peak = float(json.load(open(PEAKS))["hbm_gbs"])
    except Exception:
        pass
    print("# ncu --set full --clock-control none --import-source on (B200, sm_100a); reports are scratch (gpurun_out/), this summary is the record")
    print("# dram_traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch; GB/s against MEASURED_PEAKS.json hbm_gbs = %s" % peak)
    traffic = {}
    for spec in specs:
        path, _, fpl = spec.partition(":")
        fpl = int(fpl) if fpl else 1
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print("\n## %s (%d frame%s per launch)" % (os.path.basename(path), fpl, "" if fpl == 1 else "s"))
        seen = defaultdict(int)
        for r in rows[2:]:
            d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
            k = short(d["Kernel Name"]); seen[k] += 1
            if seen[k] > 2:
                continue
            print("\n%s  (launch %d of this kernel in the report)" % (k, seen[k]))
            for m, label in WANT.items():
                if m in d and d[m] != "":
                    print("    %-22s %s %s" % (label, d[m], u.get(m, "")))
            try:
                unit_scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tr = float(d["dram__bytes_read.sum"].replace(",", "")) * unit_scale[u["dram__bytes_read.sum"]] + float(d["dram__bytes_write.sum"].replace(",", "")) * unit_scale[u["dram__bytes_write.sum"]]
                dur_us = to_us(d["gpu__time_duration.sum"], u["gpu__time_duration.sum"])
                gbs = tr / dur_us / 1e3
                print("    %-22s %.2f MB per launch (%.2f MB per frame) -> %.1f GB/s%s" % ("dram_traffic", tr / 1e6, tr / fpl / 1e6, gbs, (" = %.4f of measured peak" % (gbs / peak)) if peak else ""))
                if seen[k] == 1:
                    traffic[k] = int(tr / fpl)
            except Exception as e:                        # a metric missing from the report
                print("    dram_traffic           n/a (%s)" % e)
    if traffic_path:
        traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch PER FRAME, first captured launch of each kernel (tools/ncu_summary.py full ...)", **traffic}
        json.dump(traffic, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "?")
    else:
        args = sys.argv[2:]
        tp = None
        if "--traffic" in args:
            i = args.index("--traffic"); tp = args[i + 1]; args = args[:i] + args[i + 2:]
        full(args, tp)
