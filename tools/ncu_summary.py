"""Summarise an `ncu --set full` report (read here with `ncu -i X.ncu-rep --page raw --csv`) into a markdown table:
per launch kernel, grid, duration, DRAM read / write bytes, tensor-pipe active %, L2 hit %, registers.
Usage: ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv"""
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    body = [r for r in rows[2:] if len(r) == len(hdr)]      # row 1 holds the units
    units = dict(zip(hdr, rows[1]))

    def col(name):
        return hdr.index(name) if name in hdr else None
    want = {"name": "Kernel Name", "grid": "Grid Size", "time": "gpu__time_duration.sum",
            "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
            "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "l2hit": "lts__t_sector_hit_rate.pct", "regs": "launch__registers_per_thread",
            "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"}
    idx = {k: col(v) for k, v in want.items()}

    def scale(v, unit, target):
        v = float(v.replace(",", ""))
        f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
             "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(unit, 1.0)
        return v * f / target
    print("| kernel | grid | time us | dram read MB | dram write MB | tensor pipe % | dram % of peak | L2 hit % | regs |")
    print("|---|---|---|---|---|---|---|---|---|")
    tot = 0.0
    n = 0
    for r in body:
        nm = re.sub(r"\(.*", "", r[idx["name"]])
        nm = re.sub(r"^void\s+", "", nm).replace("aedit::<unnamed>::", "").replace("aedit::(anonymous namespace)::", "")
        t = scale(r[idx["time"]], units[want["time"]], 1.0)
        rd = scale(r[idx["rd"]], units[want["rd"]], 1e6)
        wr = scale(r[idx["wr"]], units[want["wr"]], 1e6)
        tot += (rd + wr) * 1e6
        n += 1

        def g(k):
            return r[idx[k]] if idx[k] is not None else "-"
        print(f"| {nm} | {g('grid')} | {t:.1f} | {rd:.1f} | {wr:.1f} | {g('tensor')} | {g('dram_pct')} | {g('l2hit')} | {g('regs')} |")
    if n:
        print(f"\nMean DRAM traffic per launch over {n} launches: {tot / n / 1e6:.1f} MB")
        print(f"TRAFFIC_JSON {{\"traffic_bytes_per_launch\": {tot / n}, \"launches\": {n}}}")


if __name__ == "__main__":
    main(sys.argv[1])
