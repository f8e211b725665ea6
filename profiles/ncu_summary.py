"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small table.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xxx.txt]
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("%-44s" % "kernel" + "".join("%14s" % n for _, n in WANT))
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:43]
        cells = []
        for m, _ in WANT:
            if m in idx:
                v, u = r[idx[m]], units[idx[m]]
                try:
                    f = float(v.replace(",", ""))
                    v = "%.4g" % f
                except ValueError:
                    pass
                cells.append("%14s" % (v + (" " + u if u in ("us", "ms", "Mbyte", "Gbyte", "Kbyte", "byte") else "")))
            else:
                cells.append("%14s" % "-")
        print("%-44s" % name + "".join(cells))


if __name__ == "__main__":
    main(sys.argv[1])
