"""Derive bench.py's `roofline.traffic` from an `ncu --set full` capture of the dominant kernel.

usage: python tools/ncu_traffic.py <report.ncu-rep> --kernel k_pairs_w2 --qubits 32 [--out profiles/r02_traffic.json]
Reads the report with `ncu -i ... --page raw --csv`, averages dram__bytes_read.sum +
dram__bytes_write.sum over the captured launches of the kernel and writes a small JSON next to the
report's name, so that the number in the bench line can be traced to a committed capture.
"""
import argparse
import csv
import io
import json
import os
import subprocess


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale[unit]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", default="k_pairs_w2")
    ap.add_argument("--qubits", type=int, required=True, help="local qubits of the captured run")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    picked = []
    for r in rows[2:]:
        if a.kernel in r[col["Kernel Name"]]:
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            ms = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
            u = units[col["gpu__time_duration.sum"]]
            ms *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
            picked.append((rd, wr, ms, r[col["Kernel Name"]]))
    if not picked:
        raise SystemExit(f"no launch of {a.kernel} in {a.report}")
    rd = sum(p[0] for p in picked) / len(picked)
    wr = sum(p[1] for p in picked) / len(picked)
    out = {
        "kernel": picked[0][3], "qubits": a.qubits, "launches": len(picked),
        "dram_read_bytes_per_launch": rd, "dram_write_bytes_per_launch": wr, "dram_bytes_per_launch": rd + wr,
        "algorithmic_bytes_per_launch": 32.0 * float(1 << a.qubits), "ratio_to_algorithmic": (rd + wr) / (32.0 * float(1 << a.qubits)),
        "ncu_ms_per_launch_cold": sum(p[2] for p in picked) / len(picked),
        "source": os.path.relpath(a.report), "how": "ncu --set full --clock-control none; dram__bytes_read.sum + dram__bytes_write.sum averaged over the captured launches",
    }
    path = a.out or os.path.join(os.path.dirname(a.report) or ".", "r02_traffic.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
