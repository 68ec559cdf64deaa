#!/usr/bin/env python
"""profiles/r02_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel,
from the committed raw pages of the round's `ncu --set full` captures (bench.py's roofline.traffic)."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = {"full_n256_spin2": "r02_k1_tma_n256_uhf_ncu_raw.csv",
       "packed8_n384_spin1": "r02_k2_n384_spin1_ncu_raw.csv",
       "packed8_n384_spin2": "r02_k2_n384_spin2_ncu_raw.csv"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    out = {}
    for key, fn in SRC.items():
        rows = list(csv.reader(open(os.path.join(ROOT, "profiles", fn))))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        per = []
        for r in rows[2:]:
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[col[m]].replace(",", "")) * UNIT[units[col[m]]]
            per.append(tot)
        out[key] = {"dram_bytes_per_launch": int(round(sum(per) / len(per))), "launches_profiled": len(per),
                    "kernel": rows[2][col["Kernel Name"]].split("(")[0].replace("void <unnamed>::", ""),
                    "gpu_time_ms": [float(r[col["gpu__time_duration.sum"]]) * (1e-3 if units[col["gpu__time_duration.sum"]] in ("us", "usecond") else 1.0) for r in rows[2:]],
                    "source": f"profiles/{fn}"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
