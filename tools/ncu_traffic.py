"""Extract per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) and duration of the scan kernels from
.ncu-rep files (ncu --set full captures of tools/ncu_one.py) and write profiles/ncu_traffic.json, which bench.py quotes
as roofline.traffic.

    python tools/ncu_traffic.py configs1_f32 gpurun_out/x.ncu-rep [more.ncu-rep ...]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    key, reps = sys.argv[1], sys.argv[2:]
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    entry = out.setdefault(key, {})
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        head, units = rows[0], rows[1]
        for r in rows[2:]:
            m = dict(zip(head, r))
            u = dict(zip(head, units))
            name = m["Kernel Name"]
            which = "bwd" if "bwd" in name else "fwd"
            rd = float(m["dram__bytes_read.sum"]) * UNIT.get(u["dram__bytes_read.sum"], 1)
            wr = float(m["dram__bytes_write.sum"]) * UNIT.get(u["dram__bytes_write.sum"], 1)
            entry[which] = {"kernel": name.split("(")[0], "dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
                            "duration_us_under_ncu": float(m["gpu__time_duration.sum"]) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u["gpu__time_duration.sum"], 1),
                            "report": os.path.basename(rep)}
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
