#!/usr/bin/env python3
"""profiles/r02_roofline_traffic.json from `ncu --page raw --csv` exports of the ACSF value kernel:
   python tools/ncu_to_traffic.py c2=<raw.csv>:<atoms> [c3=<raw.csv>:<atoms>] > profiles/r02_roofline_traffic.json
(bench.py reads it for roofline.traffic -- DRAM bytes per launch of the dominant kernel, scaled by atoms)."""
import csv, json, sys
out = {}
for arg in sys.argv[1:]:
    wl, rest = arg.split("=")
    path, atoms = rest.rsplit(":", 1)
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    def get(name, scale_units=True):
        i = hdr.index(name)
        v = float(vals[i].replace(",", ""))
        u = units[i].lower()
        if scale_units:
            v *= {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        return v
    out[wl] = {
        "kernel": vals[hdr.index("Kernel Name")].split("(")[0].replace("void ", ""),
        "atoms_per_launch": int(atoms),
        "dram_bytes_read": get("dram__bytes_read.sum"),
        "dram_bytes_write": get("dram__bytes_write.sum"),
        "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
        "fp64_pipe_active_pct": get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", False),
        "warp_inst_per_atom": round(get("smsp__inst_executed.sum", False) / int(atoms)),
        "gpu_time_us": get("gpu__time_duration.sum", False) * ({"ms": 1e3, "us": 1.0, "ns": 1e-3}.get(units[hdr.index("gpu__time_duration.sum")].lower(), 1.0)),
        "source": "ncu --set full --clock-control none, one launch of %s atoms (%s)" % (atoms, path.split("/")[-1]),
    }
print(json.dumps(out, indent=1))
