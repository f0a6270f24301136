"""profiles/r02_roofline_traffic.json from ncu --set full captures (read here, without a GPU):

    python tools/ncu_traffic.py <bench key>=<file.ncu-rep> [...]

For every capture: dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the captured launches (for the pair-GEMM
instance these are the 59 launches of one step: all its shapes and epilogues, exactly the population bench.py pools under that key).
bench.py attaches the figure to `roofline.traffic` / `roofline_hbm.traffic` when the key matches the kernel it reports."""
import csv
import json
import os
import subprocess
import sys

out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_roofline_traffic.json")
res = json.load(open(out_path)) if os.path.exists(out_path) else {}
for arg in sys.argv[1:]:
    key, _, path = arg.partition("=")
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def col(r, name):
        v, u = float(r[ix[name]].replace(",", "")), units[ix[name]].lower()
        return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)

    data = rows[2:]
    rd = [col(r, "dram__bytes_read.sum") for r in data]
    wr = [col(r, "dram__bytes_write.sum") for r in data]
    dur = [float(r[ix["gpu__time_duration.sum"]].replace(",", "")) for r in data]
    names = sorted({r[ix["Kernel Name"]][:90] for r in data})
    res[key] = {"dram_bytes_per_launch": round((sum(rd) + sum(wr)) / len(data)), "dram_read_bytes_per_launch": round(sum(rd) / len(data)),
                "dram_write_bytes_per_launch": round(sum(wr) / len(data)), "launches_captured": len(data),
                "mean_duration_under_ncu_" + units[ix["gpu__time_duration.sum"]]: round(sum(dur) / len(data), 2),
                "source": f"ncu --set full --clock-control none, {os.path.basename(path)} ({len(data)} launches of one step inside bench.py)", "kernels": names}
    print(key, res[key]["dram_bytes_per_launch"], "B/launch over", len(data), "launches")
json.dump(res, open(out_path, "w"), indent=1)
print("wrote", out_path)
