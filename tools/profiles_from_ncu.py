"""Turn the ncu reports of tools/profile_round.sh into the committed summaries under profiles/:
   <tag>_ncu_k1_summary.json / <tag>_ncu_k2_summary.json (per launch: duration, DRAM bytes, DRAM / L2 / SM %, occupancy, ...)
   dram_traffic.json (the DRAM bytes per launch that bench.py quotes as roofline.traffic / roofline_k2.*.traffic)
usage: python tools/profiles_from_ncu.py gpurun_out/r2z r2z"""
import json
import os
import subprocess
import sys

src, tag = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prof = os.path.join(root, "profiles")
traffic = {}
for name in ("k1", "k2"):
    rep = os.path.join(src, name + ".ncu-rep")
    raw = os.path.join(src, name + "_raw.csv")
    if not os.path.exists(raw):
        if not os.path.exists(rep):
            continue
        with open(raw, "w") as f:
            subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL, check=True)
    out = os.path.join(prof, f"{tag}_ncu_{name}_summary.json")
    subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), raw, out], check=True)
    rows = json.load(open(out))

    def to_bytes(v, unit):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    for r in rows:
        r["dram_bytes_read_B"] = to_bytes(r.get("dram_bytes_read", 0), r.get("dram_bytes_read_unit", "byte"))
        r["dram_bytes_write_B"] = to_bytes(r.get("dram_bytes_write", 0), r.get("dram_bytes_write_unit", "byte"))
    json.dump(rows, open(out, "w"), indent=1)

    def pick(prefix, which=-1):
        c = [r for r in rows if str(r.get("kernel", "")).lstrip("void ").startswith(prefix)]
        return c[which] if c else None
    if name == "k1":
        r = pick("k_sweep_window")
        if r:
            traffic["k_sweep_window"] = {"batch": 1024, "dram_bytes_read": int(r["dram_bytes_read_B"]), "dram_bytes_write": int(r["dram_bytes_write_B"]),
                                         "source": f"profiles/{tag}_ncu_k1_summary.json (ncu --set full, bench.py --steps 1 --warmup 1)"}
    else:
        r = pick("k_raytrace")
        if r:
            traffic["k_raytrace"] = {"batch": 2000, "dram_bytes_read": int(r["dram_bytes_read_B"]), "dram_bytes_write": int(r["dram_bytes_write_B"]),
                                     "source": f"profiles/{tag}_ncu_k2_summary.json (2000-scan rebuild)"}
        r = pick("k_gm_update")
        if r:
            traffic["k_gm_update"] = {"dram_bytes_read": int(r["dram_bytes_read_B"]), "dram_bytes_write": int(r["dram_bytes_write_B"]),
                                      "source": f"profiles/{tag}_ncu_k2_summary.json (one scan)"}
        marks = [x for x in rows if str(x.get("kernel", "")).startswith("k_hs_mark")]
        applies = [x for x in rows if str(x.get("kernel", "")).startswith("k_hs_apply")]
        # launch order of tools/profile_k2.py: 4 room steps (mapping), 4 hall steps, then 2 SLAM steps
        if len(marks) >= 8 and len(applies) >= 8:
            for key, i in (("k_hs_batched_update", 2), ("k_hs_batched_update_hall", 6)):
                traffic[key] = {"batch": 128,
                                "dram_bytes_read": int(marks[i]["dram_bytes_read_B"] + applies[i]["dram_bytes_read_B"]),
                                "dram_bytes_write": int(marks[i]["dram_bytes_write_B"] + applies[i]["dram_bytes_write_B"]),
                                "source": f"profiles/{tag}_ncu_k2_summary.json (k_hs_mark + k_hs_apply of one 128-map step)"}
if traffic:
    path = os.path.join(prof, "dram_traffic.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update(traffic)
    json.dump(old, open(path, "w"), indent=1)
    print(json.dumps(traffic, indent=1))
