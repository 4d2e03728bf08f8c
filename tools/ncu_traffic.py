#!/usr/bin/env python
"""profiles/traffic.json from an ncu report of tools/ncu_targets.py: DRAM bytes (read + write) per launch.
Usage: tools/ncu_traffic.py gpurun_out/prof_targets.ncu-rep profiles/traffic.json profiles/<summary>.txt"""
import csv
import io
import json
import subprocess
import sys


def main():
    rep, out, src = sys.argv[1], sys.argv[2], sys.argv[3]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ix = {c: i for i, c in enumerate(hdr)}

    def bytes_of(r, col):
        v = float(r[ix[col]].replace(",", ""))
        u = units[ix[col]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    launches = []
    for r in rows[2:]:
        launches.append({"kernel": r[ix["Kernel Name"]][:120],
                         "dram_bytes": bytes_of(r, "dram__bytes_read.sum") + bytes_of(r, "dram__bytes_write.sum"),
                         "us": float(r[ix["gpu__time_duration.sum"]].replace(",", "")) *
                         {"nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(units[ix["gpu__time_duration.sum"]].lower(), 1)})
    gem = [l for l in launches if "gemm_tn_cg2_kernel" in l["kernel"]]
    sweep = [l for l in launches if "fit_adam_table" in l["kernel"]]
    res = {"source": src, "launches": launches}
    if len(gem) >= 4:
        res["gemm_tn_cg2_kernel"] = {"dram_bytes_per_launch": sum(l["dram_bytes"] for l in gem[:4]),
                                                 "note": "sum over the 4-launch set (qkv, proj, fc1, fc2) at batch 16"}
    if sweep:
        res["fit_adam_table_kernel"] = {"dram_bytes_per_launch": sweep[0]["dram_bytes"]}
    for key in ("attention_tc_kernel", "attention_bwd_tc_kernel"):
        hit = [l for l in launches if key in l["kernel"]]
        if hit:
            res[key] = {"dram_bytes_per_launch": hit[0]["dram_bytes"], "us": hit[0]["us"]}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1)[:1500])


if __name__ == "__main__":
    main()
