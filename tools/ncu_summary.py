#!/usr/bin/env python
"""Summarises an .ncu-rep (ncu --set full --import-source on) into the text files committed under profiles/:
key raw metrics of every captured launch + the stall-reason totals and hottest SASS lines of the source page.
Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.txt ["free-text note"]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    lines = [f"# ncu summary of {rep}", f"# {note}", ""]
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if " at::" in " " + d.get("Kernel Name", ""):   # torch helper kernels of the driving script
            continue
        lines.append(f"## launch: {d.get('Kernel Name', '?')[:150]}")
        for k in hdr:
            if k in WANT or ("inst_executed_pipe_" in k and k.endswith(".avg.pct_of_peak_sustained_active") and k.startswith("sm__")):
                lines.append(f"{k:90s} {d[k]}")
        lines.append("")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    if len(src) > 2:
        h = src[1]
        ix = {c: i for i, c in enumerate(h)}
        stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        tot, total, top = {s: 0 for s in stalls}, 0, []
        for r in src[2:]:
            try:
                n = int(r[ix["# Samples"]])
            except Exception:
                continue
            total += n
            for s in stalls:
                try:
                    tot[s] += int(r[ix[s]])
                except Exception:
                    pass
            why = sorted(((int(r[ix[s_]] or 0), s_[6:]) for s_ in stalls if (r[ix[s_]] or "0").isdigit()), reverse=True)[:2]
            top.append((n, r[ix["Source"]][:90] + "   <- " + ", ".join(f"{w}:{c}" for c, w in why if c)))
        lines.append(f"## warp-stall sampling (first launch), {total} samples")
        for s, v in sorted(tot.items(), key=lambda x: -x[1])[:8]:
            lines.append(f"{s:28s}{v:8d} {100.0 * v / max(total, 1):5.1f}%")
        lines.append("## hottest SASS lines")
        for n, s in sorted(top, reverse=True)[:40]:
            lines.append(f"{n:7d}  {s}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
