"""Turns an .ncu-rep (read here, no GPU needed) into the small tracked files under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_scan_final [scan_kernel]

writes <out>_summary.csv (selected raw metrics, one column per captured launch), <out>_stalls.txt (top
instructions by warp-stall samples) and, for the scan kernels, profiles/scan_traffic.json (scan_kernel) or
profiles/screen_traffic.json (screen_kernel)."""
import csv
import json
import subprocess
import sys

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__m_l1tex2xbar_write_sectors_mem_lg_op_st.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep, outp = sys.argv[1], sys.argv[2]
    kname = sys.argv[3] if len(sys.argv) > 3 else None
    rows = raw(rep)
    hdr, units, body = rows[0], rows[1], rows[2:]
    stall = [h for h in hdr if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")]
    with open(outp + "_summary.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch {i}" for i in range(len(body))])
        for h in KEEP + stall:
            if h in hdr:
                i = hdr.index(h)
                w.writerow([h, units[i]] + [r[i] for r in body])
    if kname:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kname],
                             capture_output=True, text=True).stdout
        srows = list(csv.reader(src.splitlines()))
        hidx = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
        if hidx:
            h = srows[hidx[0]]
            body2 = srows[hidx[0] + 1:(hidx[1] - 1 if len(hidx) > 1 else len(srows))]
            si, so = h.index("# Samples"), h.index("Source")
            stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
            agg = {c: 0 for c in stalls}
            for r in body2:
                for c in stalls:
                    try:
                        agg[c] += int(r[h.index(c)] or 0)
                    except Exception:
                        pass
            with open(outp + "_stalls.txt", "w") as f:
                tot = sum(int(r[si] or 0) for r in body2 if len(r) > si)
                f.write(f"kernel {kname}: {tot} warp-stall samples (first captured launch)\n")
                for c, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
                    f.write(f"  {c:28s} {v:8d}  {100.0 * v / max(tot, 1):5.1f}%\n")
                f.write("top instructions by samples:\n")
                for r in sorted(body2, key=lambda r: -int(r[si] or 0))[:30]:
                    f.write(f"  {r[si]:>7s}  {r[so][:110]}\n")
        if kname in ("scan_kernel", "screen_kernel"):
            # the longest captured launch is the list scan (the coarse step no longer uses this kernel)
            ti, ri, wi = hdr.index("gpu__time_duration.sum"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            ui, uw = units[ri], units[wi]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            best = max(body, key=lambda r: float(r[ti]))
            tr = float(best[ri]) * scale[ui] + float(best[wi]) * scale[uw]
            json.dump({"dram_bytes_per_launch": tr, "dram_read": float(best[ri]) * scale[ui],
                       "dram_write": float(best[wi]) * scale[uw], "kernel_time": best[ti] + " " + units[ti],
                       "source": rep.split("/")[-1]},
                      open("profiles/screen_traffic.json" if kname == "screen_kernel" else "profiles/scan_traffic.json", "w"),
                      indent=1)


if __name__ == "__main__":
    main()
