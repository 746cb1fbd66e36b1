"""Summarise an .ncu-rep (one kernel, `ncu --set full --import-source on`) into the text kept under profiles/:
headline metrics, warp-stall breakdown and the most-sampled SASS lines.  python profiles/ncu_summary.py rep.ncu-rep [N]"""
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    d = dict(zip(hdr, zip(vals, units)))
    print("kernel:", d.get("Kernel Name", ("?",))[0])
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    for k in keys:
        for h in hdr:
            if h == k or h.endswith("." + k):
                print(f"  {h} = {d[h][0]} {d[h][1]}")
                break
    src = page(rep, "source")
    h = src[1]
    s_i, src_i, ie_i = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    stalls = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    body = [r for r in src[2:] if len(r) > max(stalls) and r[s_i].isdigit()]
    tot = sum(int(r[s_i]) for r in body) or 1
    agg = {}
    for r in body:
        for i in stalls:
            agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
    print("warp-stall samples (all):", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    print(f"most-sampled SASS ({tot} samples): share, executions, instruction, top stall")
    for r in sorted(body, key=lambda r: -int(r[s_i]))[:topn]:
        top = max(((int(r[i] or 0), h[i][6:]) for i in stalls))
        print(f"  {100 * int(r[s_i]) / tot:5.1f}%  {r[ie_i]:>10}  {r[src_i].strip()[:95]:95s}  {top[1]}")


if __name__ == "__main__":
    main()
