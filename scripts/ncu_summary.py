"""Summarise an ncu report of the persistent kernel into a text file under profiles/ and refresh
profiles/gn_loop_traffic.json (the `roofline.traffic` figure of bench.py).
Usage: python scripts/ncu_summary.py gpurun_out/<tag>_gn_loop.ncu-rep profiles/<tag>_gn_loop_summary.txt "<note>" """
import csv, io, json, subprocess, sys

rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op", "sm__inst_executed_pipe_", "smsp__inst_executed.sum",
        "smsp__issue_active.avg", "issue_stalled", "launch__", "sm__warps_active.avg.pct_of_peak", "sm__cycles_elapsed.max",
        "sm__throughput.avg.pct", "gpu__compute_memory_throughput.avg.pct", "smsp__pcsamp_warps_issue_stalled")
lines = [f"# ncu --set full --clock-control none; {note}"]
traffic = None
for k, row in enumerate(rows[2:]):
    name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    lines.append(f"## launch {k}: {name}")
    vals = dict(zip(hdr, row))
    for h, u in zip(hdr, units):
        if any(t in h for t in KEEP) and "per_second" not in h and vals.get(h, "") != "":
            lines.append(f"{h} = {vals[h]} {u}")
    def num(key):
        v = vals.get(key, "").replace(",", "")
        return float(v) if v else 0.0
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = num("dram__bytes_read.sum") * scale.get(units[hdr.index("dram__bytes_read.sum")], 1.0)
    wr = num("dram__bytes_write.sum") * scale.get(units[hdr.index("dram__bytes_write.sum")], 1.0)
    traffic = int(rd + wr)
    l2_sectors = num("lts__t_sectors_srcunit_tex_op_read.sum")
    su = units[hdr.index("lts__t_sectors_srcunit_tex_op_read.sum")] if "lts__t_sectors_srcunit_tex_op_read.sum" in hdr else "sector"
    l2_bytes = int(l2_sectors * {"sector": 1.0, "Ksector": 1e3, "Msector": 1e6}.get(su, 1.0) * 32)
open(out, "w").write("\n".join(lines) + "\n")
if traffic is not None:
    json.dump({"dram_bytes_per_launch": traffic, "l2_to_l1_bytes_per_launch": l2_bytes,
               "source": f"{out}: dram__bytes_read.sum + dram__bytes_write.sum of the last captured k_gn_loop launch, ncu --set full "
                         f"(cold L2: ncu flushes caches before the replayed launch)"}, open("profiles/gn_loop_traffic.json", "w"))
print(out, "traffic bytes", traffic)
