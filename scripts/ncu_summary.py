"""Key metrics per kernel from an .ncu-rep (via `ncu --page raw --csv`).

    python scripts/ncu_summary.py report.ncu-rep                      # text summary (what profiles/*_ncu_*.txt hold)
    python scripts/ncu_summary.py report.ncu-rep --json profiles/ncu_traffic.json
        # also writes {kernel: dram bytes per launch, duration, tensor-pipe / DRAM / issue utilisation, L1 load sectors} with the
        # report's name and the current commit; bench.py takes `roofline.traffic` from that file"""
import csv, json, os, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("----", r[ix["Kernel Name"]][:90])
    for w in want:
        if w in ix:
            print(f"  {w} [{units[ix[w]]}] = {r[ix[w]]}")

if "--json" in sys.argv:
    out_path = sys.argv[sys.argv.index("--json") + 1]
    def num(r, key):
        try:
            return float(r[ix[key]].replace(",", ""))
        except (KeyError, ValueError):
            return None
    def to_bytes(r, key):
        v = num(r, key)
        if v is None:
            return None
        unit = units[ix[key]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(unit, 1)
    kernels = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("hgn::", "")
        short = name[:-len("_kernel")] if name.endswith("_kernel") else name
        rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
        if rd is None or wr is None:
            continue
        entry = kernels.setdefault(short, {"launches": 0, "dram_bytes_per_launch": 0.0})
        entry["launches"] += 1
        entry["dram_bytes_per_launch"] += rd + wr
        dur = num(r, "gpu__time_duration.sum")
        dunit = units[ix["gpu__time_duration.sum"]].lower()
        entry["duration_us_profiled"] = None if dur is None else dur * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(dunit, 1.0)
        entry["l1_lsu_data_pipe_pct"] = num(r, "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed")
        entry["tensor_pipe_active_pct"] = num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        entry["dram_throughput_pct"] = num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        entry["issue_active_pct"] = num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        entry["l1_global_load_sectors"] = num(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
        entry["registers_per_thread"] = num(r, "launch__registers_per_thread")
    for entry in kernels.values():
        entry["dram_bytes_per_launch"] /= entry["launches"]
    try:
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=os.path.dirname(os.path.abspath(__file__))).stdout.strip()
    except OSError:
        commit = ""
    with open(out_path, "w") as f:
        json.dump({"source": os.path.basename(sys.argv[1]), "commit": commit, "how": "ncu --set full --clock-control none, one launch per kernel at the cfg5 size "
                   "(scripts/time_edge.py); dram__bytes_read.sum + dram__bytes_write.sum", "kernels": kernels}, f, indent=1)
    print("wrote", out_path)
