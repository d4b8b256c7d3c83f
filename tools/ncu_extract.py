"""Key metrics of every kernel in an .ncu-rep as one CSV row each (dev tool).
    python tools/ncu_extract.py gpurun_out/x.ncu-rep [...] > profiles/summary.csv"""
import csv, io, subprocess, sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_elapsed.max",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]
out = csv.writer(sys.stdout)
out.writerow(["report", "kernel"] + WANT)
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    # TPC.TriageCompute.* style names carry a prefix: match by suffix
    def col(name):
        for i, h in enumerate(hdr):
            if h == name or h.endswith("." + name):
                return i
        return None
    idx = [col(w) for w in WANT]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        vals = []
        for i in idx:
            vals.append("" if i is None else f"{r[i]} {units[i]}".strip())
        out.writerow([rep.split("/")[-1], r[kn].split("(")[0][:60]] + vals)
