"""Summarise an .ncu-rep: key raw metrics per captured launch + stall samples by SASS region.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--src]"""
import csv, io, subprocess, sys, collections

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum ", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct", "sm__cycles_elapsed.max ", "gpc__cycles_elapsed.max.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum ",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "sm__inst_executed.sum.per_cycle_elapsed", "launch__registers_per_thread ",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "smsp__inst_executed.sum ",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "smsp__pcsamp_sample_buffer", "sm__cycles_active.avg "]

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("==", name[:90])
        for h, u, v in zip(hdr, units, vals):
            # section-prefixed columns ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime...") carry the metric
            # name after the last section prefix: match on that
            base = h.split("TriageCompute.")[-1]
            if v != "" and any((base + " ").startswith(k) or k.strip() == base for k in KEYS):
                print(f"  {h} [{u}] = {v}")

def src(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[hi], rows[hi + 1:]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp] or 0) for r in data)
    print("total samples", tot, "instructions", len(data))
    for start in range(0, len(data), 200):
        seg = data[start:start + 200]
        s = sum(int(r[isamp] or 0) for r in seg)
        ex = sum(int(r[iex] or 0) for r in seg)
        kinds = collections.Counter()
        for r in seg:
            for k in ["UTCHMMA", "UTMALDG", "LDTM", "SYNCS.PHASECHK", "SHFL", "STG", "FSETP", "LDG", "FMNMX", "HMMA", "LDS", "BAR", "ATOMS", "FFMA"]:
                if k in r[isrc]:
                    kinds[k] += 1
        if s or ex:
            print(f"  [{start:5d}] samples {s:8d} ({100 * s / max(tot, 1):5.1f}%) exec {ex:12d} {dict(kinds)}")
    top = sorted(data, key=lambda r: -int(r[isamp] or 0))[:18]
    for r in top:
        st = {hdr[i].replace("stall_", ""): int(r[i] or 0) for i in stall if int(r[i] or 0) > 0}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"  {r[isamp]:>8} x{r[iex]:>11} {r[isrc][:64]:64s} {st}")

if __name__ == "__main__":
    raw(sys.argv[1])
    if "--src" in sys.argv:
        src(sys.argv[1])
