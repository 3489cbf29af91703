"""Key metrics per kernel launch out of an `ncu --set full` report (run here, no GPU needed):

    python tools/ncu_extract.py gpurun_out/prof.ncu-rep [> profiles/rN_ncu_<kernel>.txt]

Prints, for every captured launch: duration, DRAM bytes read/written (+ throughput against the measured peak), L2->SM
bytes, every metric whose name mentions the tensor pipe (the UTCHMMA / tcgen05 evidence), achieved occupancy,
registers, shared memory."""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    name_col = hdr.index("Kernel Name")
    peaks = {}
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        print("=" * 100)
        print(r[name_col][:140])
        vals = {}
        for h, u, v in zip(hdr, units, r):
            tensor = ("tensor" in h and "sm__ops_path" not in h and ".min" not in h and ".max" not in h and ".sum" not in h
                      and "imma" not in h and "dmma" not in h and h != "device__attribute_tensor_map_access_supported")
            tf32 = "tf32" in h and h.endswith(".avg.pct_of_peak_sustained_elapsed")
            if h in KEEP or tensor or tf32:
                vals[h] = (v, u)
        for h in sorted(vals):
            v, u = vals[h]
            print(f"  {h:75s} {v:>18s} {u}")
        # where the warps wait: the six largest "stalled per issue" reasons, and shared-memory bank conflicts
        stalls = []
        for h, u, v in zip(hdr, units, r):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(v.replace(",", "")), h))
                except ValueError:
                    pass
            if "bank_conflict" in h and h.endswith(".sum"):
                print(f"  {h:75s} {v:>18s} {u}")
        for val, h in sorted(stalls, reverse=True)[:6]:
            print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {val:8.2f} warps per issue-active cycle")
        try:
            dur_ns = float(vals["gpu__time_duration.sum"][0].replace(",", ""))
            unit = vals["gpu__time_duration.sum"][1]
            dur_s = dur_ns * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit, 1e-9)

            def nbytes(k):
                v, u = vals[k]
                return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            tot = nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")
            line = f"  -> DRAM traffic {tot / 1e6:.1f} MB per launch, {tot / dur_s / 1e9:.0f} GB/s"
            if peaks.get("hbm_gbs"):
                line += f" = {tot / dur_s / 1e9 / peaks['hbm_gbs']:.3f} of the measured {peaks['hbm_gbs']:.0f} GB/s"
            print(line)
        except (KeyError, ValueError):
            pass


if __name__ == "__main__":
    main(sys.argv[1])
