"""Turns an ncu report (gpurun_out/*.ncu-rep) into the text summaries committed under profiles/:
   python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/r01x_name"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__sass_inst_executed_op_utcmma.sum",
        "smsp__sass_inst_executed_op_tmem_ldt.sum", "smsp__sass_inst_executed_op_tmem_stt.sum",
        "smsp__sass_inst_executed_op_global_ld.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main(rep, out):
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr = rows[0]
    with open(out + "_metrics.txt", "w") as f:
        for r in rows[2:]:
            f.write(f"kernel: {r[hdr.index('Kernel Name')][:110]}\n")
            for i, h in enumerate(hdr):
                if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                    f.write(f"  {h} = {r[i]} {rows[1][i]}\n")
    rows = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    with open(out + "_opcodes.txt", "w") as f:
        for k, s in enumerate(starts[:1]):
            e = starts[k + 1] if k + 1 < len(starts) else len(rows)
            hdr = rows[s]
            ci, si, sm = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
            ops, smp = collections.Counter(), collections.Counter()
            for r in rows[s + 1:e]:
                if len(r) > ci and r[0].startswith("0x"):
                    t = r[si].split()
                    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
                    ops[op] += int(r[ci])
                    smp[op] += int(r[sm])
            tot = sum(ops.values())
            f.write(f"kernel: {rows[s - 1][1][:110]}\nwarp-level instructions executed: {tot}\n")
            f.write("opcode        executed      share   stall-samples\n")
            for op, c in ops.most_common(45):
                f.write(f"{op:12s} {c:12d} {100.0 * c / tot:7.2f}% {smp[op]:8d}\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
