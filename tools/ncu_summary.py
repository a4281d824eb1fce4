"""Summarise an .ncu-rep: key launch/occupancy/throughput metrics, stall mix and per-pair instruction mix per kernel."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
pairs = float(sys.argv[2]) if len(sys.argv) > 2 else 403e6 / 32  # warp-level (t, channel, state) triples of configs[1]


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


det = list(csv.DictReader(io.StringIO(ncu("--page", "details", "--csv"))))
want = ['Duration', 'Elapsed Cycles', 'SM Frequency', 'DRAM Throughput', 'Compute (SM) Throughput', 'Registers Per Thread',
        'Achieved Occupancy', 'Theoretical Occupancy', 'Block Limit', 'Executed Ipc Active', 'Issue Slots Busy',
        'Dynamic Shared Memory Per Block', 'Waves Per SM', 'No Eligible', 'Eligible Warps', 'Active Warps Per Scheduler',
        'L2 Hit Rate', 'Mem Busy', 'Max Bandwidth', 'Memory Throughput']
kernels = []
for r in det:
    if r['Kernel Name'] not in kernels:
        kernels.append(r['Kernel Name'])
for k in kernels:
    print("=" * 10, k[:110])
    for r in det:
        if r['Kernel Name'] == k and any(w.lower() in r['Metric Name'].lower() for w in want):
            print(f"  {r['Metric Name']:42s} {r['Metric Value']:>14s} {r['Metric Unit']}")
    short = "scan_fwd" if "scan_fwd" in k else ("scan_bwd" if "scan_bwd" in k else k.split("(")[0].split("::")[-1][:30])
    raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv", "--kernel-name", f"regex:{short}"))))
    if len(raw) >= 3:
        m = dict(zip(raw[0], raw[2]))
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "smsp__inst_executed.sum",
                    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_alu.sum",
                    "sm__inst_executed_pipe_lsu.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
                    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_active.avg.per_cycle_active",
                    # the SM's L1 / shared-memory data pipe (1 wavefront per clock per SM): the resource the scan kernels run into
                    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
                    "l1tex__data_pipe_lsu_wavefronts.max.pct_of_peak_sustained_elapsed",
                    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
                    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
                    "smsp__issue_active.avg.pct_of_peak_sustained_active"):
            if key in m:
                print(f"  {key:58s} {m[key]}")
    src = ncu("--page", "source", "--csv", "--kernel-name", f"regex:{short}")
    lines = src.splitlines()
    if len(lines) > 2:
        rows = list(csv.DictReader(io.StringIO("\n".join(lines[1:]))))
        stall = [c for c in rows[0].keys() if c and c.startswith('stall_') and 'Not Issued' not in c]
        tot = {c: 0 for c in stall}
        ops, inst = {}, 0
        for r in rows:
            for c in stall:
                try:
                    tot[c] += int(r[c] or 0)
                except ValueError:
                    pass
            n = int(r['Instructions Executed'] or 0)
            inst += n
            toks = r['Source'].split()
            op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
            op = op.split('.')[0]
            ops[op] = ops.get(op, 0) + n
        S = max(1, sum(tot.values()))
        print("  stalls %:", {c[6:]: round(v / S * 100, 1) for c, v in sorted(tot.items(), key=lambda x: -x[1]) if v / S > 0.01})
        print(f"  warp instructions {inst}  -> {inst / pairs:.2f} per warp-pair")
        print("  per pair:", sorted(((o, round(v / pairs, 2)) for o, v in ops.items() if v / pairs >= 0.05), key=lambda x: -x[1]))
        top = sorted(rows, key=lambda r: -int(r['# Samples'] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 12]
        for r in top:
            t2 = sorted(((c[6:], int(r[c] or 0)) for c in stall), key=lambda x: -x[1])[:2]
            print("   ", r['Address'][-5:], r['# Samples'].rjust(6), r['Instructions Executed'].rjust(9), r['Source'][:64].ljust(64), t2)
