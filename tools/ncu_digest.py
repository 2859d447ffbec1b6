"""Digest of an `ncu --page raw --csv` export: one block per kernel launch with the metrics the rooflines are argued from.
usage: python tools/ncu_digest.py raw.csv [title] > profiles/xxx.txt"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("smsp__cycles_elapsed.avg.per_second", "SM clock"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (elapsed)"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "ALU pipe %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "LSU pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "shared-memory wavefronts %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__inst_executed_pipe_alu.sum", "  on ALU"), ("smsp__inst_executed_pipe_fmaheavy.sum", "  on FMA-heavy"), ("smsp__inst_executed_pipe_lsu.sum", "  on LSU"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
print("#", sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("unnamed>::", "")
    print(f"\n== {name}   [launch id {r[idx['ID']]}]")
    seen = set()
    for k, label in KEYS:
        if k in idx and r[idx[k]] not in ("", "n/a") and label not in seen:
            seen.add(label)
            print(f"  {label:30s} {r[idx[k]]} {units[idx[k]]}")
    st = [(hdr[i][len(STALLS):], float(r[i])) for i in range(len(hdr)) if hdr[i].startswith(STALLS) and not hdr[i].endswith("_not_issued") and r[i] not in ("", "n/a")]
    tot = sum(v for _, v in st)
    if tot:
        top = sorted(st, key=lambda x: -x[1])[:7]
        print("  stall samples: " + ", ".join(f"{n} {100 * v / tot:.1f}%" for n, v in top))
