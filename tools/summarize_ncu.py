"""Turns ncu outputs into the markdown tables of profiles/*.md.

  python tools/summarize_ncu.py launches <launches.csv> [skip_frames frames]   per-kernel average time and share
  python tools/summarize_ncu.py raw <file.ncu-rep> [kernel regex]               the metrics quoted in DESIGN.md / bench.py
  python tools/summarize_ncu.py lines <file.ncu-rep> <kernel substring> [n]     hottest source lines of one kernel
  python tools/summarize_ncu.py traffic <cascade.ncu-rep> <label> [old.json]    the counters bench.py reads (profiles/traffic.json);
                                                                               keys it cannot derive (tracker_*) are carried over from old.json

`raw` and `lines` need `ncu` (reading a report needs no GPU)."""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
           ("lts__t_sectors.sum", "L2 sectors (32 B)"), ("lts__t_sectors_srcunit_tex.sum", "L2 sectors from L1/TEX"),
           ("l1tex__t_sectors.sum", "L1/TEX sectors"), ("smsp__inst_executed.sum", "warp instructions"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
           ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
           ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
           ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "of which bank-conflict replays"),
           ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
           ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps resident %"),
           ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"),
           ("launch__occupancy_limit_shared_mem", "blocks/SM allowed by shared memory"),
           ("launch__occupancy_limit_registers", "blocks/SM allowed by registers"),
           ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier (warps per issue)"),
           ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
           ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
           ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle"),
           ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall: membar")]


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*", "", name)


def launches(path, skip=0):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    per = collections.OrderedDict()
    for r in rows[1 + skip:]:
        per.setdefault(short(r[k]), []).append(float(r[v].replace(",", "")) / 1e3)
    total = sum(sum(x) for x in per.values())
    print("| kernel | launches | avg us | share |\n|---|---|---|---|")
    for name, x in per.items():
        print(f"| {name} | {len(x)} | {sum(x) / len(x):.1f} | {100 * sum(x) / total:.1f} % |")
    print(f"\ntotal {total:.1f} us over {sum(len(x) for x in per.values())} launches")


def raw(rep, pattern=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    sel = [r for r in rows[2:] if not pattern or re.search(pattern, r[idx["Kernel Name"]])]
    print("| metric | " + " | ".join(short(r[idx["Kernel Name"]]) for r in sel) + " |\n|---|" + "---|" * len(sel))
    for m, label in METRICS:
        if m in idx:
            vals = []
            for r in sel:
                try:
                    x = float(r[idx[m]].replace(",", ""))
                    vals.append(f"{x:,.0f}" if x >= 1000 else f"{x:.2f}".rstrip("0").rstrip("."))
                except ValueError:
                    vals.append(r[idx[m]])
            print(f"| {label} ({units[idx[m]]}) | " + " | ".join(vals) + " |")


def lines(rep, kernel, n=15):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur = fn = None
    data = collections.defaultdict(list)
    for r in csv.reader(io.StringIO(out)):
        if len(r) >= 2 and r[0] == "Function Name":
            fn = r[1]
        elif len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 8 and r[0].isdigit() and r[2] == "-":
            try:
                data[fn].append((int(r[6]), int(r[7]), cur, int(r[0]), r[1].strip()[:100]))
            except ValueError:
                pass
    for fn, rows in data.items():
        if kernel not in fn:
            continue
        ts, ti = sum(r[0] for r in rows) or 1, sum(r[1] for r in rows) or 1
        print(f"{fn}: {ts} stall samples, {ti} warp instructions\n\n| samples | instructions | line | source |\n|---|---|---|---|")
        for r in sorted(rows, reverse=True)[:n]:
            print(f"| {100 * r[0] / ts:.1f} % | {100 * r[1] / ti:.1f} % | {r[2]}:{r[3]} | `{r[4]}` |")
        break


def traffic(rep, label, old=None):
    """One config-3 frame's cascade kernels (stage 0, bulk, tail) from a --set full capture -> the counters of bench.py's roofline."""
    import json
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    idx = {h: i for i, h in enumerate(rows[0])}
    for h, i in list(idx.items()):                               # "SM_B.TriageCompute.l1tex__t_sectors.sum" -> "l1tex__t_sectors.sum"
        idx.setdefault(h.split("TriageCompute.")[-1], i)
    num = lambda r, m: float(r[idx[m]].replace(",", ""))      # noqa: E731
    unit = lambda m: rows[1][idx[m]]                             # noqa: E731
    casc = [r for r in rows[2:] if re.search(r"k_stage0|k_cascade", r[idx["Kernel Name"]])]
    names = [short(r[idx["Kernel Name"]]) for r in casc]
    per_frame = len(set(names))
    casc, names = casc[:per_frame], names[:per_frame]           # the first frame of the capture
    bulk = [r for r, n in zip(casc, names) if re.search(r"k_cascade_classes|k_cascade_wide", n)]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = sum(num(r, m) * scale[unit(m)] for r in casc for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    d = json.load(open(old)) if old else {}
    d.update({
        "cascade_kernels": names,
        "cascade_dram_bytes_per_frame": dram,
        "cascade_warp_instructions_per_frame": sum(num(r, "smsp__inst_executed.sum") for r in casc),
        "source": f"{label}: dram__bytes_read.sum + dram__bytes_write.sum summed over {', '.join(names)} for one cfg3 frame (ncu --set full --clock-control none)",
        "tile_kernels": [short(r[idx["Kernel Name"]]) for r in bulk],
        "tile_shared_wavefronts_per_frame": sum(num(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") for r in bulk),
        "tile_shared_bank_conflict_wavefronts_per_frame": sum(num(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") for r in bulk),
        "tile_warp_instructions_per_frame": sum(num(r, "smsp__inst_executed.sum") for r in bulk),
        "wavefront_source": f"{label}: l1tex__data_pipe_lsu_wavefronts_mem_shared.sum and smsp__inst_executed.sum of the two bulk kernels, one cfg3 frame",
        "tile_lts_bytes_per_frame": 32 * sum(num(r, "lts__t_sectors.sum") for r in bulk),
        "tile_l1tex_bytes_per_frame": 32 * sum(num(r, "l1tex__t_sectors.sum") for r in bulk),
        "lts_ncu_pct_of_peak": [round(num(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"), 2) for r in bulk],
        "l1tex_ncu_pct_of_peak": [round(num(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), 2) for r in bulk],
        "lts_source": f"{label}: 32 B x lts__t_sectors.sum (L2) and 32 B x l1tex__t_sectors.sum (L1/TEX) of the two bulk launches of one cfg3 frame, with ncu's own lts__throughput / l1tex__throughput percentages of peak per launch",
    })
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
        sys.exit(0)
    if cmd == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    elif cmd == "raw":
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        lines(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 15)
