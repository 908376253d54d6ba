#!/usr/bin/env python
"""Turns ncu exports into the markdown summaries committed under profiles/.

    python tools/summarise_ncu.py launches gpurun_out/r1_launches.csv            > profiles/...md
    python tools/summarise_ncu.py full     gpurun_out/r2_prof.ncu-rep [N lines]  > profiles/...md

`launches`: the per-launch list of `ncu --metrics gpu__time_duration.sum --clock-control none --csv`.
`full`: one `ncu --set full --import-source on` report; prints the headline counters per captured kernel and
the source lines ranked by executed instructions / stall samples (needs the ncu CLI to read the .ncu-rep).
"""
import collections
import csv
import io
import re
import subprocess
import sys

HEADLINE = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", row["Kernel Name"])).replace("void ", "")[:60]
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    ours = sum(v for k, v in tot.items() if k.startswith("spair::"))
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, v in tot.most_common():
        print("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (k, cnt[k], v, v / cnt[k], 100 * v / total))
    print("\n%d launches, %.2f ms of kernel time in the captured window; hand-written `spair::` kernels: %.2f ms (%.1f%%)."
          % (sum(cnt.values()), total / 1e3, ours / 1e3, 100 * ours / total))


def full(path, top=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("### `%s`\n" % r[hdr.index("Kernel Name")][:90])
        print("| metric | value |\n|---|---|")
        for m in HEADLINE:
            if m in hdr:
                i = hdr.index(m)
                print("| %s | %s %s |" % (m, r[i], units[i]))
        print()
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur, per_kernel, kernel = None, collections.defaultdict(list), None
    for r in csv.reader(io.StringIO(src)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            kernel = r[1][:80]
        elif r[0].strip().isdigit() and len(r) > 7 and r[7].isdigit():
            per_kernel[kernel].append((cur, int(r[0]), r[1].strip()[:110], int(r[7]), int(r[4]) if r[4].isdigit() else 0))
    for kernel, out in per_kernel.items():
        tot = sum(o[3] for o in out) or 1
        samp = sum(o[4] for o in out) or 1
        print("### hot source lines of `%s`\n\n| file:line | instr %% | stall-sample %% | source |\n|---|---:|---:|---|" % kernel)
        agg = {}
        for f, l, s, n, sm in out:          # one entry per (file, line): inlined copies and repeated launches are summed
            e = agg.setdefault((f, l), [f, l, s, 0, 0])
            e[3] += n
            e[4] += sm
        for f, l, s, n, sm in sorted(agg.values(), key=lambda o: -o[3])[:top]:
            print("| %s:%d | %.1f | %.1f | `%s` |" % (f, l, 100 * n / tot, 100 * sm / samp, s.replace("|", "\\|")))
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
