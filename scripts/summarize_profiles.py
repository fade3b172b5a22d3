"""Turn the ncu artefacts a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py TAG        # e.g. r01b

Reads gpurun_out/launches_TAG.csv (per-launch device times of the bench command), gpurun_out/k1_full_TAG.ncu-rep and
gpurun_out/update_full_TAG.ncu-rep (ncu --set full captures) with `ncu -i ... --page raw --csv`; writes
profiles/TAG_launches.md, profiles/TAG_ncu_<kernel>.md and profiles/k1_ncu_traffic.json (dram bytes per launch of
the bench kernel, which bench.py reports as roofline.traffic).  Runs on the CPU box: ncu -i needs no GPU.
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * scale.get(unit, 1)


def ncu_raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def ncu_stalls(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    out = {}
    cur, hdr = None, None
    for row in csv.reader(io.StringIO(txt)):
        if len(row) >= 2 and row[0] == "Kernel Name":
            cur, hdr = row[1], None
            out[cur] = {"stalls": collections.Counter(), "ops": collections.Counter(), "samples": 0}
            continue
        if cur is None:
            continue
        if hdr is None:
            hdr = row
            continue
        d = dict(zip(hdr, row))
        try:
            n = int(d["# Samples"])
        except (KeyError, ValueError):
            continue
        tok = d["Source"].split()
        op = tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")
        out[cur]["ops"][op] += n
        out[cur]["samples"] += n
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v:
                out[cur]["stalls"][k] += int(v)
    return out


def summarize_rep(rep, tag, traffic_for=None):
    hdr, units, data = ncu_raw(rep)
    stalls = ncu_stalls(rep)
    written = []
    for r in data:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        kname = d["Kernel Name"]
        short = kname.split("(")[0].replace("void ", "").replace("pmc::", "").replace("<", "_").replace(">", "").replace("(int)", "").replace(", ", "x")
        lines = ["# ncu --set full: %s" % kname, "",
                 "Source: `gpurun_out/%s` (scratch), captured with `--clock-control none --import-source on`; one launch."
                 % os.path.basename(rep), "", "| metric | value | unit |", "|---|---|---|"]
        for k in WANT:
            if k in d:
                lines.append("| %s | %s | %s |" % (k, d[k], u[k]))
        rd = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"])
        wr = to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
        lines.append("| dram bytes per launch (read + write) | %.4g | byte |" % (rd + wr))
        st = None
        base = kname.split("(")[0].split("<")[0].replace("void ", "").replace("pmc::", "")
        for name, val in stalls.items():
            if name.split("(")[0].split("<")[0].replace("void ", "").replace("pmc::", "") == base:
                st = val
        if st and st["samples"]:
            lines += ["", "Warp-stall sampling (all samples = %d):" % st["samples"], "", "| reason | share |", "|---|---|"]
            for k, v in st["stalls"].most_common(8):
                lines.append("| %s | %.1f %% |" % (k, 100.0 * v / st["samples"]))
            lines += ["", "Samples by SASS opcode:", "", "| opcode | share |", "|---|---|"]
            for k, v in st["ops"].most_common(8):
                lines.append("| %s | %.1f %% |" % (k, 100.0 * v / st["samples"]))
        path = os.path.join(PROF, "%s_ncu_%s.md" % (tag, short))
        with open(path, "w") as fh:
            fh.write("\n".join(lines) + "\n")
        written.append(path)
        if traffic_for and traffic_for in kname:
            grid_rows = None
            with open(os.path.join(PROF, "k1_ncu_traffic.json"), "w") as fh:
                json.dump({"kernel": kname, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                           "rows": 10_000_000, "capture": os.path.basename(rep),
                           "gpu_time_ms": float(d["gpu__time_duration.sum"].replace(",", "")) *
                           (1e-6 if u["gpu__time_duration.sum"] in ("ns", "nsecond") else
                            1e-3 if u["gpu__time_duration.sum"] in ("us", "usecond") else 1.0)}, fh, indent=1)
    return written


def summarize_launches(tag):
    path = os.path.join(OUT, "launches_%s.csv" % tag)
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    launches = []
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        ms = v * 1e-6 if r[ui] in ("ns", "nsecond") else v * 1e-3 if r[ui] in ("us", "usecond") else v
        launches.append((int(r[ii]), r[ki], ms))
    agg = collections.OrderedDict()
    for _, k, ms in launches:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    ours = {k: v for k, v in agg.items() if "pmc::" in k and "mb_d" not in k}
    tot_ours = sum(v[1] for v in ours.values())
    lines = ["# Launch list of `python bench.py --steps 4 --warmup 3 --e2e-rows 1000000` under ncu", "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none` (one pass per kernel, serialised, cold caches:",
             "compare SHARES, not absolutes).  %d launches captured.  The torch kernels are the synthetic-sample" % len(launches),
             "generation before the timed region; a timed step launches exactly k1_prepare, k1_mma_prepare, k1_mma_eval",
             "(the matrix-instruction form, which does the work) and k1_fast_eval / k1_mixture_eval (the DFMA forms, which",
             "return at once unless k1_prepare's flags hand them the launch).", "",
             "| kernel | launches | total ms | share of the step kernels (k1_*) |", "|---|---|---|---|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = ("%.1f %%" % (100 * ms / tot_ours) if k in ours else
                 "(FP64 roofline probe, after the timed region)" if "mb_d" in k else "(torch, synthetic-sample setup)")
        lines.append("| `%s` | %d | %.3f | %s |" % (k[:90], n, ms, share))
    # the step launches in order: find the runs of prepare -> fast -> exact
    lines += ["", "First launches of this library's kernels in order (ms):", ""]
    seq = [(k.split("(")[0].replace("void pmc::", ""), ms) for _, k, ms in launches if "pmc::" in k][:24]
    lines.append(", ".join("%s %.3f" % (k, ms) for k, ms in seq))
    out = os.path.join(PROF, "%s_launches.md" % tag)
    with open(out, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return out


def main():
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    done = [summarize_launches(tag)]
    done += summarize_rep(os.path.join(OUT, "k1_full_%s.ncu-rep" % tag), tag + "_bench", traffic_for="k1_mma_eval")
    done += summarize_rep(os.path.join(OUT, "update_full_%s.ncu-rep" % tag), tag + "_update")
    for name in ("configs_%s.log" % tag, "bench_%s.log" % tag, "bench_ref_%s.log" % tag):
        src = os.path.join(OUT, name)
        if os.path.exists(src):
            with open(src) as fh, open(os.path.join(PROF, name.replace(".log", ".jsonl")), "w") as out:
                out.writelines(ln for ln in fh if ln.startswith("{"))
            done.append(os.path.join(PROF, name.replace(".log", ".jsonl")))
    print("\n".join(done))


if __name__ == "__main__":
    main()
