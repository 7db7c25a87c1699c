"""Turn the ncu outputs of scripts/gpu_ncu_r2.sh (brought back in gpurun_out/) into the tracked summaries under profiles/.

    python scripts/ncu_summarise.py r2x

reads  gpurun_out/launches_<tag>.csv, gpurun_out/prof_step_<tag>.ncu-rep (or its _raw.csv export), gpurun_out/prof_lanczos5000_<tag>.ncu-rep
writes profiles/<tag>_launches.csv, profiles/<tag>_launches_summary.md, profiles/<tag>_step_ncu_full.md,
       profiles/<tag>_lanczos5000_ncu_full.md, profiles/lanczos_traffic.json
(`ncu -i ... --page raw --csv` runs here: reading a report needs no GPU).
"""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2x"
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("pb::", "").replace("<unnamed>::", "")
    m = re.match(r"([A-Za-z0-9_]+(<[^>]*>)?)", name)
    return m.group(1) if m else name


def raw_page(rep):
    csv_path = rep.replace(".ncu-rep", "_raw.csv")
    if not os.path.exists(rep) and os.path.exists(csv_path):      # the GPU job already exported the raw page
        txt = open(csv_path).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[start], rows[start + 1], rows[start + 2:]


def launches():
    src = os.path.join(OUT, f"launches_{TAG}.csv")
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    n = 0
    for r in rows[start + 1:]:
        if len(r) <= mv or r[mv] in ("", "nan"):
            continue
        k = short(r[kn])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", "")) / 1e3
        n += 1
    tot = sum(v[1] for v in agg.values())
    shutil.copy(src, os.path.join(PROF, f"{TAG}_launches.csv"))
    with open(os.path.join(PROF, f"{TAG}_launches_summary.md"), "w") as f:
        f.write(f"# {TAG} — ncu launch list (first {n} launches of `python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-large-cone`)\n\n")
        f.write("Command: `PROXSDP_B200_LZ_COOP=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv ...` "
                "(cold-cache, serialised: compare shares, not absolute times).  The first ~60 launches are the setup of the\n"
                "warm-up solve (device-side ingest: radix sort, scans, CSR/DCSR build).\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {c} | {t:.1f} | {t / c:.2f} | {t / tot:.3f} |\n")
        f.write(f"\nTotal {tot:.1f} us over {n} launches.\n")
    return agg, tot


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__cluster_size", "smsp__inst_executed.sum", "sm__icc_request_hit_rate.pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def mbytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[unit]


def step_full():
    hdr, units, rows = raw_page(os.path.join(OUT, f"prof_step_{TAG}.ncu-rep"))
    kn = hdr.index("Kernel Name")
    dur, rd, wr = hdr.index("gpu__time_duration.sum"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    hit, lts = hdr.index("lts__t_sector_hit_rate.pct"), hdr.index("lts__throughput.avg.pct_of_peak_sustained_elapsed")
    regs, grid, blk = hdr.index("launch__registers_per_thread"), hdr.index("launch__grid_size"), hdr.index("launch__block_size")
    agg = collections.OrderedDict()
    for r in rows:
        k = short(r[kn])
        a = agg.setdefault(k, dict(n=0, us=0.0, rd=0.0, wr=0.0, hit=0.0, lts=0.0, regs=r[regs], grid=r[grid], blk=r[blk]))
        a["n"] += 1
        a["us"] += float(r[dur]) * {"us": 1.0, "ns": 1e-3, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}.get(units[dur], 1.0)
        a["rd"] += mbytes(r[rd], units[rd]); a["wr"] += mbytes(r[wr], units[wr])
        a["hit"] += float(r[hit]); a["lts"] += float(r[lts])
    with open(os.path.join(PROF, f"{TAG}_step_ncu_full.md"), "w") as f:
        f.write(f"# {TAG} — `ncu --set full --clock-control none --import-source on -s 200 -c 26`: every kernel of 2.6 consecutive PDHG iterations\n\n")
        f.write("Workload: `python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-large-cone` (Max-Cut n = 2000, N = 2 001 000 svec entries, 2000 rows).\n"
                "Each launch is replayed ~39 times with caches flushed between passes: durations are cold-cache and NOT bench numbers;\n"
                "DRAM bytes are per launch (average over the captured launches of that kernel).\n\n")
        f.write("| kernel | captured | avg us | DRAM read MB | DRAM write MB | DRAM GB/s | L2 hit % | lts throughput % | regs | grid x block |\n")
        f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
        for k, a in agg.items():
            n = a["n"]
            us = a["us"] / n
            f.write(f"| {k} | {n} | {us:.2f} | {a['rd'] / n:.3f} | {a['wr'] / n:.3f} | {(a['rd'] + a['wr']) / n / us * 1e3:.0f} | "
                    f"{a['hit'] / n:.1f} | {a['lts'] / n:.1f} | {a['regs']} | {a['grid']} x {a['blk']} |\n")
        f.write("\nReading: the streaming kernels (`k_svec_to_mat<1>`: 24N + 8n^2 = 80 MB algorithmic, of which the 32 MB matrix is written into L2;\n"
                "`k_residual_primal`: 40N = 80 MB) are the HBM-bound ones; everything with < 1 MB of traffic is a latency-bound\n"
                "chain of a few thousand rows (2000 constraint rows on this workload).\n")
    lz = next((a for k, a in agg.items() if k.startswith("k_lanczos_cl3")), None)
    return agg, lz


def lanczos_detail(rep, title, out, note):
    hdr, units, rows = raw_page(rep)
    r = rows[0]
    with open(out, "w") as f:
        f.write(f"# {TAG} — {title}\n\nKernel: `{r[hdr.index('Kernel Name')]}` (replayed ~40 times with cold caches: the duration below is NOT a bench number).\n\n")
        f.write("| metric | value | unit |\n|---|---:|---|\n")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"| {w} | {r[i]} | {units[i]} |\n")
        tot = 0
        st = []
        for i, h in enumerate(hdr):
            if h.startswith(STALLS) and not h.endswith("_not_issued"):
                v = int(float(r[i].replace(",", "")))
                st.append((h[len(STALLS):], v)); tot += v
        f.write(f"\n{note}\n\n## Warp-stall samples ({tot})\n\n| reason | samples | share |\n|---|---:|---:|\n")
        for h, v in sorted(st, key=lambda kv: -kv[1]):
            if v:
                f.write(f"| {h} | {v} | {100.0 * v / max(tot, 1):.1f}% |\n")
    i_rd, i_wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    return mbytes(r[i_rd], units[i_rd]) + mbytes(r[i_wr], units[i_wr])


def main():
    agg, tot = launches()
    step, lz = step_full()
    # the eigsolve kernel of the step capture, in detail: take its first captured launch from the step report
    hdr, units, rows = raw_page(os.path.join(OUT, f"prof_step_{TAG}.ncu-rep"))
    kn = hdr.index("Kernel Name")
    lzrow = next(r for r in rows if "k_lanczos_cl3" in r[kn])
    with open(os.path.join(PROF, f"{TAG}_lanczos_cl3_ncu_full.md"), "w") as f:
        f.write(f"# {TAG} — eigsolve kernel of the headline workload (Max-Cut n = 2000), from the `--set full` step capture\n\n"
                f"Kernel: `{lzrow[kn]}` (replayed ~39 times with cold caches: the duration below is NOT a bench number).\n\n")
        f.write("| metric | value | unit |\n|---|---:|---|\n")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"| {w} | {lzrow[i]} | {units[i]} |\n")
        st = [(h[len(STALLS):], int(float(lzrow[i].replace(",", "")))) for i, h in enumerate(hdr) if h.startswith(STALLS) and not h.endswith("_not_issued")]
        tot_s = sum(v for _, v in st)
        f.write("\nDRAM traffic of the launch = the 32 MB matrix read ONCE (it stays in L2 for the ~27 mat-vecs of the launch, and 7 of every 17 slab rows\n"
                "stay in shared memory); algorithmic bytes of the same launch: mat-vecs x (8 n^2 + 16 n) ~ 0.87 GB.\n")
        f.write(f"\n## Warp-stall samples ({tot_s})\n\n| reason | samples | share |\n|---|---:|---:|\n")
        for h, v in sorted(st, key=lambda kv: -kv[1]):
            if v:
                f.write(f"| {h} | {v} | {100.0 * v / max(tot_s, 1):.1f}% |\n")
    t5000 = lanczos_detail(os.path.join(OUT, f"prof_lanczos5000_{TAG}.ncu-rep"),
                           "`ncu --set full -k regex:k_lanczos_cl3 -s 2 -c 1 python scripts/lz_large.py 5000`: eigsolve on a side-5000 cone (200 MB matrix)",
                           os.path.join(PROF, f"{TAG}_lanczos5000_ncu_full.md"),
                           "25 mat-vecs x 200 MB = 5.0 GB algorithmic; the DRAM bytes above are what the launch really moved (the matrix does not fit L2,\n"
                           "so nearly every mat-vec streams it from HBM; the TMA prefetch `UBLKPF.L2` runs a 4-row window ahead of the loads).")
    if lz is not None:
        per_launch = (lz["rd"] + lz["wr"]) / lz["n"] * 1e6
        json.dump({"kernel": "k_lanczos_cl3", "dram_bytes_per_launch": per_launch,
                   "source": f"profiles/{TAG}_step_ncu_full.md (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, average of {lz['n']} launches)",
                   "side5000_dram_bytes_per_launch": t5000 * 1e6},
                  open(os.path.join(PROF, "lanczos_traffic.json"), "w"), indent=1)
    print("written:", [p for p in sorted(os.listdir(PROF)) if p.startswith(TAG)] + ["lanczos_traffic.json"])


if __name__ == "__main__":
    main()
