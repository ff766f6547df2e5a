"""Summarise ncu artefacts from gpurun_out/ into small text files under profiles/ (run on the CPU box)."""
import collections, csv, re, subprocess, sys
from pathlib import Path
OUT = Path("profiles")

def launch_list(path, passes, tag):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        us = v / 1000 if u.startswith("n") else (v if u.startswith("u") else v * 1000)
        agg[name][0] += 1; agg[name][1] += us
    tot = sum(v[1] for v in agg.values())
    with open(OUT / f"{tag}_launch_summary.txt", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# source: {path}; {passes} model passes in the run; total {tot/1000:.2f} ms = {tot/passes/1000:.2f} ms/pass\n")
        f.write(f"{'share':>7} {'us/pass':>10} {'launches/pass':>14}  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"{100*v[1]/tot:6.2f}% {v[1]/passes:10.1f} {v[0]/passes:14.1f}  {k}\n")
    print(open(OUT / f"{tag}_launch_summary.txt").read()[:2500])

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]

def full(rep, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(OUT / f"{tag}.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ; report {rep}\n# kernel: {vals[hdr.index('Kernel Name')]}\n")
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS or any(h.endswith(k) for k in KEYS) or "pipe_tensor" in h and "pct_of_peak_sustained_elapsed" in h and "avg" in h:
                f.write(f"{h} [{u}] = {v}\n")
    print(open(OUT / f"{tag}.txt").read())

def traffic(rep):
    """dram__bytes_read.sum + dram__bytes_write.sum of the first captured launch, in bytes."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        j = hdr.index(k)
        v = float(vals[j].replace(",", ""))
        tot += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[j]]
    return tot, vals[hdr.index("Kernel Name")]


if __name__ == "__main__":
    import json
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02l"      # file prefix under gpurun_out/
    # bench.py --steps 2 --warmup 1 --no-extras --no-cpu runs 3 warm-up + 2 timed + 2 stage + 2 per-kernel + 2+2 e2e passes = 13
    launch_list(f"gpurun_out/{tag}_launches.csv", 13, "r02_bench")
    import shutil
    shutil.copy(f"gpurun_out/{tag}_launches.csv", OUT / "r02_launches_bench.csv")
    # bench-line kernel key -> capture (the capture tools run the same instantiation at the same shape as the bench step)
    caps = {"dcn3d_kernel<64>": "dcn3d", "costvol_fwd": "costvol", "regress_fwd": "regress", "conv3d_s2 32->64": "conv_s2",
            "conv3d kind0 64->32": "conv_kdfused_64x32", "conv3d kind0 32->32": "conv_kdfused_32x32", "conv2d_tc 64->96 d1": "conv2d",
            "conv3d_head 32->1": "head", "conv3d kind2 64->32": "conv_t2"}
    tr = json.loads((OUT / "r02_traffic.json").read_text()) if (OUT / "r02_traffic.json").is_file() else {}   # keep captures of earlier calls
    caps["stem_conv 3->32 s2"] = "stem"
    caps["conv2d_tc 96->32 d1"] = "conv2d_96x32"
    for key, name in caps.items():
        rep = Path(f"gpurun_out/{tag}_{name}.ncu-rep")
        if not rep.is_file():
            continue
        full(str(rep), f"r02_{name}")
        b, kname = traffic(str(rep))
        tr[key] = {"bytes_per_launch": b, "kernel": kname, "source": f"profiles/r02_{name}.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch)"}
    (OUT / "r02_traffic.json").write_text(json.dumps(tr, indent=1))
