#!/usr/bin/env python
"""Turn the artefacts of one `profiles/run_gpu.sh <tag>` call (gpurun_out/<tag>_*) into small tracked summaries:

    profiles/<tag>_bench.json        the bench line (pretty-printed)
    profiles/<tag>_launches.csv      ncu launch list: kernel, duration (us)   [cold-cache, serialised: shares only]
    profiles/<tag>_ncu_summary.md    per kernel: duration, DRAM bytes, DRAM / SM / L2 throughput %, occupancy, registers,
                                     top stall reasons  (from `ncu --set full`, read with --page raw)
    profiles/ncu_traffic.json        kernel -> dram bytes per launch (read + write), consumed by bench.py `roofline.traffic`

Runs here (no GPU): only reads files.   python profiles/summarize.py r1a
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

# bench.py stage name -> kernels that belong to it
STAGES = {
    "blur": ["blur15"], "pack": ["pack_masks"], "prep_setup": ["prep_setup"], "prep": ["prep_main"], "heat_tables": ["heat_prefix", "heat_consts", "heat_resize"], "grid_heat_pool": ["mask_rows", "mask_grid", "mask_area"], "pool_score": ["pool_score"], "score_select": ["score_select", "score_text"], "iou": ["iou_zero", "iou_kernel"],
    "mask_pool": ["mask_pool"], "token_mask_fuse": ["token_mask_fuse"], "rle_to_bits": ["rle_"], "gem_token": ["gem_"],
}


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("hgl::", "")
    return n.strip()


def stage_of(kname):
    for st, keys in STAGES.items():
        if any(k in kname for k in keys):
            return st
    return "other"


def launches(tag):
    path = os.path.join(OUT, f"{tag}_launches.csv")
    if not os.path.exists(path):
        return None
    text = open(path).read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    out = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        out.append((short(r["Kernel Name"]), us))
    with open(os.path.join(PROF, f"{tag}_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare SHARES only)\n")
        f.write("launch,kernel,stage,duration_us\n")
        for i, (k, us) in enumerate(out):
            f.write(f"{i},{k},{stage_of(k)},{us:.2f}\n")
        tot = sum(us for _, us in out)
        f.write("# share of the profiled launches per bench stage\n")
        agg = {}
        for k, us in out:
            agg[stage_of(k)] = agg.get(stage_of(k), 0.0) + us
        for st, us in sorted(agg.items(), key=lambda x: -x[1]):
            f.write(f"# {st},{us:.1f} us,{us / tot:.3f}\n")
    return out


def ncu_full(tag):
    rep = os.path.join(OUT, f"{tag}_prof.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, default=float("nan")):
        i = col.get(name)
        if i is None or r[i] in ("", "n/a"):
            return default
        try:
            return float(r[i].replace(",", ""))
        except ValueError:
            return default

    def bytes_of(r, name):
        v = get(r, name, 0.0)
        u = units[col[name]] if name in col else "byte"
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    def us_of(r):
        v = get(r, "gpu__time_duration.sum")
        u = units[col["gpu__time_duration.sum"]]
        return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u.replace("second", "s").replace("usecond", "us"), 1e-3 if u.startswith("n") else 1)

    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    lines = [f"# ncu --set full summary, tag `{tag}` (one pass of the path; cold-ish caches, clock-control none)", "",
             "| kernel | us | DRAM rd MB | DRAM wr MB | DRAM GB/s | DRAM % | SM % | L2 % | warps active % | regs | grid x block | top stalls (cycles per issue) |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    traffic = {}
    tensor_rows = []
    for r in body:
        name = short(r[col["Kernel Name"]])
        if get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) > 0.0:
            tensor_rows.append(
                f"| {name} | {us_of(r):.1f} | {get(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.2f} | "
                f"{get(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.2f} | "
                f"{get(r, 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active'):.2f} | "
                f"{get(r, 'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active'):.3f} | "
                f"{get(r, 'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active'):.3f} | "
                f"{get(r, 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.2f} | "
                f"{get(r, 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active'):.2f} |")
        us = us_of(r)
        rd, wr = bytes_of(r, "dram__bytes_read.sum"), bytes_of(r, "dram__bytes_write.sum")
        st = sorted(((get(r, s, 0.0), s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for s in stalls), reverse=True)[:3]
        lines.append(f"| {name} | {us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / us / 1e3:.0f} | "
                     f"{get(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                     f"{get(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                     f"{get(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                     f"{get(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
                     f"{get(r, 'launch__registers_per_thread'):.0f} | {get(r, 'launch__grid_size'):.0f} x {get(r, 'launch__block_size'):.0f} | "
                     + ", ".join(f"{n} {v:.1f}" for v, n in st) + " |")
        stg = stage_of(name)
        traffic.setdefault(stg, {"bytes": 0.0, "kernels": {}})
        traffic[stg]["kernels"][name] = rd + wr
    if tensor_rows:
        lines += ["", "Tensor pipe (tcgen05) kernels -- % of peak: `sm__pipe_tensor_cycles_active` (sustained active / elapsed), its hmma sub-pipe, "
                  "`sm__inst_executed_pipe_tensor_subpipe_hmma`, `sm__inst_executed_pipe_tmem`, `sm__mem_tensor_cycles_active`, "
                  "`sm__inst_executed_pipe_uniform`:", "",
                  "| kernel | us | pipe_tensor active % | pipe_tensor elapsed % | hmma sub-pipe active % | inst pipe_tensor hmma % | inst pipe_tmem % | mem_tensor active % | inst pipe_uniform % |",
                  "|---|---|---|---|---|---|---|---|---|"] + tensor_rows
    for stg, d in traffic.items():
        d["bytes"] = sum(d["kernels"].values())
    with open(os.path.join(PROF, f"{tag}_ncu_summary.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(os.path.join(PROF, "ncu_traffic.json"), "w") as f:
        json.dump({"tag": tag, "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, summed over the kernels of each bench stage",
                   **{k: v["bytes"] for k, v in traffic.items()}, "detail": {k: v["kernels"] for k, v in traffic.items()}}, f, indent=1)
    print("\n".join(lines))


def main():
    tag = sys.argv[1]
    b = os.path.join(OUT, f"{tag}_bench.json")
    if os.path.exists(b):
        txt = open(b).read().strip().splitlines()
        for ln in txt:
            if ln.startswith("{"):
                with open(os.path.join(PROF, f"{tag}_bench.json"), "w") as f:
                    json.dump(json.loads(ln), f, indent=1)
    for extra in ("pytest.log", "smoke.log", "gpu.csv", "membw.log"):
        p = os.path.join(OUT, f"{tag}_{extra}")
        if os.path.exists(p):
            with open(os.path.join(PROF, f"{tag}_{extra}"), "w") as f:
                f.write(open(p).read()[-4000:])
    launches(tag)
    ncu_full(tag)


if __name__ == "__main__":
    main()
