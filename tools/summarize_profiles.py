"""Turn the raw ncu artefacts in gpurun_out/ into the committed summaries under profiles/ (run here, no GPU needed).

    python tools/summarize_profiles.py r01
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__cycles_elapsed.max",
        "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_elapsed.max.per_second",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(tag, steps=2, src_name="launches.csv", out_name="launch_list_summary.md", title=None):
    src = os.path.join(GP, src_name)
    if not os.path.exists(src):
        return
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    n = 0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    with open(os.path.join(OUT, f"{tag}_{out_name}"), "w") as f:
        f.write((title or f"# {tag}: kernels of the timed region of `python bench.py --steps {steps} --warmup 3` under ncu") + "\n\n"
                "`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised, no power cap:\n"
                "compare SHARES, not absolutes).  Raw list: gpurun_out/launches.csv (scratch).\n\n"
                f"{n} launches, {tot / 1e3:.2f} ms kernel time over {steps} steps = {tot / steps / 1e3:.2f} ms/step\n\n"
                "| share | us/step | launches/step | avg us | kernel |\n|---|---|---|---|---|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {t / tot * 100:.2f}% | {t / steps:.1f} | {c / steps:.1f} | {t / c:.1f} | `{k[:120]}` |\n")
    print("wrote launch summary:", n, "launches")


def raw(tag, rep):
    src = os.path.join(GP, rep + ".ncu-rep")
    pre = os.path.join(GP, rep + "_raw.csv")  # exported on the GPU box by tools/profile_cmds.sh
    if os.path.exists(pre):
        out = open(pre).read()
    elif os.path.exists(src):
        out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [i for i, h in enumerate(hdr) if h in KEYS or h == "Kernel Name"]
    with open(os.path.join(OUT, f"{tag}_{rep}_ncu_full.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in cols])
        w.writerow([units[i] for i in cols])
        for r in data:
            w.writerow([re.sub(r"\(CUtensorMap.*", "", r[i]) if hdr[i] == "Kernel Name" else r[i] for i in cols])
    print("wrote", rep, len(data), "kernel instances")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    launches(tag, steps=3, src_name="train_launches.csv", out_name="train_launch_list_summary.md",
             title=f"# {tag}: every kernel of `python tools/bench_train.py --batch 1 --steps 1 --warmup 1 --layers 2 4` under ncu "
                   "(whole process: model init + 3 train-step passes over 2 double + 4 single blocks; per-step columns = totals / 3)")
    launches(tag, steps=2, src_name="vae_launches.csv", out_name="vae_launch_list_summary.md",
             title=f"# {tag}: every kernel of `python tools/bench_vae.py --steps 1 --warmup 1` under ncu (2 decodes of one 1024px image "
                   "+ model init; per-step columns = totals / 2)")
    launches(tag, steps=1, src_name="cn_launches.csv", out_name="controlnext_launch_list_summary.md",
             title=f"# {tag}: one ControlNeXt net on a 1024x1024 hint (`tools/profile_controlnext.py`, profiled pass only) under ncu")
    launches(tag, steps=1, src_name="lc_train_launches.csv", out_name="lightcontrol_train_launch_list_summary.md",
             title=f"# {tag}: one LightControl train step (`X2I_NCU=1 python tools/bench_lightcontrol_train.py --steps 1 --warmup 1`: FLUX-dev, 19 "
                   "trainable ControlNeXt nets, 1024px, B = 1, VAE encode included) under ncu")
    launches(tag, steps=2, src_name="mllm_launches.csv", out_name="mllm_prefill_launch_list_summary.md",
             title=f"# {tag}: every kernel of `python tools/bench_mllm.py --steps 1 --warmup 1` under ncu (2 prefills of one 512-token prompt through the "
                   "Qwen2.5-VL-3B text decoder + Proj7Exp, + model init; per-step columns = totals / 2)")
    for rep in ("prof_attn", "prof_gemm", "prof_rowwise", "prof_bwd", "prof_vae", "prof_train"):
        raw(tag, rep)
