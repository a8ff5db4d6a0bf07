"""gpurun_out/parity_table.jsonl (appended by tests/parity.py during `pytest -m gpu`) -> profiles/<tag>_parity_table.md, and the per-depth
error curve gpurun_out/parity_depth_curve.json -> profiles/<tag>_parity_depth_curve.json (+ a short table).

    python tools/summarize_parity.py r02
"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(tag):
    src = os.path.join(ROOT, "gpurun_out", "parity_table.jsonl")
    rows = {}
    for ln in open(src):
        r = json.loads(ln)
        rows[r["name"]] = r  # last run of a test wins
    out = os.path.join(ROOT, "profiles", f"{tag}_parity_table.md")
    with open(out, "w") as f:
        f.write(f"# {tag}: parity of every bf16 path against the fp32 oracle, measured on B200 (`pytest -m gpu`, tests/parity.py)\n\n"
                "`err` = relative Frobenius error of the x2i_b200 CUDA path; `eager bf16` = the reference's own path (oracle modules in bf16, stock\n"
                "PyTorch ops on the same GPU) against the same fp32 ground truth.  Bar: `err <= 1e-2`, or `err <= eager bf16` where the reference's\n"
                "own bf16 path misses 1e-2 on those inputs (random-weight deep stacks amplify rounding noise).  `recorded only` rows are small\n"
                "tensors (< 4096 elements) whose relative error is noise in both paths.\n\n"
                "| comparison | err | eager bf16 | bar met by |\n|---|---|---|---|\n")
        for name, r in rows.items():
            e, y, tol = r["err"], r.get("eager"), r.get("tol", 1e-2)
            how = "1e-2" if e <= tol else ("eager" if (y is not None and e <= y) else ("-" if "recorded only" in name or y is None else "FAIL"))
            ys = "-" if y is None else f"{y:.5f}"
            f.write(f"| {name} | {e:.5f} | {ys} | {how} |\n")
    print("wrote", out, len(rows), "rows")
    dc = os.path.join(ROOT, "gpurun_out", "parity_depth_curve.json")
    if os.path.exists(dc):
        dst = os.path.join(ROOT, "profiles", f"{tag}_parity_depth_curve.json")
        shutil.copy(dc, dst)
        d = json.load(open(dc))
        with open(os.path.join(ROOT, "profiles", f"{tag}_parity_depth_curve.md"), "w") as f:
            f.write(f"# {tag}: relative error of the 76 hooked attention-module outputs through the full 19 + 38 block FLUX transformer\n\n"
                    "1024 px (512 text + 4096 latent tokens), D = 3072, B = 1, random weights N(0, 0.02^2); x2i_b200 and the eager-bf16 reference\n"
                    "path, both against the fp32 oracle on the GPU (tests/test_gpu_parity_full.py).  Full curves: the .json next to this file.\n\n")
            for tagk, v in d.items():
                f.write(f"## {tagk}\n\n| layer | x2i_b200 | eager bf16 |\n|---|---|---|\n")
                n = len(v["layers"])
                pick = sorted(set(list(range(0, n, 6)) + [18, 19, 37, 38, n - 1]))
                for i in pick:
                    f.write(f"| {v['layers'][i]} | {v['x2i_b200'][i]:.5f} | {v['eager_bf16'][i]:.5f} |\n")
                f.write(f"\nmax over all 76: x2i_b200 {max(v['x2i_b200']):.5f}, eager bf16 {max(v['eager_bf16']):.5f}\n\n")
        print("wrote depth curve")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
