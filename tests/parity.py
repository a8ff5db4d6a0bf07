"""Shared parity bookkeeping of the -m gpu tests.

Every bf16 comparison records TWO numbers against the same fp32 oracle ground truth:
  err    relative (Frobenius) error of the x2i_b200 CUDA path,
  eager  relative error of the reference's own bf16 path (the oracle modules run in bf16 with torch eager ops on the GPU),
and asserts ``err <= tol`` (1e-2) or, only where the reference's bf16 path itself misses ``tol``,
``err <= eager`` -- never a scaled or floating bound.  The records are appended to ``gpurun_out/parity_table.jsonl``; the
committed copy of the table is ``profiles/r02_parity_table.md`` (tools/summarize_parity.py).
"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABLE = os.path.join(ROOT, "gpurun_out", "parity_table.jsonl")
TOL = 1e-2


def record(name, err, eager=None, tol=TOL, **extra):
    row = dict(name=name, err=float(err), eager=None if eager is None else float(eager), tol=tol, **extra)
    try:
        os.makedirs(os.path.dirname(TABLE), exist_ok=True)
        with open(TABLE, "a") as f:
            f.write(json.dumps(row) + "\n")
    except OSError:
        pass
    return row


def check(name, err, eager, tol=TOL, **extra):
    """Record (err, eager) and enforce the bar: tol, or the reference's own bf16 error where that path misses tol."""
    record(name, err, eager, tol, **extra)
    err, eager = float(err), float(eager)
    print(f"[parity] {name}: x2i_b200 {err:.5f}  eager-bf16 reference path {eager:.5f}  (tol {tol:g})")
    if err <= tol:
        return
    assert err <= eager, f"{name}: rel err {err:.5f} exceeds both tol {tol:g} and the reference's own bf16 path ({eager:.5f})"
