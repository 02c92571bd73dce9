"""TEST INFRASTRUCTURE — loader that executes the *reference's own* function bodies.

Only `tests/golden/make_golden.py` (run in the build container, where
/root/reference is mounted read-only) and the optional `-m "not gpu"`
cross-check tests import this.  Nothing on the product path may.

The reference scripts cannot run as files on a modern stack:
`torch.potrf` / `torch.gesv` were removed, matplotlib/xlrd are absent, and the
KIN40K workbook is not in the repo.  The *definitions* do run under a two-line
compatibility shim, which is what this module provides:

  torch.potrf(A)      -> upper Cholesky factor   (0.4-era default, KF:26)
  torch.gesv(B, A)    -> (solve(A, B), None)     (LU solve, KF:27-28)

Nothing from the reference is copied into this repository: the def-block and
the loop-body statements are read from /root/reference at call time and
`exec`-ed, so the goldens are outputs of the reference's own source text run
with float64 tensors.
"""
import os
import sys
import textwrap
import types

REF_ROOT = os.environ.get("GPS_REFERENCE_ROOT", "/root/reference")

# script -> (file name, last line of the def-block, i.e. the line before
# `import pandas`)
SCRIPTS = {
    "KF": ("kin40k-FULL-compare.py", 134),
    "K20": ("KIN40K-COMPARE-ALL-FITC-20.py", 132),
    "SF": ("SIMPLE-DATA FULL-comapre.py", 119),
    "SC": ("SIMPLE-FITC--comapre.py", 119),
}


def available():
    return all(os.path.isfile(os.path.join(REF_ROOT, f)) for f, _ in SCRIPTS.values())


def _read_lines(script):
    fname, _ = SCRIPTS[script]
    with open(os.path.join(REF_ROOT, fname), "r", encoding="utf-8", errors="replace") as fh:
        return fh.read().split("\n")


def _install_shims():
    import torch

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if not hasattr(torch, "potrf"):
        torch.potrf = lambda A, upper=True: torch.linalg.cholesky(A).mT
    if not hasattr(torch, "gesv"):
        torch.gesv = lambda B, A: (torch.linalg.solve(A, B), None)


def load_namespace(script):
    """exec the def-block of one reference script; returns its globals dict.

    The namespace is given float64 defaults: the reference is written for
    `dtype = torch.FloatTensor` (KF:205) but the oracle and the build are
    float64, so `dtype = torch.DoubleTensor` is injected and
    the torch default dtype is switched for the duration of every call made
    through `run_block`.
    """
    import torch

    _install_shims()
    lines = _read_lines(script)
    _, last = SCRIPTS[script]
    src = "\n".join(lines[:last])
    ns = {"__name__": "reference_" + script}
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        exec(compile(src, SCRIPTS[script][0], "exec"), ns)
    finally:
        torch.set_default_dtype(prev)
    ns["dtype"] = torch.DoubleTensor
    return ns


def run_block(ns, script, first, last, _default_dtype=None, **inject):
    """exec reference lines [first, last] (1-based, inclusive) verbatim in `ns`.

    The lines are loop-body statements (indented in the file); they are
    dedented as a block and executed with float64 as torch's default dtype so
    `torch.eye(n)` / `torch.tensor([...])` inside them are float64.
    """
    import torch

    ns.update(inject)
    lines = _read_lines(script)[first - 1:last]
    # drop pure-comment lines: some are at column 0 inside an indented block,
    # which would defeat dedent.
    lines = [ln for ln in lines if not ln.lstrip().startswith("#")]
    src = textwrap.dedent("\n".join(lines))
    prev = torch.get_default_dtype()
    torch.set_default_dtype(_default_dtype or torch.float64)
    try:
        exec(compile(src, "%s:%d-%d" % (SCRIPTS[script][0], first, last), "exec"), ns)
    finally:
        torch.set_default_dtype(prev)
    return ns
