"""Prints, for every golden case and both cores, the worst error / tolerance ratio per output key
(tolerance = max(1e-4, 2 * ||ref32 - ref64||_inf)).  Run on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import CASES, OUT_KEYS, linf, load_case, load_params
from test_render_gpu import _build, _run_kernel

impls = sys.argv[1:] or ["ffma", "tcgen05"]
for impl in impls:
    worst = {}
    for name in CASES:
        meta, inp, r32, r64 = load_case(name)
        if meta["n_importance"] > 0:
            keys = ("color_fine", "weight_sum", "gradient_error", "surface_loss")
        else:
            keys = OUT_KEYS
        P, r = _build(meta, impl)
        out = _run_kernel(r, inp, meta)
        row = []
        for k in keys:
            err = linf(out[k], r64[k]); floor = linf(r32[k], r64[k]); tol = max(1e-4, 2 * floor)
            row.append((err / tol, k, err, floor))
            worst[k] = max(worst.get(k, 0), err / tol)
        row.sort(reverse=True)
        print(f"{impl:8s} {name:18s} " + "  ".join(f"{k}:{e:.1e}/{f:.1e}({r:.2f})" for r, k, e, f in row[:4]))
    print(f"{impl:8s} WORST err/tol per key:", {k: round(v, 3) for k, v in sorted(worst.items(), key=lambda kv: -kv[1])[:6]})
