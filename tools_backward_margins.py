import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
import test_backward_gpu as T
worst_all = {}
for rep in range(3):
    for name in ("cfg1_n16_m0", "cfgd_n16_m4_D8"):
        for impl in ("ffma", "tcgen05"):
            meta, c, w = T._inputs(name, 64)
            wr = T._check(meta, c, w, T.ADJ_KEYS, 0.3, impl=impl, seed=rep)
            worst_all[(name, impl, rep)] = wr
    meta, c, w = T._inputs("cfgd_n16_m4_D8", 50, n_inst=2)
    for impl in ("ffma", "tcgen05"):
        worst_all[("ragged", impl, rep)] = T._check(meta, c, w, ["color_fine", "weight_sum", "gradient_error", "weights", "gradients", "raw_color"], 1.0, impl=impl, seed=rep)
    meta, c, w = T._inputs("cfgd_n16_m4_D8", 96)
    t_rand = torch.rand(96, 1, device="cuda") - 0.5
    worst_all[("hier", "auto", rep)] = T._check(meta, c, w, ["color_fine", "weight_sum", "gradient_error"], 0.5, n_importance=4, t_rand=t_rand, seed=rep)
for k, v in sorted(worst_all.items(), key=lambda kv: -kv[1][0])[:12]:
    print(f"{v[0]:.3f} of tol  {v[1]:45s} {k}")
