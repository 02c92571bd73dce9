"""Generate tests/golden/*.npz by running the REFERENCE's own source text.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

For every case the reference's def-block (ARD, chol_solve, Q, crps, logs,
cal_mean_and_cov, spgp_cal_mean_and_cov, SMSE, trivial_loss) and the
loop-body statements of the scripts are `exec`-ed verbatim from the reference
files through `oracle/ref_shim.py`, with float64 tensors, on seeded synthetic
inputs.  The inputs and every output (objective, autograd gradients, LOO
mean/variance, predictions, test metrics) are stored, so the tests never need
the reference again.  The reference holds no golden vectors of its own
(SURVEY.md §4, §8c): these files are the pin.

Reference line ranges replayed (file:first-last):
  full GP   CRPS  KF:239-245   NLML KF:329-334   logs KF:416-424   4-fold DSS KF:499-538
            predict + metrics  KF:267-292
  FITC      CRPS  K20:222-234  NLML K20:329-340  logs K20:434-447  4-fold DSS K20:538-582  kc K20:669-714
            predict + metrics  K20:270-296
  SIMPLE    data  SF:161-181 / SC:161-181 (torch.manual_seed(0));
            loops SF:206-213, SF:291-296, SF:384-392 / SC:208-220, SC:321-333, SC:441-452
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from gpscore_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def T(a, grad=False):
    t = torch.tensor(np.asarray(a, dtype=np.float64), dtype=torch.float64)
    if grad:
        t.requires_grad_(True)
    return t


def leaves(theta, d_b, U=None):
    """Leaf tensors shaped like the scripts': para_k[1], para_l[1,D] or [1], para_noise[1]."""
    a, b, c = theta[0], theta[1:-1], theta[-1]
    para_k = T([a], True)
    if d_b == 1:
        para_l = T([b[0]], True)            # isotropic 1-element (K20:422, SF:199)
    else:
        para_l = T(b.reshape(1, -1), True)  # [1, D] (KF:226)
    para_noise = T([c], True)
    out = dict(para_k=para_k, para_l=para_l, para_noise=para_noise)
    if U is not None:
        out["inducing_x"] = T(U, True)
    return out


def grads_of(obj, lv):
    for v in lv.values():
        if v.grad is not None:
            v.grad = None
    obj.sum().backward()
    return {k: v.grad.detach().numpy().copy() for k, v in lv.items()}


FULL_BLOCKS = {  # script -> score -> (first, last, name of result)
    "KF": {"crps": (239, 245, "CRPS_ave"), "nlml": (329, 334, "Neg_logL"), "logs": (416, 424, "logs_ave")},
    "SF": {"crps": (206, 214, "CRPS_ave"), "nlml": (291, 296, "Neg_logL"), "logs": (384, 392, "logs_ave")},
}
FITC_BLOCKS = {
    "K20": {"crps": (222, 234, "CRPS_ave"), "nlml": (329, 340, "Neg_logL"), "logs": (434, 447, "logs_ave")},
    "SC": {"crps": (208, 220, "CRPS_ave"), "nlml": (321, 333, "Neg_logL"), "logs": (441, 452, "logs_ave")},
}
# SF has its metric lines commented out (SF:245-260): its prediction lines are replayed and
# the identical metric statements are taken from KF:276-292, executed in SF's namespace.
PREDICT = {"KF": [("KF", 267, 292)], "K20": [("K20", 270, 296)],
           "SF": [("SF", 236, 242), ("KF", 276, 292)], "SC": [("SC", 252, 278)]}


def run_case(name, script, X, y, Xs, ys, theta, d_b, U=None):
    fitc = U is not None
    ns = ref_shim.load_namespace(script)
    rec = dict(X=X, y=y, Xs=Xs, ys=ys, theta=np.asarray(theta, dtype=np.float64),
               d_b=np.int64(d_b), script=np.array(script))
    if fitc:
        rec["U"] = U
    common = dict(train_x=T(X), train_y=T(y), test_x=T(Xs), test_y=T(ys),
                  num_train=X.shape[0], num_test=Xs.shape[0])
    blocks = (FITC_BLOCKS if fitc else FULL_BLOCKS)[script]
    for score, (first, last, res) in blocks.items():
        lv = leaves(theta, d_b, U)
        ref_shim.run_block(ns, script, first, last, **common, **lv)
        obj = ns[res]
        rec["obj_" + score] = np.float64(obj.detach().numpy().reshape(-1)[0])
        g = grads_of(obj, lv)
        rec["grad_k_" + score] = g["para_k"]
        rec["grad_l_" + score] = g["para_l"]
        rec["grad_noise_" + score] = g["para_noise"]
        if fitc:
            rec["grad_u_" + score] = g["inducing_x"]
        if score in ("crps", "logs"):
            rec["loo_mean_" + score] = ns["mean_term"].detach().numpy().copy()
            rec["loo_var_" + score] = ns["cov_term"].detach().numpy().copy()
    # 4-fold DSS objective (KF:499-538); the script sizes every fold with index1, so it needs 4 | N
    if script == "KF" and X.shape[0] % 4 == 0:
        lv = leaves(theta, d_b, U)
        ref_shim.run_block(ns, "KF", 499, 538, **common, **lv)
        obj = ns["dss_ave"]
        rec["obj_dss"] = np.float64(obj.detach().numpy().reshape(-1)[0])
        g = grads_of(obj, lv)
        rec["grad_k_dss"] = g["para_k"]
        rec["grad_l_dss"] = g["para_l"]
        rec["grad_noise_dss"] = g["para_noise"]
    # FITC 4-fold DSS (K20:538-582) and block-CRPS "kc" (K20:669-714); same 4 | N requirement
    if script == "K20" and X.shape[0] % 4 == 0:
        for score, (first, last, res) in {"dss": (538, 582, "dss_ave"), "kc": (669, 714, "kc_ave")}.items():
            lv = leaves(theta, d_b, U)
            ref_shim.run_block(ns, "K20", first, last, **common, **lv)
            obj = ns[res]
            rec["obj_" + score] = np.float64(obj.detach().numpy().reshape(-1)[0])
            g = grads_of(obj, lv)
            rec["grad_k_" + score] = g["para_k"]
            rec["grad_l_" + score] = g["para_l"]
            rec["grad_noise_" + score] = g["para_noise"]
            rec["grad_u_" + score] = g["inducing_x"]
    # prediction + test metrics with the same hyper-parameters
    lv = leaves(theta, d_b, U)
    with torch.no_grad():
        ns["sigma_noise_sq"] = torch.exp(lv["para_noise"])
        for src, first, last in PREDICT[script]:
            ref_shim.run_block(ns, src, first, last, **common, **lv)
    mean_name, var_name, pre = {
        "KF": ("y_mean_crps", "y_cov_crps_diag", "crps_"),
        "SF": ("y_mean_crps", "y_cov_crps_diag", "crps_"),
        "K20": ("mean_crps", "cov_crps_diag", ""),
        "SC": ("mean_crps", "cov_crps_diag", ""),
    }[script]
    rec["pred_mean"] = ns[mean_name].detach().numpy().copy()
    rec["pred_var"] = ns[var_name].detach().numpy().copy()
    if script in ("KF", "SF"):
        rec["m_mse"] = np.float64(ns["crps_mse"])
        rec["m_smse"] = np.float64(ns["smse_crps"])
        rec["m_logs"] = np.float64(ns["crps_test_logs"])
        rec["m_crps"] = np.float64(ns["crps_test_crps"])
        rec["m_msll"] = np.float64(ns["MSLL_crps"])
    else:
        rec["m_mse"] = np.float64(ns["mse_crps"])
        rec["m_smse"] = np.float64(ns["smse_crps"])
        rec["m_logs"] = np.float64(ns["logs_crps"])
        rec["m_crps"] = np.float64(ns["crps_crps"])
        rec["m_msll"] = np.float64(ns["MSLL_crps"])
    rec["m_coverage"] = np.float64(ns["res"])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **rec)
    print("%-28s obj crps %.12g  nlml %.12g  logs %.12g  -> %s" % (
        name, rec["obj_crps"], rec["obj_nlml"], rec["obj_logs"], os.path.basename(path)))
    return rec


def simple_data(script):
    """Replay the toy data generation of SF:161-181 / SC:161-181 with torch.manual_seed(0)."""
    ns = ref_shim.load_namespace(script)
    torch.manual_seed(0)
    # the toy data is drawn in float32 exactly as the script does (dtype = FloatTensor, SF:165),
    # then widened to float64 once; every evaluation after that is float64.
    ref_shim.run_block(ns, script, 161, 181, _default_dtype=torch.float32, p=0)
    ns["dtype"] = torch.DoubleTensor
    f64 = lambda t: t.detach().double().numpy().copy()
    return f64(ns["train_x"]), f64(ns["train_y"]), f64(ns["test_x"]), f64(ns["test_y"])


def main():
    assert ref_shim.available(), "reference not mounted at " + ref_shim.REF_ROOT
    # ---- C1: SIMPLE-DATA FULL, N=120, T=300, D=1 (SF:161-163) -------------------------
    X, y, Xs, ys = simple_data("SF")
    run_case("c1_simple_full_init", "SF", X, y, Xs, ys, np.array([1.0, 1.0, 1.0]), 1)
    run_case("c1_simple_full_late", "SF", X, y, Xs, ys, np.array([0.2, 0.1, np.log(0.09)]), 1)
    # ---- C2: SIMPLE-FITC, M=5 (SC:187, SC:200) ------------------------------------------
    X, y, Xs, ys = simple_data("SC")
    torch.manual_seed(0)
    U = torch.randint(-3, 3, (5, 1)).double().numpy()
    U = U + 0.25 * np.arange(5).reshape(5, 1)  # SC:200 can draw duplicates; keep inducing inputs distinct
    run_case("c2_simple_fitc_init", "SC", X, y, Xs, ys, np.array([1.0, 1.0, 1.0]), 1, U=U)
    run_case("c2_simple_fitc_late", "SC", X, y, Xs, ys, np.array([0.2, 0.1, np.log(0.09)]), 1, U=U)
    # ---- C3: kin40k-FULL as the script runs it, N=T=500, D=8 (KF:196-213) ----------------
    X, y, Xs, ys = synth.kin40k_like(500, 500)
    run_case("c3_kin_full_P1", "KF", X, y, Xs, ys, synth.hyper_point("P1"), 8)
    run_case("c3_kin_full_P2", "KF", X, y, Xs, ys, synth.hyper_point("P2"), 8)
    # ragged: N not a multiple of any tile, T != N
    X, y, Xs, ys = synth.kin40k_like(333, 77, seed=10)
    run_case("c3_kin_full_ragged", "KF", X, y, Xs, ys, synth.hyper_point("P1", seed=11), 8)
    # scaled: N=1500
    X, y, Xs, ys = synth.kin40k_like(1500, 600, seed=20)
    run_case("c3_kin_full_1500_P2", "KF", X, y, Xs, ys, synth.hyper_point("P2"), 8)
    # ---- C4: FITC-20, N=T=500, D=8, M=20 (K20:185-215) ----------------------------------
    X, y, Xs, ys = synth.kin40k_like(500, 500)
    U = synth.inducing_init(20)
    run_case("c4_kin_fitc_P1", "K20", X, y, Xs, ys, synth.hyper_point("P1"), 8, U=U)
    run_case("c4_kin_fitc_P2", "K20", X, y, Xs, ys, synth.hyper_point("P2"), 8, U=U)
    # isotropic 1-element para_l as in the K20 log-score run (K20:422)
    run_case("c4_kin_fitc_iso", "K20", X, y, Xs, ys, np.array([1.0, 1.0, 1.0]), 1, U=U)
    X, y, Xs, ys = synth.kin40k_like(333, 77, seed=10)
    run_case("c4_kin_fitc_ragged", "K20", X, y, Xs, ys, synth.hyper_point("P1", seed=11), 8,
             U=synth.inducing_init(7, seed=12))
    X, y, Xs, ys = synth.kin40k_like(1500, 600, seed=20)
    run_case("c4_kin_fitc_1500_P2", "K20", X, y, Xs, ys, synth.hyper_point("P2"), 8,
             U=synth.inducing_init(20, seed=21))


if __name__ == "__main__":
    main()
