"""Generate tests/golden/reference_goldens.json by running the REAL reference.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_goldens.py``.

The reference imports ``streamlit`` from ``src/utils/decorators/caching.py:3``
(reached through ``src/__init__.py:24-31``); a pass-through stub module is put
on ``sys.modules`` so nothing under ``/root/reference`` is modified.  Normal
draws depend on the NumPy build, so the NumPy version is recorded and the
tests compare bit-for-bit only when it matches.
"""

from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import scipy

REFERENCE_ROOT = os.environ.get("OPTIONSLAB_REFERENCE", "/root/reference")


def import_reference():
    st = types.ModuleType("streamlit")

    def _passthrough(*dargs, **dkwargs):
        if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
            return dargs[0]
        return lambda fn: fn

    st.cache_data = _passthrough
    st.cache_resource = _passthrough
    sys.modules.setdefault("streamlit", st)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from src.greeks.unified_greeks import ExoticAdapter, compute_greeks_unified
    from src.pricing_models.black_scholes import black_scholes
    from src.pricing_models.exotic_options import AsianOption, AutocallableOption, BarrierOption, CliquetOption, LookbackOption
    from src.pricing_models.heston import HestonPricer
    from src.pricing_models.jump_diffusion import KouJumpDiffusion, MertonJumpDiffusion
    from src.pricing_models.monte_carlo import MCMethod, MonteCarloPricer
    from src.pricing_models.monte_carlo_unified import MonteCarloPricerUni

    return dict(HestonPricer=HestonPricer, MertonJumpDiffusion=MertonJumpDiffusion, KouJumpDiffusion=KouJumpDiffusion, MCMethod=MCMethod, MonteCarloPricer=MonteCarloPricer, MonteCarloPricerUni=MonteCarloPricerUni,
                AsianOption=AsianOption, BarrierOption=BarrierOption, LookbackOption=LookbackOption,
                AutocallableOption=AutocallableOption, CliquetOption=CliquetOption,
                compute_greeks_unified=compute_greeks_unified, ExoticAdapter=ExoticAdapter,
                black_scholes=black_scholes)


def main():
    ref = import_reference()
    MCP, Uni = ref["MonteCarloPricer"], ref["MonteCarloPricerUni"]
    P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)
    g = {"numpy": np.__version__, "scipy": scipy.__version__, "point": P, "seed": 42}

    g["black_scholes"] = {"call": float(ref["black_scholes"](**P, option_type="call")),
                          "put": float(ref["black_scholes"](**P, option_type="put")),
                          "call_q2": float(ref["black_scholes"](**P, option_type="call", q=0.02))}

    # --- MonteCarloPricer (monte_carlo.py:108-152) ---------------------------------
    eu = {}
    for n_sims, n_steps in [(10000, 50), (100000, 252), (100000, 1), (4096, 7)]:
        for ot in ("call", "put"):
            pr = MCP(n_sims, n_steps, seed=42)
            res = pr.price(**P, option_type=ot, return_error=True)
            term = pr._simulate(P["S"], P["T"], P["r"], P["sigma"], 0.0)
            pay = np.maximum(term - P["K"], 0.0) if ot == "call" else np.maximum(P["K"] - term, 0.0)
            eu[f"{n_sims}x{n_steps}_{ot}"] = {
                "price": res.price, "std_error": res.std_error, "n_paths": res.n_paths,
                "payoff_head": pay[:8].tolist(), "payoff_mirror_head": pay[n_sims:n_sims + 8].tolist(),
                "payoff_sum": float(np.sum(pay)),
            }
    pr = MCP(20000, 64, seed=7)
    eu["20000x64_call_q"] = {"price": pr.price(105.0, 95.0, 0.75, 0.03, 0.35, "call", q=0.02, return_error=True).price}
    g["european"] = eu
    g["control_variate"] = {
        "10000x50_call": MCP(10000, 50, seed=42).price_with_control_variate(**P, option_type="call"),
        "10000x50_put": MCP(10000, 50, seed=42).price_with_control_variate(**P, option_type="put"),
        "100000x1_call": MCP(100000, 1, seed=42).price_with_control_variate(**P, option_type="call"),
        "20000x64_call_q": MCP(20000, 64, seed=7).price_with_control_variate(105.0, 95.0, 0.75, 0.03, 0.35, "call", q=0.02),
    }

    # --- MCMethod.QMC backend (monte_carlo.py:94-97 -> gbm_qmc.py:14-47) -----------------
    qmc = {}
    import warnings
    for n_sims, n_steps in [(4096, 7), (16384, 64), (65536, 252), (10000, 50)]:
        for ot in ("call", "put"):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")  # scipy warns when N is not a power of two; the reference lets it through
                pr = MCP(n_sims, n_steps, seed=42, method=ref["MCMethod"].QMC)
                res = pr.price(**P, option_type=ot, return_error=True)
                term = pr._simulate(P["S"], P["T"], P["r"], P["sigma"], 0.0)
            pay = np.maximum(term - P["K"], 0.0) if ot == "call" else np.maximum(P["K"] - term, 0.0)
            qmc[f"{n_sims}x{n_steps}_{ot}"] = {"price": res.price, "std_error": res.std_error, "n_paths": res.n_paths,
                                               "payoff_head": pay[:8].tolist(), "payoff_sum": float(np.sum(pay))}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        qmc["16384x32_call_q"] = {"price": MCP(16384, 32, seed=7, method=ref["MCMethod"].QMC).price(105.0, 95.0, 0.75, 0.03, 0.35, "call", q=0.02)}
    g["qmc"] = qmc

    # --- MonteCarloPricerUni NumPy backend (monte_carlo_unified.py:451-689) ---------
    uni = Uni(10000, 50, seed=42, use_numba=False, use_gpu=False)
    g["uni"] = {"price_call": uni.price(**P, option_type="call"), "price_put": uni.price(**P, option_type="put")}
    S_v = np.array([100.0, 110.0, 90.0, 100.0, 100.0])
    K_v = np.array([100.0, 100.0, 100.0, 95.0, 105.0])
    T_v = np.array([1.0, 1.0, 1.0, 0.5, 0.5])
    r_v = np.full(5, 0.05)
    s_v = np.array([0.2, 0.2, 0.2, 0.3, 0.15])
    q_v = np.array([0.0, 0.0, 0.0, 0.02, 0.01])
    uni_b = Uni(2000, 20, seed=42, use_numba=False, use_gpu=False)
    g["uni"]["batch_inputs"] = dict(S=S_v.tolist(), K=K_v.tolist(), T=T_v.tolist(), r=r_v.tolist(),
                                    sigma=s_v.tolist(), q=q_v.tolist(), num_simulations=2000, num_steps=20)
    g["uni"]["price_batch_call"] = uni_b.price_batch(S_v, K_v, T_v, r_v, s_v, "call", q_v).tolist()
    g["uni"]["price_batch_put"] = uni_b.price_batch(S_v, K_v, T_v, r_v, s_v, "put", q_v).tolist()
    d, gm = uni_b.delta_gamma_batch(S_v, K_v, T_v, r_v, s_v, "call", q_v, h=1.0)
    g["uni"]["delta_gamma_batch_h1"] = {"delta": d.tolist(), "gamma": gm.tolist()}
    d1, g1 = Uni(10000, 50, seed=42, use_numba=False).delta_gamma(**P, option_type="call", h=1.0, seed=42)
    g["uni"]["delta_gamma_h1_seed42"] = [d1, g1]

    # --- compute_greeks_unified (unified_greeks.py:235-367) -------------------------
    gk = {}
    for ot in ("call", "put"):
        out = ref["compute_greeks_unified"](MCP(100000, 252, seed=42), **P, option_type=ot)
        gk[f"100000x252_{ot}"] = {k: float(v) for k, v in out.items()}
    out = ref["compute_greeks_unified"](MCP(20000, 16, seed=3), 105.0, 95.0, 0.75, 0.03, 0.35, "put", q=0.02)
    gk["20000x16_put_q"] = {k: float(v) for k, v in out.items()}
    out = ref["compute_greeks_unified"](MCP(20000, 16, seed=3), 100.0, 100.0, 0.002, 0.05, 0.2, "call")
    gk["20000x16_call_shortT"] = {k: float(v) for k, v in out.items()}
    g["greeks"] = gk

    # --- Exotics (exotic_options.py:97-131,174-224,368-401) -------------------------
    ex = {}
    for n_paths, n_steps in [(100000, 252), (5000, 12)]:
        a = ref["AsianOption"](**P, seed=42)
        tag = f"{n_paths}x{n_steps}"
        ex[f"asian_arith_call_{tag}"] = float(a.price(n_paths, n_steps, "arithmetic", "call"))
        ex[f"asian_arith_put_{tag}"] = float(a.price(n_paths, n_steps, "arithmetic", "put"))
        ex[f"asian_geom_call_{tag}"] = float(a.price(n_paths, n_steps, "geometric", "call"))
        lb = ref["LookbackOption"](**P, seed=42)
        for lt in ("floating", "fixed"):
            for ot in ("call", "put"):
                ex[f"lookback_{lt}_{ot}_{tag}"] = float(lb.price(n_paths, n_steps, lt, ot))
    ex["asian_geom_closed_form_call"] = float(ref["AsianOption"](**P).price_geometric_closed_form("call"))
    for n_paths, n_steps in [(100000, 365), (5000, 12)]:
        tag = f"{n_paths}x{n_steps}"
        for B, kinds in [(120.0, ("up-and-out", "up-and-in")), (85.0, ("down-and-out", "down-and-in"))]:
            b = ref["BarrierOption"](**P, seed=42, barrier=B)
            for bt in kinds:
                for ot in ("call", "put"):
                    ex[f"barrier_{bt}_{ot}_B{int(B)}_{tag}"] = float(b.price(n_paths, n_steps, bt, ot))
    ad = ref["ExoticAdapter"](ref["AsianOption"](**P, seed=42), n_paths=20000, n_steps=32, avg_type="arithmetic")
    out = ref["compute_greeks_unified"](ad, **P, option_type="call")
    ex["asian_adapter_greeks_20000x32"] = {k: float(v) for k, v in out.items()}
    g["exotics"] = ex

    # --- structured products of the same file (exotic_options.py:404-552): autocallable and cliquet ----------------
    sp = {}
    for n_paths, n_steps, freq, nper in [(100000, 252, 21, 12), (5000, 12, 3, 4), (20000, 100, 7, 9), (4097, 37, 37, 37)]:
        tag = f"{n_paths}x{n_steps}"
        sp[f"autocallable_default_{tag}_f{freq}"] = float(ref["AutocallableOption"](**P, seed=42).price(n_paths, n_steps, freq))
        sp[f"autocallable_tight_{tag}_f{freq}"] = float(ref["AutocallableOption"](
            **P, seed=42, autocall_barrier=1.05, coupon_barrier=0.9, coupon_rate=0.08, ki_barrier=0.75).price(n_paths, n_steps, freq))
        sp[f"cliquet_default_{tag}_p{nper}"] = float(ref["CliquetOption"](**P, seed=42).price(n_paths, n_steps, nper))
        sp[f"cliquet_wide_{tag}_p{nper}"] = float(ref["CliquetOption"](
            **P, seed=42, local_cap=0.08, local_floor=-0.03, global_cap=0.5, global_floor=-0.1).price(n_paths, n_steps, nper))
    sp["autocallable_q_sigma_20000x64_f8"] = float(ref["AutocallableOption"](105.0, 100.0, 1.5, 0.03, 0.35, 0.02, seed=7).price(20000, 64, 8))
    sp["cliquet_q_sigma_20000x64_p8"] = float(ref["CliquetOption"](105.0, 100.0, 1.5, 0.03, 0.35, 0.02, seed=7).price(20000, 64, 8))
    ad = ref["ExoticAdapter"](ref["CliquetOption"](**P, seed=42), n_paths=20000, n_steps=36)
    sp["cliquet_adapter_greeks_20000x36"] = {k: float(v) for k, v in ref["compute_greeks_unified"](ad, **P, option_type="call", n_periods=6).items()}
    g["structured"] = sp

    # --- Heston / Merton / Kou Monte Carlo (heston.py:184-255, jump_diffusion.py:160-225, :325-377) ------------
    import warnings
    md = {}
    hp = dict(kappa=2.0, theta=0.04, sigma_v=0.3, rho=-0.7, v0=0.04)
    hes = ref["HestonPricer"](**hp)
    md["heston_params"] = hp
    for n_paths, n_steps in [(20000, 50), (4097, 7), (100000, 252)]:
        for ot in ("call", "put"):
            md[f"heston_{ot}_{n_paths}x{n_steps}"] = float(hes.price_monte_carlo(100.0, 100.0, 1.0, 0.05, 0.01, ot, n_paths, n_steps, seed=42))
    md["heston_analytic_call"] = float(hes.price_european(100.0, 100.0, 1.0, 0.05, 0.01, "call"))
    md["heston_analytic_put"] = float(hes.price_european(100.0, 100.0, 1.0, 0.05, 0.01, "put"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hes2 = ref["HestonPricer"](kappa=1.0, theta=0.09, sigma_v=0.8, rho=-0.3, v0=0.02)  # Feller violated: truncation active
    md["heston_feller_violated_call_20000x50"] = float(hes2.price_monte_carlo(100.0, 110.0, 0.5, 0.03, 0.0, "call", 20000, 50, seed=7))
    mp = dict(lambda_j=1.0, mu_j=-0.1, sigma_j=0.15)
    mer = ref["MertonJumpDiffusion"](**mp)
    md["merton_params"] = mp
    for n_paths, n_steps in [(5000, 20), (20000, 50)]:
        for ot in ("call", "put"):
            md[f"merton_{ot}_{n_paths}x{n_steps}"] = float(mer.price_monte_carlo(100.0, 100.0, 1.0, 0.05, 0.2, ot, 0.01, n_paths, n_steps, seed=42))
    md["merton_analytic_call"] = float(mer.price(100.0, 100.0, 1.0, 0.05, 0.2, "call", 0.01))
    md["merton_analytic_put"] = float(mer.price(100.0, 100.0, 1.0, 0.05, 0.2, "put", 0.01))
    kp = dict(lambda_j=2.0, p=0.4, eta1=10.0, eta2=5.0)
    kou = ref["KouJumpDiffusion"](**kp)
    md["kou_params"] = kp
    for n_paths, n_steps in [(5000, 20), (20000, 50)]:
        for ot in ("call", "put"):
            md[f"kou_{ot}_{n_paths}x{n_steps}"] = float(kou.price_monte_carlo(100.0, 100.0, 1.0, 0.05, 0.2, ot, 0.01, n_paths, n_steps, seed=42))
    g["models"] = md

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")
    with open(path, "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
