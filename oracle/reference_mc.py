"""TEST INFRASTRUCTURE ONLY — NumPy restatement of OptionsLab's Monte Carlo hot path.

This file is the parity oracle for the CUDA engine.  It is never imported by
``optionslab_b200``; see ``oracle/__init__.py`` for who may use it.

Every function restates one piece of the reference's arithmetic in FP64 with
the SAME floating-point expression shape (so that on the same NumPy build the
results are bit-identical to the reference; ``tests/test_oracle_golden.py``
checks that against ``tests/golden/reference_goldens.json``).  Citations are
relative to the reference repository root.
"""

from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

# --------------------------------------------------------------------------
# Normal draws, exactly as the reference obtains them
# --------------------------------------------------------------------------


def normals_generator(seed: int, shape) -> np.ndarray:
    """PCG64 + ziggurat draws used by the European simulators.

    src/simulation/gbm_numpy.py:32,43 (``default_rng(seed).standard_normal((N, n))``),
    :77 (shape ``(N,)`` for the single-step variant) and
    src/pricing_models/monte_carlo_unified.py:321,329 (shape ``(n_opt, N, n)``).
    """
    return np.random.default_rng(seed).standard_normal(shape)


def normals_legacy(seed: Optional[int], shape) -> np.ndarray:
    """Legacy MT19937 draws used by the exotic path generator.

    src/pricing_models/exotic_options.py:51-52,59 seeds the *global* legacy
    state; ``RandomState(seed)`` yields the identical stream without touching
    global state.
    """
    return np.random.RandomState(seed).standard_normal(shape)


# --------------------------------------------------------------------------
# European: terminal prices, payoff, discounted mean, standard error
# --------------------------------------------------------------------------


def gbm_terminal_from_normals(S, T, r, sigma, q, Z, antithetic: bool = True) -> np.ndarray:
    """Terminal prices from a given ``Z`` of shape ``(N, n_steps)``.

    src/simulation/gbm_numpy.py:35-53: per-step drift and vol, the log-price is
    ``ln S + drift*n + vol*sum_j Z_ij`` (NumPy pairwise row sum); the mirrored
    paths use ``- vol*sum``; output is ``[+Z paths..., -Z paths...]``.
    """
    n_steps = Z.shape[1]
    dt = T / n_steps
    drift = (r - q - 0.5 * sigma * sigma) * dt
    vol = sigma * np.sqrt(dt)
    total_drift = drift * n_steps
    log_S0 = np.log(S)
    W = np.sum(Z, axis=1)
    up = log_S0 + total_drift + vol * W
    if not antithetic:
        return np.exp(up)
    down = log_S0 + total_drift - vol * W
    return np.concatenate([np.exp(up), np.exp(down)])


def gbm_terminal_single_step(S, T, r, sigma, q, Z) -> np.ndarray:
    """Single-step closed form, ``Z`` of shape ``(N,)``.

    src/simulation/gbm_numpy.py:71-83.
    """
    drift = (r - q - 0.5 * sigma * sigma) * T
    vol = sigma * np.sqrt(T)
    log_S0 = np.log(S)
    return np.concatenate([np.exp(log_S0 + drift + vol * Z), np.exp(log_S0 + drift - vol * Z)])


def vanilla_payoffs(terminal: np.ndarray, K, option_type: str) -> np.ndarray:
    """src/pricing_models/monte_carlo.py:140-143 (also monte_carlo_unified.py:503-506)."""
    if option_type == "call":
        return np.maximum(terminal - K, 0.0)
    return np.maximum(K - terminal, 0.0)


def discounted_mean(payoffs: np.ndarray, r, T) -> float:
    """src/pricing_models/monte_carlo.py:145-146."""
    return float(np.exp(-r * T) * np.mean(payoffs))


def discounted_std_error(payoffs: np.ndarray, r, T) -> float:
    """Population std (ddof=0) over all 2N antithetic samples / sqrt(2N).

    src/pricing_models/monte_carlo.py:148-150.
    """
    return float(np.exp(-r * T) * np.std(payoffs) / np.sqrt(len(payoffs)))


@dataclass
class OracleResult:
    price: float
    std_error: float
    n_paths: int
    payoffs: np.ndarray


def european_price(S, K, T, r, sigma, option_type, q=0.0, *, num_simulations, num_steps, seed) -> OracleResult:
    """``MonteCarloPricer(num_simulations, num_steps, seed).price(..., return_error=True)``.

    src/pricing_models/monte_carlo.py:74-106 (dispatch: ``num_steps == 1`` takes
    the single-step simulator), :133-152.
    """
    if T <= 0:
        intrinsic = max(S - K, 0) if option_type == "call" else max(K - S, 0)
        return OracleResult(intrinsic, 0.0, 0, np.empty(0))
    if num_steps == 1:
        Z = normals_generator(seed, num_simulations)
        terminal = gbm_terminal_single_step(S, T, r, sigma, q, Z)
    else:
        Z = normals_generator(seed, (num_simulations, num_steps))
        terminal = gbm_terminal_from_normals(S, T, r, sigma, q, Z)
    pay = vanilla_payoffs(terminal, K, option_type)
    return OracleResult(discounted_mean(pay, r, T), discounted_std_error(pay, r, T), len(pay), pay)


def qmc_uniforms(seed, n_points: int, n_steps: int) -> np.ndarray:
    """The reference's point set: scrambled Sobol, d = min(n_steps, 21201) (src/simulation/gbm_qmc.py:30-33)."""
    from scipy.stats.qmc import Sobol

    return Sobol(d=min(n_steps, 21201), scramble=True, seed=seed).random(n_points)


def qmc_normals_from_uniforms(uniforms: np.ndarray) -> np.ndarray:
    """src/simulation/gbm_qmc.py:36: ``norm.ppf(np.clip(u, 1e-10, 1 - 1e-10))``."""
    from scipy.stats import norm

    return norm.ppf(np.clip(uniforms, 1e-10, 1 - 1e-10))


def qmc_terminal_from_normals(S, T, r, sigma, q, normals: np.ndarray) -> np.ndarray:
    """src/simulation/gbm_qmc.py:38-47: ``ln S + drift*n + vol*sum(normals, axis=1)`` -> exp (N prices, no mirror)."""
    effective_steps = normals.shape[1]
    dt = T / effective_steps
    drift = (r - q - 0.5 * sigma * sigma) * dt
    vol = sigma * np.sqrt(dt)
    log_S0 = np.log(S)
    log_S_T = log_S0 + drift * effective_steps + vol * np.sum(normals, axis=1)
    return np.exp(log_S_T)


def european_price_qmc(S, K, T, r, sigma, option_type, q=0.0, *, num_simulations, num_steps, seed) -> OracleResult:
    """``MonteCarloPricer(N, n, seed, method=MCMethod.QMC).price(..., return_error=True)``:
    src/pricing_models/monte_carlo.py:94-97 -> gbm_qmc.py:14-47, then :140-150 (std error over the N samples)."""
    if T <= 0:
        intrinsic = max(S - K, 0) if option_type == "call" else max(K - S, 0)
        return OracleResult(intrinsic, 0.0, 0, np.empty(0))
    normals = qmc_normals_from_uniforms(qmc_uniforms(seed, num_simulations, num_steps))
    pay = vanilla_payoffs(qmc_terminal_from_normals(S, T, r, sigma, q, normals), K, option_type)
    return OracleResult(discounted_mean(pay, r, T), discounted_std_error(pay, r, T), len(pay), pay)


def control_variate_from_terminal(terminal: np.ndarray, S, K, T, r, q, option_type: str) -> float:
    """``price_with_control_variate`` given the simulated terminals.

    src/pricing_models/monte_carlo.py:168-186: discounted payoffs, ``np.cov`` (ddof=1) against the
    terminal spot, beta zeroed when var(S_T) <= 1e-10, adjustment by mean(S_T) - S*exp((r-q)T).
    """
    pay = vanilla_payoffs(terminal, K, option_type)
    discounted = np.exp(-r * T) * pay
    control_mean = np.mean(terminal)
    forward = S * np.exp((r - q) * T)
    cov = np.cov(discounted, terminal)
    beta = cov[0, 1] / cov[1, 1] if cov[1, 1] > 1e-10 else 0.0
    return float(np.mean(discounted) - beta * (control_mean - forward))


def european_price_control_variate(S, K, T, r, sigma, option_type, q=0.0, *, num_simulations, num_steps, seed) -> float:
    """``MonteCarloPricer(...).price_with_control_variate`` (monte_carlo.py:154-186)."""
    if num_steps == 1:
        terminal = gbm_terminal_single_step(S, T, r, sigma, q, normals_generator(seed, num_simulations))
    else:
        terminal = gbm_terminal_from_normals(S, T, r, sigma, q, normals_generator(seed, (num_simulations, num_steps)))
    return control_variate_from_terminal(terminal, S, K, T, r, q, option_type)


# --------------------------------------------------------------------------
# MonteCarloPricerUni: batched terminals (NumPy backend)
# --------------------------------------------------------------------------


def uni_terminal_from_normals(S_arr, T_arr, r_arr, sigma_arr, q_arr, Z) -> np.ndarray:
    """``Z`` of shape ``(n_opt, N, n_steps)`` -> terminals ``(n_opt, 2N)``.

    src/pricing_models/monte_carlo_unified.py:324-343: per-option drift/vol,
    sequential ``cumsum`` of the increments along the step axis (once for +Z,
    once for -Z), last column, ``exp``.
    """
    n_steps = Z.shape[2]
    dt = T_arr / n_steps
    drift = (r_arr - q_arr - 0.5 * sigma_arr**2)[:, None] * dt[:, None]
    vol = sigma_arr[:, None] * np.sqrt(dt[:, None])
    log_S = np.log(S_arr)[:, None, None]
    pos = log_S + np.cumsum(drift[:, None, :] + vol[:, None, :] * Z, axis=2)
    neg = log_S + np.cumsum(drift[:, None, :] - vol[:, None, :] * Z, axis=2)
    return np.concatenate([np.exp(pos[:, :, -1]), np.exp(neg[:, :, -1])], axis=1)


def uni_price_batch(S_vals, K_vals, T_vals, r_vals, sigma_vals, option_type, q_vals=0.0, *,
                    num_simulations, num_steps, seed) -> np.ndarray:
    """``MonteCarloPricerUni(..., use_numba=False).price_batch``.

    src/pricing_models/monte_carlo_unified.py:601-631.
    """
    S_vals = np.asarray(S_vals, dtype=np.float64)
    K_vals = np.asarray(K_vals, dtype=np.float64)
    T_vals = np.asarray(T_vals, dtype=np.float64)
    r_vals = np.asarray(r_vals, dtype=np.float64)
    sigma_vals = np.asarray(sigma_vals, dtype=np.float64)
    if isinstance(q_vals, (int, float)):
        q_vals = np.full_like(S_vals, q_vals)
    else:
        q_vals = np.asarray(q_vals, dtype=np.float64)
    Z = normals_generator(seed, (len(S_vals), num_simulations, num_steps))
    terminal = uni_terminal_from_normals(S_vals, T_vals, r_vals, sigma_vals, q_vals, Z)
    if option_type == "call":
        pay = np.maximum(terminal - K_vals[:, None], 0.0)
    else:
        pay = np.maximum(K_vals[:, None] - terminal, 0.0)
    return np.exp(-r_vals * T_vals) * np.mean(pay, axis=1)


def uni_price(S, K, T, r, sigma, option_type, q=0.0, *, num_simulations, num_steps, seed) -> float:
    """``MonteCarloPricerUni.price`` (src/pricing_models/monte_carlo_unified.py:493-508)."""
    Z = normals_generator(seed, (1, num_simulations, num_steps))
    terminal = uni_terminal_from_normals(np.array([S]), np.array([T]), np.array([r]),
                                         np.array([sigma]), np.array([q]), Z)[0]
    pay = vanilla_payoffs(terminal, K, option_type)
    return float(np.exp(-r * T) * np.mean(pay))


def central_delta_gamma(price_up, price_mid, price_down, h):
    """src/pricing_models/monte_carlo_unified.py:557-558 / :684-685."""
    return (price_up - price_down) / (2 * h), (price_up - 2 * price_mid + price_down) / (h**2)


# --------------------------------------------------------------------------
# Exotics: full paths, Asian / barrier / lookback payoffs
# --------------------------------------------------------------------------


def exotic_paths_from_normals(S, T, r, sigma, q, Z) -> np.ndarray:
    """Full path array ``(N, n_steps+1)`` with column 0 = S.

    src/pricing_models/exotic_options.py:54-67: sequential cumsum of
    ``drift + diffusion*Z`` added to ``ln S``, then ``exp`` of everything.
    """
    n_paths, n_steps = Z.shape
    dt = T / n_steps
    drift = (r - q - 0.5 * sigma**2) * dt
    diffusion = sigma * np.sqrt(dt)
    log_S = np.zeros((n_paths, n_steps + 1))
    log_S[:, 0] = np.log(S)
    log_S[:, 1:] = np.log(S) + np.cumsum(drift + diffusion * Z, axis=1)
    return np.exp(log_S)


def asian_payoffs(paths, K, avg_type="arithmetic", option_type="call") -> np.ndarray:
    """src/pricing_models/exotic_options.py:118-128 (average excludes column 0)."""
    if avg_type == "arithmetic":
        avg = np.mean(paths[:, 1:], axis=1)
    else:
        avg = np.exp(np.mean(np.log(paths[:, 1:]), axis=1))
    if option_type == "call":
        return np.maximum(avg - K, 0)
    return np.maximum(K - avg, 0)


def barrier_payoffs(paths, K, barrier, barrier_type="up-and-out", option_type="call") -> np.ndarray:
    """src/pricing_models/exotic_options.py:198-222 (monitoring includes column 0, >= / <=)."""
    if barrier <= 0:
        raise ValueError("Barrier must be positive")
    if barrier_type.startswith("up"):
        crossed = np.any(paths >= barrier, axis=1)
    else:
        crossed = np.any(paths <= barrier, axis=1)
    active = ~crossed if barrier_type.endswith("out") else crossed
    S_T = paths[:, -1]
    pay = np.maximum(S_T - K, 0) if option_type == "call" else np.maximum(K - S_T, 0)
    return pay * active


def lookback_payoffs(paths, K, lookback_type="floating", option_type="call") -> np.ndarray:
    """src/pricing_models/exotic_options.py:380-399 (extrema include column 0)."""
    S_T = paths[:, -1]
    S_max = np.max(paths, axis=1)
    S_min = np.min(paths, axis=1)
    if lookback_type == "floating":
        return S_T - S_min if option_type == "call" else S_max - S_T
    if option_type == "call":
        return np.maximum(S_max - K, 0)
    return np.maximum(K - S_min, 0)


def autocallable_payoffs(paths, S, T, r, autocall_barrier=1.0, coupon_barrier=0.8, coupon_rate=0.10, ki_barrier=0.6,
                         observation_freq=21) -> np.ndarray:
    """ALREADY DISCOUNTED per-path payoffs (fraction of notional) of AutocallableOption.price,
    src/pricing_models/exotic_options.py:438-488: early redemption at the first observation with S_t/S >= autocall
    barrier pays (1 + coupon_rate * (i+1)/n_obs * T) * exp(-r t dt); otherwise 1 (+ coupon_rate*T above the coupon
    barrier) at maturity, replaced by S_T/S when the path minimum (t = 0 included) touched the knock-in barrier and
    S_T < S; discounted with exp(-rT).  The price is the plain mean of these."""
    n_paths, n_cols = paths.shape
    n_steps = n_cols - 1
    dt = T / n_steps
    obs_times = list(range(observation_freq, n_steps + 1, observation_freq))
    n_obs = len(obs_times)
    payoffs = np.zeros(n_paths)
    redeemed = np.zeros(n_paths, dtype=bool)
    knocked_in = np.min(paths / S, axis=1) <= ki_barrier
    for i, t in enumerate(obs_times):
        active = ~redeemed
        S_rel = paths[active, t] / S
        idx = np.where(active)[0][S_rel >= autocall_barrier]
        coupon = coupon_rate * ((i + 1) / n_obs) * T
        payoffs[idx] = (1 + coupon) * np.exp(-r * t * dt)
        redeemed[idx] = True
    still = ~redeemed
    S_rel_final = paths[still, -1] / S
    final = np.ones(int(np.sum(still)))
    final[S_rel_final >= coupon_barrier] += coupon_rate * T
    loss = knocked_in[still] & (S_rel_final < 1.0)
    final[loss] = S_rel_final[loss]
    payoffs[still] = final * np.exp(-r * T)
    return payoffs


def cliquet_payoffs(paths, S, local_cap=0.05, local_floor=-0.05, global_cap=0.30, global_floor=0.0, n_periods=12) -> np.ndarray:
    """UNDISCOUNTED per-path payoffs of CliquetOption.price, src/pricing_models/exotic_options.py:525-552: sum over
    n_periods reset periods of n_steps // n_periods steps of the locally clipped simple returns, clipped globally,
    floored at 0, times S.  (Steps beyond n_periods * (n_steps // n_periods) are simulated and ignored.)"""
    n_paths, n_cols = paths.shape
    n_steps = n_cols - 1
    spp = n_steps // n_periods
    total = np.zeros(n_paths)
    for p in range(n_periods):
        S_start, S_end = paths[:, p * spp], paths[:, (p + 1) * spp]
        total += np.clip((S_end - S_start) / S_start, local_floor, local_cap)
    total = np.clip(total, global_floor, global_cap)
    return np.maximum(total, 0) * S


def exotic_price(kind, S, K, T, r, sigma, q=0.0, *, seed, n_paths, n_steps, option_type="call",
                 avg_type="arithmetic", barrier=0.0, barrier_type="up-and-out",
                 lookback_type="floating", return_payoffs=False, **product):
    """``AsianOption/BarrierOption/LookbackOption(...).price(n_paths, n_steps, ...)``.

    src/pricing_models/exotic_options.py:97-131, :174-224, :368-401.  Chunk-free:
    materialises the full path array like the reference does.
    """
    if kind == "barrier" and barrier <= 0:
        raise ValueError("Barrier must be positive")
    Z = normals_legacy(seed, (n_paths, n_steps))
    paths = exotic_paths_from_normals(S, T, r, sigma, q, Z)
    if kind == "asian":
        pay = asian_payoffs(paths, K, avg_type, option_type)
    elif kind == "barrier":
        pay = barrier_payoffs(paths, K, barrier, barrier_type, option_type)
    elif kind == "lookback":
        pay = lookback_payoffs(paths, K, lookback_type, option_type)
    elif kind == "cliquet":
        pay = cliquet_payoffs(paths, S, **product)
    elif kind == "autocallable":  # discounting happens per path, inside the payoffs (exotic_options.py:466,486-488)
        pay = autocallable_payoffs(paths, S, T, r, **product)
        price = np.mean(pay)
        return (price, pay) if return_payoffs else price
    else:
        raise ValueError(kind)
    price = np.exp(-r * T) * np.mean(pay)
    return (price, pay) if return_payoffs else price


def asian_geometric_closed_form(S, K, T, r, sigma, q=0.0, option_type="call") -> float:
    """Continuous-averaging closed form, src/pricing_models/exotic_options.py:143-160."""
    from scipy.stats import norm

    sigma_adj = sigma / np.sqrt(3)
    r_adj = 0.5 * (r - q - sigma**2 / 6)
    d1 = (np.log(S / K) + (r_adj + 0.5 * sigma_adj**2) * T) / (sigma_adj * np.sqrt(T))
    d2 = d1 - sigma_adj * np.sqrt(T)
    if option_type == "call":
        return S * np.exp((r_adj - r) * T) * norm.cdf(d1) - K * np.exp(-r * T) * norm.cdf(d2)
    return K * np.exp(-r * T) * norm.cdf(-d2) - S * np.exp((r_adj - r) * T) * norm.cdf(-d1)


# --------------------------------------------------------------------------
# Bump-and-revalue Greeks
# --------------------------------------------------------------------------

H_SIGMA = 0.01
H_R = 1e-4
H_T = 1 / 365.0


def greek_bumps(S):
    """src/greeks/unified_greeks.py:274-277."""
    return max(1e-4, 0.01 * S), max(1e-4, 0.01), 1e-4, 1 / 365.0


def greeks_bump_and_revalue(price_fn: Callable[..., float], S, K, T, r, sigma, q=0.0,
                            include_second_order: bool = True) -> "OrderedDict[str, float]":
    """Finite-difference Greeks over ``price_fn(S, K, T, r, sigma, q)``.

    src/greeks/unified_greeks.py:274-362: memoised scenario prices; delta/gamma
    central in S; vega central in sigma; theta one-sided over one day (n_steps
    kept, so dt shrinks); rho central in r; vanna 4-point cross; charm from the
    delta at T - 1/365; vomma second difference in sigma.
    """
    h_S, h_sigma, h_r, h_T = greek_bumps(S)
    cache = {}

    def P(S_=S, T_=T, r_=r, sigma_=sigma):
        key = (S_, K, T_, r_, sigma_, q)
        if key not in cache:
            cache[key] = price_fn(S_, K, T_, r_, sigma_, q)
        return cache[key]

    p_mid = P()
    p_S_up, p_S_down = P(S_=S + h_S), P(S_=S - h_S)
    delta = (p_S_up - p_S_down) / (2 * h_S)
    gamma = (p_S_up - 2 * p_mid + p_S_down) / (h_S**2)
    p_v_up, p_v_down = P(sigma_=sigma + h_sigma), P(sigma_=sigma - h_sigma)
    vega = (p_v_up - p_v_down) / (2 * h_sigma)
    if T > h_T:
        theta = (P(T_=T - h_T) - p_mid) / h_T
    else:
        theta = -p_mid / max(T, 1e-6)
    rho = (P(r_=r + h_r) - P(r_=r - h_r)) / (2 * h_r)
    out = OrderedDict(price=p_mid, delta=delta, gamma=gamma, vega=vega, theta=theta, rho=rho)
    if include_second_order:
        vanna = (P(S_=S + h_S, sigma_=sigma + h_sigma) - P(S_=S + h_S, sigma_=sigma - h_sigma)
                 - P(S_=S - h_S, sigma_=sigma + h_sigma) + P(S_=S - h_S, sigma_=sigma - h_sigma)) / (
            4 * h_S * h_sigma)
        if T > h_T:
            d_T = (P(S_=S + h_S, T_=T - h_T) - P(S_=S - h_S, T_=T - h_T)) / (2 * h_S)
            charm = (d_T - delta) / h_T
        else:
            charm = 0.0
        vomma = (p_v_up - 2 * p_mid + p_v_down) / (h_sigma**2)
        out["vanna"], out["charm"], out["vomma"] = vanna, charm, vomma
    return out


# --------------------------------------------------------------------------
# Black-Scholes closed form (analytic oracle for the 3-standard-error checks)
# --------------------------------------------------------------------------


def black_scholes(S, K, T, r, sigma, option_type="call", q=0.0) -> float:
    """src/pricing_models/black_scholes.py:33-52."""
    from scipy.stats import norm

    if S <= 0 or K <= 0 or T < 0 or sigma < 0:
        raise ValueError("Invalid input")
    if T == 0:
        return max(S - K, 0.0) if option_type == "call" else max(K - S, 0.0)
    d1 = (np.log(S / K) + (r - q + 0.5 * sigma**2) * T) / (sigma * np.sqrt(T))
    d2 = d1 - sigma * np.sqrt(T)
    if option_type == "call":
        return float(S * np.exp(-q * T) * norm.cdf(d1) - K * np.exp(-r * T) * norm.cdf(d2))
    return float(K * np.exp(-r * T) * norm.cdf(-d2) - S * np.exp(-q * T) * norm.cdf(-d1))


def black_scholes_greeks(S, K, T, r, sigma, option_type="call", q=0.0):
    """Analytic delta/gamma/vega (textbook BSM; used only as a statistical anchor)."""
    from scipy.stats import norm

    d1 = (np.log(S / K) + (r - q + 0.5 * sigma**2) * T) / (sigma * np.sqrt(T))
    pdf = norm.pdf(d1)
    delta = np.exp(-q * T) * (norm.cdf(d1) if option_type == "call" else norm.cdf(d1) - 1.0)
    gamma = np.exp(-q * T) * pdf / (S * sigma * np.sqrt(T))
    vega = S * np.exp(-q * T) * pdf * np.sqrt(T)
    return float(delta), float(gamma), float(vega)


# --------------------------------------------------------------------------
# Timed CPU baseline used by bench.py (bounded samples of the benchmark workload)
# --------------------------------------------------------------------------


def grid_workload(n_strikes=64, n_maturities=64):
    """The C5 grid of SURVEY.md §8(d): 64 strikes x 64 maturities, S=100, r=5%, sigma=20%."""
    K = np.linspace(60.0, 140.0, n_strikes)
    T = np.linspace(1.0 / 12.0, 2.0, n_maturities)
    KK, TT = np.meshgrid(K, T, indexing="ij")
    n = KK.size
    return dict(S=np.full(n, 100.0), K=KK.ravel().copy(), T=TT.ravel().copy(),
                r=np.full(n, 0.05), sigma=np.full(n, 0.2), q=np.zeros(n))


def cpu_grid_sample(option_indices, num_simulations, num_steps, seed, option_type="call"):
    """Price a slice of the C5 grid the way the reference would on CPU.

    One option at a time through the ``simulate_gbm_numpy`` arithmetic
    (gbm_numpy.py:32-53 + monte_carlo.py:140-146); the reference's own batch
    backend (monte_carlo_unified.py:329) would allocate ``n_opt*N*n`` doubles,
    which does not fit, so the slice is priced option by option with option
    ``i`` drawing from ``default_rng(seed + i)`` (the Numba backend's seeding,
    monte_carlo_unified.py:190).
    """
    g = grid_workload()
    out = np.empty(len(option_indices))
    for j, i in enumerate(option_indices):
        res = european_price(g["S"][i], g["K"][i], g["T"][i], g["r"][i], g["sigma"][i], option_type,
                             g["q"][i], num_simulations=num_simulations, num_steps=num_steps,
                             seed=seed + int(i))
        out[j] = res.price
    return out


# --------------------------------------------------------------------------
# Heston and jump-diffusion Monte Carlo
# --------------------------------------------------------------------------


_NUMBA_TERMINAL = None


def numba_backend_terminal(S, T, r, sigma, q, n_paths, n_steps, seed):
    """The reference's multi-core backend, MCMethod.NUMBA -> ``_gbm_terminal_parallel`` (src/simulation/gbm_numba.py:74-97),
    restated for TIMING on the host cores: a prange over paths, each path re-seeding Numba's per-thread generator with
    ``seed + i`` and walking its steps with both mirrored log-prices.  Returns 2 * n_paths terminal prices ([i] = +z path,
    [i + n_paths] = -z path).  Needs numba; raises ImportError otherwise."""
    global _NUMBA_TERMINAL
    if _NUMBA_TERMINAL is None:
        import numba

        @numba.njit(parallel=True, fastmath=True)
        def kernel(S, T, r, sigma, q, n_paths, n_steps, seed):
            dt = T / n_steps
            mu = (r - q - 0.5 * sigma * sigma) * dt
            vol = sigma * np.sqrt(dt)
            start = np.log(S)
            out = np.empty(2 * n_paths)
            for i in numba.prange(n_paths):
                np.random.seed(seed + i)
                up = start
                down = start
                for _ in range(n_steps):
                    z = np.random.randn()
                    up += mu + vol * z
                    down += mu - vol * z
                out[i] = np.exp(up)
                out[n_paths + i] = np.exp(down)
            return out

        _NUMBA_TERMINAL = kernel
    return _NUMBA_TERMINAL(float(S), float(T), float(r), float(sigma), float(q), int(n_paths), int(n_steps), int(seed))


def heston_draws(seed: Optional[int], n_paths: int, n_steps: int) -> np.ndarray:
    """The normals ``HestonPricer.price_monte_carlo`` consumes, in its order: per step ``standard_normal(n_paths)``
    for Z1 and again for the independent part of Z2 (src/pricing_models/heston.py:207-229, legacy global
    generator).  One C-ordered draw of shape (n_steps, 2, n_paths) is the same stream."""
    return np.random.RandomState(seed).standard_normal((n_steps, 2, n_paths))


def heston_payoffs_from_normals(S, K, T, r, q, kappa, theta, sigma_v, rho, v0, Z, option_type="call") -> np.ndarray:
    """src/pricing_models/heston.py:210-252 with the draws supplied (full-truncation Euler)."""
    n_steps, _, n_paths = Z.shape
    dt = T / n_steps
    sqrt_dt = np.sqrt(dt)
    log_S = np.full(n_paths, np.log(S))
    v = np.full(n_paths, v0)
    rho_sqrt = np.sqrt(1 - rho**2)
    for t in range(n_steps):
        Z1 = Z[t, 0]
        Z2 = rho * Z1 + rho_sqrt * Z[t, 1]
        v_pos = np.maximum(v, 0)
        sqrt_v = np.sqrt(v_pos)
        log_S += (r - q - 0.5 * v_pos) * dt + sqrt_v * sqrt_dt * Z1
        v += kappa * (theta - v_pos) * dt + sigma_v * sqrt_v * sqrt_dt * Z2
        v = np.maximum(v, 0)
    return vanilla_payoffs(np.exp(log_S), K, option_type)


def heston_price_mc(S, K, T, r, q=0.0, option_type="call", *, kappa, theta, sigma_v, rho, v0, n_paths, n_steps, seed) -> float:
    """``HestonPricer(kappa, theta, sigma_v, rho, v0).price_monte_carlo(...)`` (heston.py:184-255)."""
    pay = heston_payoffs_from_normals(S, K, T, r, q, kappa, theta, sigma_v, rho, v0, heston_draws(seed, n_paths, n_steps), option_type)
    return discounted_mean(pay, r, T)


def merton_kappa(mu_j, sigma_j) -> float:
    """src/pricing_models/jump_diffusion.py:65-67."""
    return float(np.exp(mu_j + 0.5 * sigma_j**2) - 1)


def kou_kappa(p, eta1, eta2) -> float:
    """src/pricing_models/jump_diffusion.py:302-308."""
    return p * eta1 / (eta1 - 1) + (1 - p) * eta2 / (eta2 + 1) - 1


def merton_draws(seed, lambda_j, mu_j, sigma_j, T, n_paths, n_steps):
    """Replay of the generator calls of ``MertonJumpDiffusion.price_monte_carlo`` (jump_diffusion.py:203-216):
    per step ``standard_normal(n_paths)``, ``poisson(lambda_j*dt, n_paths)``, then ``normal(mu_j, sigma_j, k)`` for each
    path with k > 0 jumps in path order (the legacy Gaussian keeps its cached second value across calls, so one
    draw of all the step's jump sizes is the same stream).  -> dW [n_steps, n_paths], J [n_steps, n_paths] with
    J[t, i] = ``np.sum(jump_sizes)`` of path i at step t (0 where there is no jump)."""
    rs = np.random.RandomState(seed)
    dt = T / n_steps
    dW = np.empty((n_steps, n_paths))
    J = np.zeros((n_steps, n_paths))
    for t in range(n_steps):
        dW[t] = rs.standard_normal(n_paths)
        n_jumps = rs.poisson(lambda_j * dt, n_paths)
        hit = np.flatnonzero(n_jumps)
        if hit.size:
            sizes = rs.normal(mu_j, sigma_j, int(n_jumps[hit].sum()))
            idx = 0
            for i in hit:
                J[t, i] = np.sum(sizes[idx:idx + n_jumps[i]])
                idx += n_jumps[i]
    return dW, J


def kou_draws(seed, lambda_j, p, eta1, eta2, T, n_paths, n_steps):
    """Replay of ``KouJumpDiffusion.price_monte_carlo`` (jump_diffusion.py:352-367) and ``simulate_jump`` (:310-323):
    per step ``standard_normal``, ``poisson``, and when any jump occurred ``uniform(0, 1, total)``, then
    ``exponential(1/eta1, #up)`` and ``exponential(1/eta2, #down)`` scattered by the up/down masks."""
    rs = np.random.RandomState(seed)
    dt = T / n_steps
    dW = np.empty((n_steps, n_paths))
    J = np.zeros((n_steps, n_paths))
    for t in range(n_steps):
        dW[t] = rs.standard_normal(n_paths)
        n_jumps = rs.poisson(lambda_j * dt, n_paths)
        total = int(np.sum(n_jumps))
        if total > 0:
            jumps = np.zeros(total)
            u = rs.uniform(0, 1, total)
            up = u < p
            jumps[up] = rs.exponential(1 / eta1, int(np.sum(up)))
            jumps[~up] = -rs.exponential(1 / eta2, int(np.sum(~up)))
            idx = 0
            for i in np.flatnonzero(n_jumps):
                J[t, i] = np.sum(jumps[idx:idx + n_jumps[i]])
                idx += n_jumps[i]
    return dW, J


def jump_payoffs_from_draws(S, K, T, r, sigma, q, lambda_kappa, dW, J, option_type="call") -> np.ndarray:
    """jump_diffusion.py:197-223 / :345-375 with the draws supplied: per step ``log_S += drift + vol*dW`` and then the
    step's jump sum on the paths it hit."""
    n_steps, n_paths = dW.shape
    dt = T / n_steps
    drift = (r - q - lambda_kappa - 0.5 * sigma**2) * dt
    vol = sigma * np.sqrt(dt)
    log_S = np.full(n_paths, np.log(S))
    for t in range(n_steps):
        log_S += drift + vol * dW[t]
        hit = J[t] != 0.0
        log_S[hit] += J[t][hit]
    return vanilla_payoffs(np.exp(log_S), K, option_type)


def merton_price_mc(S, K, T, r, sigma, option_type="call", q=0.0, *, lambda_j, mu_j, sigma_j, n_paths, n_steps, seed) -> float:
    dW, J = merton_draws(seed, lambda_j, mu_j, sigma_j, T, n_paths, n_steps)
    pay = jump_payoffs_from_draws(S, K, T, r, sigma, q, lambda_j * merton_kappa(mu_j, sigma_j), dW, J, option_type)
    return discounted_mean(pay, r, T)


def kou_price_mc(S, K, T, r, sigma, option_type="call", q=0.0, *, lambda_j, p, eta1, eta2, n_paths, n_steps, seed) -> float:
    dW, J = kou_draws(seed, lambda_j, p, eta1, eta2, T, n_paths, n_steps)
    pay = jump_payoffs_from_draws(S, K, T, r, sigma, q, lambda_j * kou_kappa(p, eta1, eta2), dW, J, option_type)
    return discounted_mean(pay, r, T)


def merton_series_price(S, K, T, r, sigma, option_type="call", q=0.0, *, lambda_j, mu_j, sigma_j, n_terms=50) -> float:
    """``MertonJumpDiffusion.price`` (jump_diffusion.py:69-132): Poisson-weighted Black-Scholes prices with
    lambda' = lambda(1+kappa), sigma_n^2 = sigma^2 + n sigma_j^2/T, r_n = r - lambda kappa + n ln(1+kappa)/T."""
    from math import factorial

    if T <= 0:
        return max(S - K, 0) if option_type == "call" else max(K - S, 0)
    kappa = merton_kappa(mu_j, sigma_j)
    lambda_prime = lambda_j * (1 + kappa)
    price = 0.0
    for n in range(n_terms):
        poisson_weight = np.exp(-lambda_prime * T) * (lambda_prime * T) ** n / factorial(n)
        sigma_n = np.sqrt(sigma**2 + n * sigma_j**2 / T)
        r_n = r - lambda_j * kappa + n * np.log(1 + kappa) / T
        price += poisson_weight * black_scholes(S, K, T, r_n, sigma_n, option_type, q)
        if poisson_weight < 1e-12:
            break
    return float(price)


def jump_terminal_exact_law(model, S, T, r, sigma, q, jump_params: dict, n_paths: int, seed: int) -> np.ndarray:
    """Independent statistical cross-check (NOT the reference's algorithm): S_T of a Merton / Kou jump diffusion
    sampled from its exact law in one step -- Gaussian diffusion over [0, T], N ~ Poisson(lambda T) jumps with the
    model's jump-size law, compensated drift.  Same distribution as the reference's step-wise scheme."""
    rng = np.random.default_rng(seed)
    lam = jump_params["lambda_j"]
    n_jumps = rng.poisson(lam * T, n_paths)
    total = int(n_jumps.sum())
    owner = np.repeat(np.arange(n_paths), n_jumps)
    if model == "merton":
        kappa = merton_kappa(jump_params["mu_j"], jump_params["sigma_j"])
        sizes = rng.normal(jump_params["mu_j"], jump_params["sigma_j"], total)
    else:
        kappa = kou_kappa(jump_params["p"], jump_params["eta1"], jump_params["eta2"])
        up = rng.uniform(0, 1, total) < jump_params["p"]
        sizes = np.where(up, rng.exponential(1 / jump_params["eta1"], total), -rng.exponential(1 / jump_params["eta2"], total))
    jump_sum = np.bincount(owner, weights=sizes, minlength=n_paths)
    log_S_T = np.log(S) + (r - q - lam * kappa - 0.5 * sigma**2) * T + sigma * np.sqrt(T) * rng.standard_normal(n_paths) + jump_sum
    return np.exp(log_S_T)
