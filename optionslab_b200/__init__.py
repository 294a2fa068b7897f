"""optionslab_b200 — B200-native Monte Carlo engine behind OptionsLab's pricer API.

Only the Monte Carlo hot path of Diegotistical/OptionsLab lives here: European /
Asian / barrier / lookback pricing and bump-and-revalue Greeks, executed by hand-written sm_100a
kernels in ``libb200mc.so`` (C ABI: ``include/b200mc.h``).  Importing this package does not touch
the GPU; the first pricing call creates the engine and raises ``AccelerationError`` if the
library or the device is missing — there is no CPU fallback.
"""

from .exceptions import AccelerationError, ConvergenceError, GreeksError, InputValidationError, MonteCarloError
from .exotic_options import (AsianOption, AutocallableOption, BarrierOption, CliquetOption, LookbackOption, price_asian,
                             price_barrier, price_lookback)
from .greeks import (ExerciseStyle, ExoticAdapter, HestonAdapter, JumpDiffusionAdapter, OptionType, PricerProtocol,
                     compute_greeks_unified, greeks_heston, greeks_jump_diffusion)
from .models import HestonPricer, KouJumpDiffusion, MertonJumpDiffusion
from .monte_carlo import MCMethod, MCResult, MonteCarloPricer
from .monte_carlo_unified import MonteCarloPricerUni
from . import simulation
from .validation import monte_carlo_convergence_test

__version__ = "0.1.0"
__all__ = [
    "MonteCarloPricer", "MCMethod", "MCResult", "MonteCarloPricerUni",
    "AsianOption", "BarrierOption", "LookbackOption", "AutocallableOption", "CliquetOption", "price_asian", "price_barrier",
    "price_lookback",
    "HestonPricer", "MertonJumpDiffusion", "KouJumpDiffusion",
    "monte_carlo_convergence_test", "simulation",
    "PricerProtocol", "ExoticAdapter", "HestonAdapter", "JumpDiffusionAdapter", "compute_greeks_unified", "greeks_heston",
    "greeks_jump_diffusion", "OptionType", "ExerciseStyle",
    "MonteCarloError", "InputValidationError", "ConvergenceError", "AccelerationError", "GreeksError",
]
