"""Exception classes of the Monte Carlo path — same names, hierarchy and message shapes as the
reference (src/exceptions/montecarlo_exceptions.py:24-131, src/exceptions/greek_exceptions.py:4-15)
so callers' ``except`` clauses keep working."""

__all__ = ["MonteCarloError", "InputValidationError", "ConvergenceError", "AccelerationError", "GreeksError"]


class MonteCarloError(Exception):
    """Base class of every Monte Carlo pricer error."""

    def __init__(self, message: str = "Monte Carlo computation error"):
        self.message = message
        super().__init__(self.message)


class InputValidationError(MonteCarloError):
    """Invalid pricer inputs (non-positive S/K/T, negative sigma, unknown option type, ...)."""

    def __init__(self, message: str = "Invalid input parameters"):
        super().__init__(f"Input validation failed: {message}")


class ConvergenceError(MonteCarloError):
    def __init__(self, message: str = "Simulation did not converge", iterations: int = 0):
        self.iterations = iterations
        super().__init__(f"{message} (after {iterations} iterations)")


class AccelerationError(MonteCarloError):
    """The CUDA engine failed or is unavailable.  There is no CPU fallback: this is raised loudly."""

    def __init__(self, message: str = "Hardware acceleration failed", backend: str = "unknown"):
        self.backend = backend
        super().__init__(f"{message} (backend: {backend})")


class GreeksError(Exception):
    def __init__(self, message: str = "An error occurred in Greeks calculations."):
        super().__init__(message)
