"""Exceptions of the public API (same names, attributes and messages as the reference's
machupX/exceptions.py:4-65 so that callers' ``except`` clauses and message checks keep working)."""


class SolverNotConvergedError(Exception):
    """The Newton solver (or scipy's fsolve) stopped before reaching the convergence threshold.

    Attributes: ``solver_type`` ("nonlinear" or "scipy_fsolve"), ``final_error`` (last residual norm), ``message``.
    """

    def __init__(self, solver_type, final_error):
        self.solver_type = solver_type
        self.final_error = final_error
        advice = " Try relaxing the solver, increasing the maximum iterations, or increasing the convergence threshold."
        if solver_type == "scipy_fsolve":
            self.message = ("The scipy solver failed to converge, after which the nonlinear solver also failed to converge."
                            + advice)
        else:
            self.message = "The nonlinear solver failed to converge." + advice
        super().__init__(self.message)

    def __str__(self):
        return "{0} The final error was {1}.".format(self.message, self.final_error)


class MaxIterationError(Exception):
    """An outer iteration (pitch trim, target CL, ...) ran out of iterations.

    Attributes: ``throwing_function``, ``final_error``, ``message``.
    """

    def __init__(self, throwing_function, final_error):
        self.throwing_function = throwing_function
        self.final_error = final_error
        self.message = ("The solver within {0} failed to converge. The target state may not be attainable on this "
                        "aircraft.".format(throwing_function))
        super().__init__(self.message)

    def __str__(self):
        return "{0} The final error was {1}.".format(self.message, self.final_error)
