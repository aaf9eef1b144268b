class DatabaseBoundsError(Exception):
    """Raised by the real airfoil_db when a lookup leaves the data table; never raised by the linear model."""

    def __init__(self, airfoil="", exception_indices=(), inputs_dict=None):
        self.airfoil = airfoil
        self.exception_indices = exception_indices
        self.inputs_dict = inputs_dict or {}
        super().__init__("The inputs to the airfoil database fell outside the bounds of available data.")
