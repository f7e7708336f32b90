from .parameters import ADCParameters, Parameters  # noqa: F401
