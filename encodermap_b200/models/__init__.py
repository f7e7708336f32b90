from .layers import BackMapLayer, PairwiseDistances, PeriodicInput  # noqa: F401
