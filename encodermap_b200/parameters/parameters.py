"""Attribute bags with the reference's defaults for the knobs that parameterise the hot path.

Reference: encodermap/parameters/parameters.py:611-638 (``Parameters._defaults``) and :794-828
(``ADCParameters._defaults``).  Only the fields the operator layer reads are kept; the reference's
own ``Parameters`` objects work unchanged wherever one of these is accepted (duck typing)."""
from __future__ import annotations

from math import pi
from typing import Any


class Parameters:
    _defaults = dict(
        n_neurons=[128, 128, 2],
        activation_functions=["", "tanh", "tanh", ""],
        periodicity=2 * pi,
        learning_rate=0.001,
        n_steps=1000,
        batch_size=256,
        dist_sig_parameters=(4.5, 12, 6, 1, 2, 6),
        distance_cost_scale=500,
        auto_cost_scale=1,
        auto_cost_variant="mean_abs",
        center_cost_scale=0.0001,
        l2_reg_constant=0.001,
    )

    def __init__(self, **kwargs: Any) -> None:
        for k, v in self._defaults.items():
            setattr(self, k, v)
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def defaults(self):
        return dict(self._defaults)


class ADCParameters(Parameters):
    _defaults = dict(
        Parameters._defaults,
        cartesian_pwd_start=None,
        cartesian_pwd_stop=None,
        cartesian_pwd_step=None,
        use_backbone_angles=False,
        use_sidechains=False,
        cartesian_cost_scale=1,
        cartesian_cost_variant="mean_abs",
        cartesian_cost_reference=1,
        cartesian_dist_sig_parameters=Parameters._defaults["dist_sig_parameters"],
        cartesian_distance_cost_scale=1,
        auto_cost_scale=None,
        distance_cost_scale=None,
        reconstruct_sidechains=False,
    )
