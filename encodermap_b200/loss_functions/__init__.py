from .loss_functions import cartesian_distance_loss, distance_loss, sigmoid_loss  # noqa: F401
