from .backmapping import (  # noqa: F401
    dihedral_to_cartesian_tf_one_way_layers,
    dihedrals_to_cartesian_tf_layers,
    rotation_matrix,
    split_and_reverse_cartesians,
    split_and_reverse_dihedrals,
)
from .distances import pairwise_dist, pairwise_dist_periodic, periodic_distance, sigmoid  # noqa: F401
