from .backmapping import chain_in_plane, dihedral_to_cartesian_tf_one_way, dihedrals_to_cartesian_tf  # noqa: F401
