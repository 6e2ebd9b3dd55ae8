"""Host-side math that the scene loader needs (transforms, unit geometry)."""
