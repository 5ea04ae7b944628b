"""Mirror of ``xpsi.cellmesh`` integrators (GPU)."""
