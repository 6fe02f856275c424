"""Plugin surface mirroring the reference's `library/` package (bodies, geometry, assets,
environment, actions, observations, mods)."""
