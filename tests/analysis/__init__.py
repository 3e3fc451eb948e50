"""Analysis scripts that use the CPU oracle (test infrastructure): cull statistics and per-tile list statistics of the
bench scene.  Not collected by pytest."""
