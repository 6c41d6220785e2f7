"""TEST INFRASTRUCTURE ONLY (see package docstring)."""
colormaps = {}
