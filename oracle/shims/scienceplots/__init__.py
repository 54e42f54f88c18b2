"""Empty stand-in for scienceplots (reference use: dgfem/visualization.py:19)."""
