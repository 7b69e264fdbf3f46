"""Import-time placeholder for liesel.contrib.splines (the in-tree twin bspline/approx.py is
what the fixture generator runs)."""
