"""Import-time placeholder for liesel (type names only) — test infrastructure."""
