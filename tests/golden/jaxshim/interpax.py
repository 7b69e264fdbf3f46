"""Import-time placeholder for interpax (only the off-path inverse uses it)."""


def interp1d(*_a, **_k):
    raise NotImplementedError("interpax is not available in this image")
