"""Import stand-in for `thop` (only imported, never used, by the reference's transformer files)."""


def profile(*args, **kwargs):
    raise NotImplementedError("thop stand-in")


def clever_format(*args, **kwargs):
    raise NotImplementedError("thop stand-in")
