from .irlb import lanczos, LanczosResult  # noqa: F401
