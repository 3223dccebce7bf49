from . import registration  # noqa
