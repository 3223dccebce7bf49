"""TEST INFRASTRUCTURE ONLY -- container-only stand-in for `gym` (no arithmetic)."""
from . import spaces, utils, envs  # noqa
from .envs.registration import register, make  # noqa


class Env(object):
    pass
