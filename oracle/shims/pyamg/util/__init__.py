from . import utils, params, linalg  # noqa: F401
