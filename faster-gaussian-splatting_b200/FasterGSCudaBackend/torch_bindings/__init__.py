from fgs_b200.torch_bindings import *  # noqa: F401,F403
from fgs_b200.torch_bindings import __all__  # noqa: F401
