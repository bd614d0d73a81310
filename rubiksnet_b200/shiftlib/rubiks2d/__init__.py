from .layer import *  # noqa: F401,F403
from .primitive import *  # noqa: F401,F403
