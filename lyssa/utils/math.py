from lyssandra_b200.utils.math import *  # noqa: F401,F403
