from lyssandra_b200.dict_learning.utils import *  # noqa: F401,F403
