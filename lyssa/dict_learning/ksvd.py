from lyssandra_b200.dict_learning.ksvd import *  # noqa: F401,F403
