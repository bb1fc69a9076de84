from lyssandra_b200.dict_learning.online_dict_learn import *  # noqa: F401,F403
