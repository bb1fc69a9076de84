from lyssandra_b200.dict_learning.gradient_descent import *  # noqa: F401,F403
