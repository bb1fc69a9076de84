"""lyssa.dict_learning.class_dict_learn -> lyssandra_b200 (reference: lyssa/dict_learning/class_dict_learn.py)."""
from lyssandra_b200.dict_learning.class_dict_learn import class_ksvd_coder, class_dict_learn  # noqa: F401
