"""lyssa.dict_learning -> lyssandra_b200.dict_learning (reference: lyssa/dict_learning/__init__.py:1-5)."""
from lyssandra_b200.dict_learning import (ksvd_coder, ksvd_dict_learn, approx_ksvd,  # noqa: F401
                                          online_dictionary_coder, online_dict_learn,
                                          dictionary_learner, init_dictionary, approx_error,
                                          class_ksvd_coder, class_dict_learn)
