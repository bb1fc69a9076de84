"""Mirror of lyssa/dict_learning/__init__.py:1-5 for the learners on the hot path."""
from .ksvd import ksvd_coder, ksvd_dict_learn, approx_ksvd, ksvd  # noqa: F401
from .online_dict_learn import online_dictionary_coder, online_dict_learn  # noqa: F401
from .gradient_descent import dictionary_learner  # noqa: F401
from .class_dict_learn import class_ksvd_coder, class_dict_learn  # noqa: F401
from .utils import init_dictionary, approx_error  # noqa: F401
