"""ScSPM pooling behind the reference's feature_extract API (first "next" row)."""
from .pooling import sc_max_pooling, max_pooling, sum_pooling, average_pooling  # noqa: F401
from .preproc import l2_normalizer  # noqa: F401
from .spatial_pyramid import sc_spm_extractor, dsift_extractor, spm_pool, pyramid_feat_extract  # noqa: F401
from .dsift import DsiftExtractor  # noqa: F401
