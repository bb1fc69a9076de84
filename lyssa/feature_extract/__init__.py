"""lyssa.feature_extract -> lyssandra_b200.feature_extract (reference: lyssa/feature_extract/{spatial_pyramid,pooling,preproc}.py)."""
from lyssandra_b200.feature_extract import (sc_spm_extractor, dsift_extractor, pyramid_feat_extract, sc_max_pooling, max_pooling,  # noqa: F401
                                            sum_pooling, average_pooling, l2_normalizer)
