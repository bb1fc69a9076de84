from lyssandra_b200.feature_extract.preproc import l2_normalizer  # noqa: F401
