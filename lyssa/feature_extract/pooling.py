from lyssandra_b200.feature_extract.pooling import sc_max_pooling, max_pooling, sum_pooling, average_pooling  # noqa: F401
