from lyssandra_b200.feature_extract.dsift import DsiftExtractor, gen_dgauss  # noqa: F401
