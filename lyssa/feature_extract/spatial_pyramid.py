from lyssandra_b200.feature_extract.spatial_pyramid import sc_spm_extractor, dsift_extractor, pyramid_feat_extract  # noqa: F401
