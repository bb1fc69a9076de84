"""lyssa.feature_encoding -> lyssandra_b200.feature_encoding (reference: lyssa/feature_encoding.py:14-89)."""
from lyssandra_b200.feature_encoding import feature_encoder, soft_thresholding, sign_splitting  # noqa: F401
