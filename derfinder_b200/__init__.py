"""derfinder_b200 -- B200-native engine for derfinder's single-base F-statistic pipeline.

The product is the CUDA shared library `libderfinder_b200.so` (C ABI: include/derfinder_b200.h);
`derfinder_b200.api` mirrors the reference's R functions on top of it.
"""
from .api import (Coverage, Engine, analyzeChr, calcFstatCutoff, calcPval, calculateFWER, calculatePvalues,  # noqa: F401
                  calculateStats, clusterMakerRle, coverageToExon, default_engine, filterData, findRegions, getSegmentsRle,
                  makeModels, mergeResults, preprocessCoverage, railMatrix, regionMatrix)

__version__ = "0.1.0"
