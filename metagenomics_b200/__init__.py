"""B200-native overlap-graph construction engine (drop-in for the Dataset -> HashTable -> OverlapGraph
path of abiswas-odu/metagenomics). The compute path is libogb.so (CUDA, sm_100a); see include/ogb.h."""
from .api import Context, Dataset, HashTable, OverlapGraph, OgbError, edges_as_tuples, nccl_unique_id  # noqa: F401
