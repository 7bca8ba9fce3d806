"""image_search_engine_b200 -- B200-native retrieval core of ManuelZ/image-search-engine.

Hot path only: k-means codebook training, descriptor quantisation + BoVW histogram /
Okapi tf, and flat kNN search, behind the reference's own Python call surface:

    FaissKMeans                                   (backend/kmeans_faiss.py)
    BOVW, run_clustering, load_cluster_model,
    train_bovw_model                              (backend/bag_of_visual_words.py)
    OkapiTransformer, create_search_index, chunkIt,
    calc_sampled_cluster_score                    (backend/utils.py)
    run_image_query (+ QueryBatcher)              (backend/engine.py)
    query_index                                   (backend/siamese/test_index.py)
    faiss_compat                                  (the subset of the `faiss` module those files call)

All arithmetic runs in hand-written sm_100a kernels exported by libise.so (include/ise.h).
Importing the package never touches the GPU; the first compute call requires the built
extension and a B200, and raises otherwise -- there is no CPU fallback.
"""
from . import faiss_compat
from ._lib import IseError
from .bag_of_visual_words import (BOVW, PackedDescriptions, load_cluster_model, pack_descriptions, run_clustering,
                                   train_bovw_model)
from .engine import QueryBatcher, query_index, run_image_query
from .kmeans_faiss import FaissKMeans
from .utils import OkapiTransformer, calc_sampled_cluster_score, chunkIt, create_search_index

__all__ = ["faiss_compat", "IseError", "BOVW", "load_cluster_model", "run_clustering", "train_bovw_model",
           "run_image_query", "query_index", "QueryBatcher", "FaissKMeans", "OkapiTransformer", "chunkIt",
           "create_search_index", "calc_sampled_cluster_score", "pack_descriptions", "PackedDescriptions"]
