"""Drop-in for the query path of backend/engine.py: run_image_query (:46-65).

The Flask plumbing, thumbnail encoding and descriptor extraction (:68-137) are outside the
retrieval core; what is kept is the search call contract:
  * torch tensors are flattened to one (1, d) query (:49-50);
  * ``normalize=True`` L2-normalises the query in place first (:52-53; no reference caller sets it);
  * ``index.search(q, n_images)`` -> flattened distance / id lists (:55-57); inner-product indexes
    return descending scores, L2 indexes ascending squared distances;
  * every id is mapped through ``images_paths`` (:61) -- ids of -1 (k > ntotal)
    are dropped here instead of silently aliasing ``images_paths[-1]``.
The reference reads module globals ``index`` / ``images_paths`` (set under __main__, :110-135); they
can be set on this module the same way or passed explicitly.
"""
from __future__ import annotations

import numpy as np
import torch

from . import faiss_compat as faiss

index = None
images_paths = None
get_image = None  # optional callable path -> base64 thumbnail (utils.py:44-62, UI helper, out of scope)


def run_image_query(image_features, n_images, normalize=False, *, index=None, images_paths=None,
                    get_image=None):
    idx = index if index is not None else globals()["index"]
    paths = images_paths if images_paths is not None else globals()["images_paths"]
    thumb = get_image if get_image is not None else globals()["get_image"]
    if idx is None:
        raise RuntimeError("run_image_query: no index loaded")
    if isinstance(image_features, torch.Tensor) and not image_features.is_cuda:
        image_features = image_features.detach().cpu().numpy().reshape(1, -1)
    elif isinstance(image_features, torch.Tensor):
        image_features = image_features.detach().reshape(1, -1).to(torch.float32)
    else:
        image_features = np.asarray(image_features)
    if normalize:
        faiss.normalize_L2(image_features)
    distances, indices = idx.search(image_features, n_images)
    if isinstance(distances, torch.Tensor):
        distances, indices = distances.cpu().numpy(), indices.cpu().numpy()
    predictions = []
    for dist, i in zip(distances.ravel().tolist(), indices.ravel().tolist()):
        if i < 0:
            continue
        path = paths[i] if paths is not None else i
        predictions.append((dist, thumb(path) if thumb else None, str(path)))
    return predictions


def query_index(embedding, index, index_type, n_results):
    """Drop-in for backend/siamese/test_index.py:query_index (:49-71).

    "faiss": L2-normalise the query in place and search the (inner-product) index (:52-56).
    "dict" : the reference's fallback over a pickled [n, d] array -- query divided by its norm, Euclidean
             distance to every row, ascending (:58-69); here the array is searched on the device through an
             L2 index and the square roots are taken on the k winners.
    Returns (indices, distances) like the reference.
    """
    if index_type == "faiss":
        faiss.normalize_L2(embedding)
        distances, indices = index.search(embedding, n_results)
        return indices.ravel().tolist(), distances.ravel().tolist()
    if index_type == "dict":
        q = np.asarray(embedding, dtype=np.float64)
        q = (q / np.linalg.norm(q)).astype(np.float32).reshape(1, -1)
        flat = faiss.IndexFlatL2(q.shape[1])
        flat.add(np.ascontiguousarray(np.asarray(index), dtype=np.float32))
        d2, ids = flat.search(q, n_results)
        keep = ids.ravel() >= 0
        return ids.ravel()[keep], np.sqrt(d2.ravel()[keep])
    raise ValueError(f"unknown index_type {index_type!r}")
